// host_math.hpp — double-precision 4x4 helpers used at scene/source set-up time by the engine
// (Fortran array layout: element (i,j), 1-based, at m[(j-1)*4 + (i-1)]; row-vector convention p' = p.M,
// src/vector_class.f90:292-304).  Mirrors the semantics of src/sdfs/sdfHelpers.f90 and src/mat_class.f90.
#pragma once
#include <cmath>

namespace smcrt_math {

struct M44 {
    double m[16];
    double& a(int i, int j) { return m[(j - 1) * 4 + (i - 1)]; }
    double a(int i, int j) const { return m[(j - 1) * 4 + (i - 1)]; }
};
inline M44 identity() {
    M44 r{};
    for (int i = 1; i <= 4; ++i) r.a(i, i) = 1.0;
    return r;
}
inline M44 translate(double x, double y, double z) {  // sdfHelpers.f90:168-182 (o in row 4)
    M44 r = identity();
    r.a(4, 1) = x; r.a(4, 2) = y; r.a(4, 3) = z;
    return r;
}
inline M44 matmul(const M44& x, const M44& y) {
    M44 r{};
    for (int i = 1; i <= 4; ++i)
        for (int j = 1; j <= 4; ++j) {
            double s = 0;
            for (int k = 1; k <= 4; ++k) s += x.a(i, k) * y.a(k, j);
            r.a(i, j) = s;
        }
    return r;
}
// rotationAlign(a,b) = I + [v]x + [v]x^2/(1+a.b), v = a x b  (sdfHelpers.f90:114-140; the skew matrix is
// filled column by column there: v_x(:,1) = [0,-vz,vy,0] ...)
inline M44 rotation_align(const double a[3], const double b[3]) {
    const double v[3] = {a[1] * b[2] - a[2] * b[1], -a[0] * b[2] + a[2] * b[0], a[0] * b[1] - a[1] * b[0]};
    const double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2], k = 1.0 / (1.0 + c);
    M44 vx{};
    vx.a(1, 1) = 0;      vx.a(2, 1) = -v[2];  vx.a(3, 1) = v[1];
    vx.a(1, 2) = v[2];   vx.a(2, 2) = 0;      vx.a(3, 2) = -v[0];
    vx.a(1, 3) = -v[1];  vx.a(2, 3) = v[0];   vx.a(3, 3) = 0;
    M44 vx2 = matmul(vx, vx), r = identity();
    for (int i = 0; i < 16; ++i) r.m[i] += vx.m[i] + vx2.m[i] * k;
    return r;
}
// inverse of an affine row-vector matrix [R 0; t 1]: [R^-1 0; -t R^-1 1] with a general 3x3 inverse
inline M44 invert_affine(const M44& in) {
    const double a = in.a(1, 1), b = in.a(1, 2), c = in.a(1, 3), d = in.a(2, 1), e = in.a(2, 2), f = in.a(2, 3),
                 g = in.a(3, 1), h = in.a(3, 2), i = in.a(3, 3);
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C, id = 1.0 / det;
    M44 r = identity();
    r.a(1, 1) = A * id;                  r.a(1, 2) = -(b * i - c * h) * id;  r.a(1, 3) = (b * f - c * e) * id;
    r.a(2, 1) = B * id;                  r.a(2, 2) = (a * i - c * g) * id;   r.a(2, 3) = -(a * f - c * d) * id;
    r.a(3, 1) = C * id;                  r.a(3, 2) = -(a * h - b * g) * id;  r.a(3, 3) = (a * e - b * d) * id;
    for (int j = 1; j <= 3; ++j)
        r.a(4, j) = -(in.a(4, 1) * r.a(1, j) + in.a(4, 2) * r.a(2, j) + in.a(4, 3) * r.a(3, j));
    return r;
}
inline void normalise(double v[3]) {
    const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= l; v[1] /= l; v[2] /= l;
}

}  // namespace smcrt_math
