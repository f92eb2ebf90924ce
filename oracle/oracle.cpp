// ============================================================================================
// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// A double-precision CPU restatement (C++17 + OpenMP) of the photon-packet hot path of
// the-professor510/RSMCRT (fork of signedMCRT).  It exists only so tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs can check and time-compare the sm_100a engine.
// Nothing under rsmcrt_b200/ may include, link, import or execute it.
//
// Why a restatement: the reference is Fortran 2018 with un-vendored, un-pinned git dependencies
// (fpm.toml:8-16) and this image has neither a Fortran compiler nor fpm, so the reference cannot be
// compiled or imported here.  Parity of this oracle with the reference is pinned by
// transcribing the reference's own known-answer tests (test/SDF, test/fresnel, test/detector,
// test/geometry, test/photon, test/matrix, test/vector, test/end_to_end) and the literature targets in
// tools/validate*.py into tests/test_oracle_*.py.  The RNG stream itself is unpinned BY DESIGN in the
// reference (compiler intrinsic random_number, src/random_mod.f90:83-90; its tests pin only ranges /
// moments), so whole-simulation parity is statistical.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// The loop structure deliberately mirrors the reference (nested while loops, ds(:) arrays) so that it
// is an independent check of the engine's state-machine formulation.
// ============================================================================================
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- constants (src/constants.f90:18-30)
constexpr double PI = 3.14159265358979323846;
constexpr double TWOPI = 2.0 * PI;
constexpr double REF_THRESHOLD = 0.01;
constexpr double REF_CHANCE = 0.1;

// node / source / detector codes: numerically identical to include/smcrt.h (the interface contract),
// restated here so the oracle does not depend on product headers.
enum { K_SPHERE = 1, K_BOX, K_TORUS, K_CYLINDER, K_TRIPRISM, K_SEGMENT, K_CAPSULE, K_CONE, K_EGG, K_PLANE,
       K_UNION = 20, K_SMOOTHUNION, K_SUBTRACTION, K_INTERSECTION,
       K_REVOLUTION = 30, K_EXTRUDE, K_ONION, K_TWIST, K_BEND, K_ELONGATE };
enum { SRC_POINT = 1, SRC_PENCIL, SRC_UNIFORM, SRC_CIRCULAR, SRC_FOCUS, SRC_ANNULUS, SRC_DSLIT, SRC_APERTURE };
enum { SP_POS = 0, SP_DIR = 3, SP_P1 = 6, SP_P2 = 9, SP_P3 = 12, SP_RADIUS = 15, SP_FOCAL = 16, SP_BEAM = 17,
       SP_RLO = 18, SP_RHI = 19, SP_SIGMA = 20, SP_ROT = 21, SP_N = 24 };
enum { DET_CIRCLE = 1, DET_ANNULUS, DET_FIBRE, DET_CAMERA };
constexpr int DET_P = 20, NODE_P = 8;
enum { TALLY_ABSORB = 1, TALLY_PATHLENGTH = 2, TALLY_EMISSION = 4 };

// ---------------------------------------------------------------- vector (src/vector_class.f90)
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }           // :280-290
inline V3 cross(V3 a, V3 b) {                                                          // :306-318
    return {a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x};
}
inline double length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }    // :405-411
inline V3 magnitude(V3 a) { return a / length(a); }                                    // :392-402 (normalises!)
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }      // :157-164
inline V3 vmax0(V3 a) { return {std::max(a.x, 0.0), std::max(a.y, 0.0), std::max(a.z, 0.0)}; }
inline bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }         // :131-146

// 4x4 matrix stored exactly like the Fortran array: element (i,j) (1-based) at m[(j-1)*4 + (i-1)]
struct M4 {
    double m[16];
    double& at(int i, int j) { return m[(j - 1) * 4 + (i - 1)]; }
    double at(int i, int j) const { return m[(j - 1) * 4 + (i - 1)]; }
};
// vec .dot. mat, row-vector affine (src/vector_class.f90:292-304)
inline V3 vec_dot_mat(V3 a, const M4& b) {
    return {b.at(1, 1) * a.x + b.at(2, 1) * a.y + b.at(3, 1) * a.z + b.at(4, 1),
            b.at(1, 2) * a.x + b.at(2, 2) * a.y + b.at(3, 2) * a.z + b.at(4, 2),
            b.at(1, 3) * a.x + b.at(2, 3) * a.y + b.at(3, 3) * a.z + b.at(4, 3)};
}
M4 m_identity() {  // src/sdfs/sdfHelpers.f90:142-152
    M4 r{};
    for (int i = 1; i <= 4; ++i) r.at(i, i) = 1.0;
    return r;
}
M4 m_matmul(const M4& a, const M4& b) {  // Fortran matmul
    M4 r{};
    for (int i = 1; i <= 4; ++i)
        for (int j = 1; j <= 4; ++j) {
            double s = 0;
            for (int k = 1; k <= 4; ++k) s += a.at(i, k) * b.at(k, j);
            r.at(i, j) = s;
        }
    return r;
}
inline double deg2rad(double a) { return a * PI / 180.0; }  // utils:deg2rad, pinned by test_rotate_* KATs
// r(:,1) = [..] assigns COLUMN 1 (src/sdfs/sdfHelpers.f90:23-83)
M4 m_from_cols(const double c1[4], const double c2[4], const double c3[4], const double c4[4]) {
    M4 r{};
    for (int i = 0; i < 4; ++i) {
        r.m[0 + i] = c1[i];
        r.m[4 + i] = c2[i];
        r.m[8 + i] = c3[i];
        r.m[12 + i] = c4[i];
    }
    return r;
}
M4 m_rotate_x(double angle) {  // sdfHelpers.f90:23-41
    double a = deg2rad(angle), c = std::cos(a), s = std::sin(a);
    double c1[4] = {1, 0, 0, 0}, c2[4] = {0, c, -s, 0}, c3[4] = {0, s, c, 0}, c4[4] = {0, 0, 0, 1};
    return m_from_cols(c1, c2, c3, c4);
}
M4 m_rotate_y(double angle) {  // sdfHelpers.f90:43-62
    double a = deg2rad(angle), c = std::cos(a), s = std::sin(a);
    double c1[4] = {c, 0, s, 0}, c2[4] = {0, 1, 0, 0}, c3[4] = {-s, 0, c, 0}, c4[4] = {0, 0, 0, 1};
    return m_from_cols(c1, c2, c3, c4);
}
M4 m_rotate_z(double angle) {  // sdfHelpers.f90:64-83
    double a = deg2rad(angle), c = std::cos(a), s = std::sin(a);
    double c1[4] = {c, -s, 0, 0}, c2[4] = {s, c, 0, 0}, c3[4] = {0, 0, 1, 0}, c4[4] = {0, 0, 0, 1};
    return m_from_cols(c1, c2, c3, c4);
}
M4 m_rotmat(V3 axis, double angle) {  // sdfHelpers.f90:85-112
    V3 ax = magnitude(axis);
    double a = deg2rad(angle), s = std::sin(a), c = std::cos(a), oc = 1.0 - c;
    double c1[4] = {oc * ax.x * ax.x + c, oc * ax.x * ax.y - ax.z * s, oc * ax.z * ax.x + ax.y * s, 0};
    double c2[4] = {oc * ax.x * ax.y + ax.z * s, oc * ax.y * ax.y + c, oc * ax.y * ax.z - ax.x * s, 0};
    double c3[4] = {oc * ax.z * ax.x - ax.y * s, oc * ax.y * ax.z + ax.x * s, oc * ax.z * ax.z + c, 0};
    double c4[4] = {0, 0, 0, 1};
    return m_from_cols(c1, c2, c3, c4);
}
M4 m_skew(V3 a) {  // sdfHelpers.f90:154-166
    double c1[4] = {0, -a.z, a.y, 0}, c2[4] = {a.z, 0, -a.x, 0}, c3[4] = {-a.y, a.x, 0, 0}, c4[4] = {0, 0, 0, 0};
    return m_from_cols(c1, c2, c3, c4);
}
M4 m_rotation_align(V3 a, V3 b) {  // sdfHelpers.f90:114-140  (I + [v]x + [v]x^2 / (1 + a.b))
    V3 v = cross(a, b);
    double c = dot(a, b), k = 1.0 / (1.0 + c);
    M4 vx = m_skew(v), vx2 = m_matmul(vx, vx), r = m_identity();
    for (int i = 0; i < 16; ++i) r.m[i] += vx.m[i] + vx2.m[i] * k;
    return r;
}
M4 m_translate(V3 o) {  // sdfHelpers.f90:168-182 : o goes into ROW 4
    double c1[4] = {1, 0, 0, o.x}, c2[4] = {0, 1, 0, o.y}, c3[4] = {0, 0, 1, o.z}, c4[4] = {0, 0, 0, 1};
    return m_from_cols(c1, c2, c3, c4);
}
// general 4x4 inverse (src/mat_class.f90:154-214 computes the adjugate / determinant directly; here the
// same closed form is reached through 2x2 sub-determinants of row pairs, which is the standard
// Laplace-expansion arrangement)
M4 m_invert(const M4& A) {
    const double* a = A.m;  // treat as a[col*4+row]; inverse of a matrix is layout-agnostic if we are consistent
    double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
    double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
    double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
    double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
    double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    double id = 1.0 / det;
    M4 B{};
    double* b = B.m;
    b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * id;
    b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * id;
    b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * id;
    b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * id;
    b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * id;
    b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * id;
    b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * id;
    b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * id;
    b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * id;
    b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * id;
    b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * id;
    b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * id;
    b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * id;
    b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * id;
    b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * id;
    b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * id;
    return B;
}
inline double clampd(double v, double lo, double hi) { return std::min(std::max(v, lo), hi); }  // utils:clamp
inline double fsign(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }      // Fortran sign()

// ---------------------------------------------------------------- scene
struct Optics {  // mono, src/opticalProps/opticalProperties.f90:107-125
    double mus, mua, hgg, g2, n, kappa, albedo;
    void init(double mus_, double mua_, double hgg_, double n_) {
        mus = mus_;
        mua = mua_;
        kappa = mus + mua;
        albedo = (mua < 1e-9) ? 1.0 : mus / kappa;
        hgg = hgg_;
        g2 = hgg_ * hgg_;
        n = n_;
    }
};
struct Node {
    int kind, first_child, n_child;
    M4 xf;
    double p[NODE_P];
};
struct Grid {  // cart_grid, src/grid.f90:14-25,119-159
    int nxg = 0, nyg = 0, nzg = 0;
    double xmax = 1, ymax = 1, zmax = 1;
    std::vector<double> xface, yface, zface;
    void init(int nx, int ny, int nz, double xm, double ym, double zm) {
        nxg = nx; nyg = ny; nzg = nz; xmax = xm; ymax = ym; zmax = zm;
        xface.resize(nx + 1); yface.resize(ny + 1); zface.resize(nz + 2);
        for (int i = 1; i <= nx + 1; ++i) xface[i - 1] = (i - 1) * 2.0 * xm / nx;
        for (int i = 1; i <= ny + 1; ++i) yface[i - 1] = (i - 1) * 2.0 * ym / ny;
        for (int i = 1; i <= nz + 2; ++i) zface[i - 1] = (i - 1) * 2.0 * zm / nz;
    }
    // get_voxel_cart, src/grid.f90:51-78
    void get_voxel(V3 pos, int cell[3]) const {
        cell[0] = (int)std::floor(nxg * (pos.x + xmax) / (2.0 * xmax)) + 1;
        cell[1] = (int)std::floor(nyg * (pos.y + ymax) / (2.0 * ymax)) + 1;
        cell[2] = (int)std::floor(nzg * (pos.z + zmax) / (2.0 * zmax)) + 1;
        if (cell[0] < 1 || cell[0] > nxg) cell[0] = -1;
        if (cell[1] < 1 || cell[1] > nyg) cell[1] = -1;
        if (cell[2] < 1 || cell[2] > nzg) cell[2] = -1;
    }
};
struct Detector {  // src/detectors/detectors.f90 init_* :107-445
    int kind;
    V3 pos, dir;
    double radius = 0, r1 = 0, r2 = 0;
    double f1 = 0, f2 = 0, f1Ap = 0, f2Ap = 0, frontOff = 0, backOff = 0, frontToPin = 0, pinToBack = 0, pinAp = 0,
           acceptAngle = 0, coreDiameter = 0;
    V3 p2, p3, e1, e2, n;  // camera
    double width = 0, height = 0;
    int nbins = 0, nbinsX = 0, nbinsY = 0;  // STORED counts (= user nbins + 1)
    double bin_wid = 1, bin_wid_x = 1, bin_wid_y = 1;
    int64_t offset = 0;  // into the concatenated bins array
    int64_t count() const { return kind == DET_CAMERA ? (int64_t)nbinsX * nbinsY : nbins; }
};
struct Source {
    int kind = SRC_POINT, subtype = 0;
    double p[SP_N] = {0};
};
struct Scene {
    std::vector<Node> nodes;
    std::vector<int> top;
    std::vector<Optics> opt;
    Grid grid;
    Source src;
    std::vector<Detector> dets;
    int64_t det_total = 0;
    bool bugcompat = true;      // keep quirks Q1..Q5
    bool launch_mask_le = false;  // test_kernel's mask=(distances<=0), kernelsMod.f90:2136
};

// ---------------------------------------------------------------- SDF evaluation
double eval_node(const Scene& s, int ni, V3 pos);

// primitives, src/sdfs/sdfs.f90:494-735
double eval_prim(const Node& nd, V3 pos) {
    V3 p = vec_dot_mat(pos, nd.xf);
    const double* q = nd.p;
    switch (nd.kind) {
        case K_SPHERE:  // :494-508
            return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z) - q[0];
        case K_BOX: {  // :510-525
            V3 d = vabs(p) - V3{q[0], q[1], q[2]};
            return length(vmax0(d)) + std::min(std::max(d.x, std::max(d.y, d.z)), 0.0);
        }
        case K_TORUS: {  // :527-542
            V3 t{length(V3{p.x, 0.0, p.z}) - q[0], p.y, 0.0};
            return length(t) - q[1];
        }
        case K_CYLINDER: {  // :544-581
            V3 a{q[0], q[1], q[2]}, b{q[3], q[4], q[5]};
            double radius = q[6];
            V3 ba = b - a, pa = p - a;
            double baba = dot(ba, ba), paba = dot(pa, ba);
            double x = length(pa * baba - ba * paba) - radius * baba;
            double y = std::fabs(paba - baba * 0.5) - baba * 0.5;
            double x2 = x * x, y2 = (y * y) * baba, d;
            if (std::max(x, y) < 0.0) {
                d = -std::min(x2, y2);
            } else {
                if (x > 0.0 && y > 0.0) d = x2 + y2;
                else if (x > 0.0) d = x2;
                else if (y > 0.0) d = y2;
                else d = 0.0;
            }
            return fsign(std::sqrt(std::fabs(d)) / baba, d);
        }
        case K_TRIPRISM: {  // :583-597
            V3 a = vabs(p);
            return std::max(a.z - q[1], std::max(a.x * 0.866025 + p.y * 0.5, -p.y) - q[0] * 0.5);
        }
        case K_SEGMENT:    // :599-626 (radius fixed 0.1)
        case K_CAPSULE: {  // :628-648
            V3 a{q[0], q[1], q[2]}, b{q[3], q[4], q[5]};
            V3 pa = p - a, ba = b - a;
            double h = clampd(dot(pa, ba) / dot(ba, ba), 0.0, 1.0);
            return length(pa - ba * h) - (nd.kind == K_SEGMENT ? 0.1 : q[6]);
        }
        case K_CONE: {  // :650-686
            V3 a{q[0], q[1], q[2]}, b{q[3], q[4], q[5]};
            double ra = q[6], rb = q[7];
            double rba = rb - ra;
            double baba = dot(b - a, b - a);
            double papa = dot(p - a, p - a);
            double paba = dot(p - a, b - a) / baba;
            double x = std::sqrt(papa - baba * paba * paba);
            double cax = (paba < 0.5) ? std::max(0.0, x - ra) : std::max(0.0, x - rb);
            double cay = std::fabs(paba - 0.5) - 0.5;
            double k = rba * rba + baba;
            double f = clampd((rba * (x - ra) + paba * baba) / k, 0.0, 1.0);
            double cbx = x - ra - f * rba;
            double cby = paba - f;
            double sgn = (cbx < 0.0 && cay < 0.0) ? -1.0 : 1.0;
            return sgn * std::sqrt(std::min(cax * cax + baba * cay * cay, cbx * cbx + baba * cby * cby));
        }
        case K_EGG: {  // :688-718
            double r1 = q[0], r2 = q[1], h = q[2];
            V3 pin = p;
            pin.x = std::fabs(p.x);
            double r = r1 - r2;
            double hin = h + r;
            double l = (hin * hin - r * r) / (2.0 * r);
            if (pin.y <= 0.0) return length(pin) - r1;
            if ((pin.y - hin) * l > pin.x * hin)
                return length(pin - V3{0.0, hin, 0.0}) - ((r1 + l) - length(V3{hin, l, 0.0}));
            return length(pin + V3{l, 0.0, 0.0}) - (r1 + l);
        }
        case K_PLANE:  // :720-735
            return dot(p, V3{q[0], q[1], q[2]});
    }
    return std::numeric_limits<double>::quiet_NaN();
}

// CSG ops, src/sdfs/sdfModifiers.f90:428-491
inline double csg_op(int kind, double d1, double d2, double k) {
    switch (kind) {
        case K_UNION: return std::min(d1, d2);
        case K_SMOOTHUNION: {
            double h = std::max(k - std::fabs(d1 - d2), 0.0) / k;
            return std::min(d1, d2) - h * h * h * k * (1.0 / 6.0);
        }
        case K_SUBTRACTION: return std::max(-d1, d2);
        case K_INTERSECTION: return std::max(d1, d2);
    }
    return d1;
}

double eval_node(const Scene& s, int ni, V3 pos) {
    const Node& nd = s.nodes[ni];
    if (nd.kind <= K_PLANE) return eval_prim(nd, pos);
    if (nd.kind >= K_UNION && nd.kind <= K_INTERSECTION) {  // eval_model, src/sdfs/sdf_base.f90:146-161
        double res = eval_node(s, nd.first_child, pos);
        for (int i = 1; i < nd.n_child; ++i) res = csg_op(nd.kind, res, eval_node(s, nd.first_child + i, pos), nd.p[0]);
        return res;
    }
    const int c = nd.first_child;
    switch (nd.kind) {
        case K_EXTRUDE: {  // sdfModifiers.f90:286-301
            double d = eval_node(s, c, pos);
            double wx = d, wy = std::fabs(pos.z) - nd.p[0];
            return std::min(std::max(wx, wy), 0.0) + length(V3{std::max(wx, 0.0), std::max(wy, 0.0), 0.0});
        }
        case K_REVOLUTION: {  // :303-321
            V3 pin = pos - V3{nd.p[1], nd.p[2], nd.p[3]};
            V3 q{length(V3{pin.x, 0.0, pin.z}) - nd.p[0], pin.y, 0.0};
            return eval_node(s, c, q);
        }
        case K_ONION:  // :323-333
            return std::fabs(eval_node(s, c, pos)) - nd.p[0];
        case K_ELONGATE: {  // :335-351
            V3 q = vabs(pos) - V3{nd.p[0], nd.p[1], nd.p[2]};
            double w = std::min(std::max(q.x, std::max(q.y, q.z)), 0.0);
            return eval_node(s, c, vmax0(q)) + w;
        }
        case K_TWIST: {  // :353-371
            double cc = std::cos(nd.p[0] * pos.z), ss = std::sin(nd.p[0] * pos.z);
            return eval_node(s, c, V3{cc * pos.x - ss * pos.y, ss * pos.x + cc * pos.y, pos.z});
        }
        case K_BEND: {  // :373-391
            double cc = std::cos(nd.p[0] * pos.x), ss = std::sin(nd.p[0] * pos.x);
            return eval_node(s, c, V3{cc * pos.x - ss * pos.y, ss * pos.x + cc * pos.y, pos.z});
        }
    }
    return std::numeric_limits<double>::quiet_NaN();
}
inline double eval_top(const Scene& s, int t /*0-based*/, V3 pos) { return eval_node(s, s.top[t], pos); }

// calcNormal, src/sdfs/sdf_base.f90:166-190 (tetrahedral 4-tap, h = 1e-6)
V3 calc_normal(const Scene& s, int t, V3 p) {
    const double h = 1e-6;
    V3 xyy{1, -1, -1}, yyx{-1, -1, 1}, yxy{-1, 1, -1}, xxx{1, 1, 1};
    V3 n = xyy * eval_top(s, t, p + xyy * h) + yyx * eval_top(s, t, p + yyx * h) + yxy * eval_top(s, t, p + yxy * h) +
           xxx * eval_top(s, t, p + xxx * h);
    return magnitude(n);
}

// maxloc(ds, dim=1, mask=(ds<0)) -> 1-based index, 0 when the mask is empty; ties -> lowest index
inline int maxloc_neg(const double* ds, int n, bool le = false) {
    int best = 0;
    double bv = 0;
    for (int i = 0; i < n; ++i) {
        bool m = le ? (ds[i] <= 0.0) : (ds[i] < 0.0);
        if (m && (best == 0 || ds[i] > bv)) {
            best = i + 1;
            bv = ds[i];
        }
    }
    return best;
}

// ---------------------------------------------------------------- RNG
// Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants).  The engine keys a packet's stream by
// (seed, packet id) and consumes ONE 4-word block per "event" (emit attempt / interaction / Fresnel).
struct Philox {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1],
                 n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // `sub`: 0 = the event's block; 1 = its second block (emitters that draw more than three uniforms: dslit, aperture)
    static void block(uint64_t seed, uint64_t id, uint32_t event, uint32_t out[4], uint32_t sub = 0u) {
        uint32_t c[4] = {event, (uint32_t)id, (uint32_t)(id >> 32), sub};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        for (int r = 0; r < 10; ++r) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        std::memcpy(out, c, sizeof(uint32_t) * 4);
    }
};
// word -> uniform conversions, done in IEEE binary32 exactly as the engine does, then widened
inline double u01(uint32_t w) { return (double)((float)(w >> 8) * 5.9604644775390625e-08f); }  // [0,1)
inline double u01_open0(uint32_t w) {                                                             // (0,1]
    float f = (float)w;
    f = f + 1.0f;
    return (double)(f * 2.3283064365386963e-10f);
}
// xoshiro256** (what gfortran's random_number uses); sequential stream for the "reference-like" mode
struct Xoshiro {
    uint64_t s[4];
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    void seed(uint64_t sd) {  // splitmix64 expansion
        for (int i = 0; i < 4; ++i) {
            sd += 0x9E3779B97F4A7C15ull;
            uint64_t z = sd;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
// One RNG front-end for the restated routines.  In PHILOX mode draws are addressed by slot inside the
// current event block (slot 3 of an emit/interact block is the tau of the following tauint2); in STREAM
// mode every draw is the next value of a sequential generator, exactly like ran2() in the reference.
struct Rng {
    int mode = 0;  // 0 philox-event, 1 xoshiro stream
    uint64_t seed = 0, id = 0;
    uint32_t event = 0;
    uint32_t w[4] = {0, 0, 0, 0};
    Xoshiro xo;
    void start_packet(uint64_t sd, uint64_t pid) {
        seed = sd; id = pid; event = 0;
    }
    void begin_event() {
        if (mode == 0) {
            Philox::block(seed, id, event, w);
            ++event;
        }
    }
    double draw(int slot) { return mode == 0 ? u01(w[slot]) : xo.uni(); }
    // uniforms 4 and 5 of an emission (dslit / aperture): words 0, 1 of the event's second Philox block
    void draw_extra(double out[2]) {
        if (mode == 0) {
            uint32_t x[4];
            Philox::block(seed, id, event - 1u, x, 1u);
            out[0] = u01(x[0]); out[1] = u01(x[1]);
        } else { out[0] = xo.uni(); out[1] = xo.uni(); }
    }
    double draw_tau() { return mode == 0 ? u01_open0(w[3]) : xo.uni(); }
};

// ---------------------------------------------------------------- packet (src/photon.f90:24-56)
struct Packet {
    V3 pos;
    double nxp, nyp, nzp;
    double sint, cost, sinp, cosp, phi;
    double phase = 0;
    int xcell, ycell, zcell;
    bool tflag;
    int layer;
    long cnts;
    int bounces;
    double weight;
    int step;
};
struct Tally {
    float *jmean = nullptr, *absorb = nullptr, *emission = nullptr;
    double* det = nullptr;
    int mode = TALLY_ABSORB;
    bool parallel = false;
};
struct ThreadCounters {
    double nscatt = 0, cnts = 0, bounces = 0, retries = 0, lost = 0, det_hits = 0;
};

inline void atomic_add_f(float* a, float v, bool par) {
    if (par) {
#pragma omp atomic
        *a += v;
    } else
        *a += v;
}
inline void atomic_add_d(double* a, double v, bool par) {
    if (par) {
#pragma omp atomic
        *a += v;
    } else
        *a += v;
}

// ---------------------------------------------------------------- detectors
// intersectPlane / intersectCircle, src/geometryMod.f90:217-270
bool intersect_plane(V3 n, V3 p0, V3 l0, V3 l, double& t) {
    double denom = dot(n, l);
    if (denom > 1e-6) {
        V3 p0l0 = p0 - l0;
        t = dot(p0l0, n);
        t = t / denom;
        if (t > -1e-6) return true;
    }
    return false;
}
bool intersect_circle(V3 n, V3 p0, double radius, V3 l0, V3 l, double& t, double& d2) {
    t = 0.0;
    if (intersect_plane(n, p0, l0, l, t)) {
        V3 p = l0 + l * t;
        V3 v = p - p0;
        d2 = std::sqrt(dot(v, v));
        if (d2 <= radius) return true;
    }
    return false;
}
struct Hit {  // hit_t, src/detectors/detector_base.f90:9-22
    V3 pos, dir;
    double pointSep, value1D, weight;
};
// check_hit_* , src/detectors/detectors.f90:147-164, 212-244, 323-393, 447-469
bool check_hit(const Detector& d, Hit& h) {
    double t = 0;
    switch (d.kind) {
        case DET_CIRCLE: {
            bool hit = intersect_circle(d.dir, d.pos, d.radius, h.pos, h.dir, t, h.value1D);
            if (hit && (t <= 0.0 || t > h.pointSep)) hit = false;
            return hit;
        }
        case DET_ANNULUS: {
            bool res = false;
            bool hit_r1 = intersect_circle(d.dir, d.pos, d.r1, h.pos, h.dir, t, h.value1D);
            bool hit_r2 = intersect_circle(d.dir, d.pos, d.r2, h.pos, h.dir, t, h.value1D);
            if (!hit_r1 && hit_r2) res = !(t <= 0.0 || t > h.pointSep);
            h.value1D = h.value1D - d.r1;  // Q8: happens even on a miss
            return res;
        }
        case DET_FIBRE: {
            bool hit = intersect_circle(d.dir, d.pos + d.dir * d.frontOff, d.f1Ap, h.pos, h.dir, t, h.value1D);
            if (hit && (t <= 0.0 || t > h.pointSep)) hit = false;
            if (!hit) return false;
            double costt = dot(d.dir, h.dir);
            if (costt > 1.0) costt = 1.0;
            double sintt = std::sqrt(1.0 - costt * costt);
            double gradient = sintt / costt;
            double radius = h.value1D;
            gradient = -radius / d.f1 + gradient;      // thin lens 1
            radius = radius + gradient * d.frontToPin;  // to the pinhole
            if (radius > d.pinAp) return false;
            radius = radius + gradient * d.pinToBack;  // to lens 2
            if (radius > d.f2Ap) return false;
            gradient = -radius / d.f2 + gradient;
            radius = radius + gradient * d.backOff;  // to the fibre face
            double angle = std::fabs(std::atan(gradient)) * 360.0 / TWOPI;
            if (angle > d.acceptAngle || radius > d.coreDiameter / 2.0) hit = false;
            h.value1D = std::fabs(radius);
            return hit;
        }
        case DET_CAMERA: {
            double tt = dot(d.pos - h.pos, d.n) / dot(h.dir, d.n);
            if (tt >= 0.0) {
                V3 v = (h.pos + tt * h.dir) - d.pos;
                double proj1 = dot(v, d.e1) / d.width, proj2 = dot(v, d.e2) / d.height;
                if (proj1 < d.width && proj1 > 0.0 && proj2 < d.height && proj2 > 0.0) return true;
            }
            return false;
        }
    }
    return false;
}
// Fortran nint: round half away from zero
inline long f_nint(double v) { return (long)std::llround(v); }
// record_hit_1D_sub / record_hit_2D_sub, src/detectors/detector_base.f90:137-163, 206-235.
// returns the 1-based flat bin or 0 on miss
int64_t record_hit_bin(const Detector& d, Hit& h) {
    if (!check_hit(d, h)) return 0;
    if (d.kind == DET_CAMERA) {
        double x = h.pos.z + d.pos.x;  // sic: the reference bins the segment START point
        double y = h.pos.y + d.pos.y;
        long idx = std::min((long)(x / d.bin_wid_x) + 1, (long)d.nbinsX);  // int() truncates toward zero
        long idy = std::min((long)(y / d.bin_wid_y) + 1, (long)d.nbinsY);
        if (idx < 1) idx = d.nbinsX;
        if (idy < 1) idy = d.nbinsY;
        return idx + (idy - 1) * (int64_t)d.nbinsX;
    }
    long idx = std::min(f_nint(h.value1D / d.bin_wid) + 1, (long)d.nbins);
    return idx;
}
void record_segment(const Scene& s, Tally& T, ThreadCounters& C, V3 start, V3 dir, double sep, int layer, double weight) {
    for (const Detector& d : s.dets) {
        Hit h{start, dir, sep, (double)layer, weight};  // hit_t(startPos, dir, pointSep, packet%layer, packet%weight)
        int64_t b = record_hit_bin(d, h);
        if (b > 0) {
            if (b > d.count()) b = d.count();  // (Fortran would write out of bounds for a negative annulus value; guard)
            if (b < 1) b = 1;
            double w = (d.kind == DET_CAMERA) ? 1.0 : h.weight;
            if (T.det) atomic_add_d(&T.det[d.offset + b - 1], w, T.parallel);
            C.det_hits += 1;
        }
    }
}

// ---------------------------------------------------------------- grid walk (src/inttau2.f90:367-614)
// update_voxels :587-614 (pos already shifted to the corner origin)
void update_voxels(const Grid& g, V3 pos, int& ci, int& cj, int& ck) {
    ci = (int)std::floor(g.nxg * pos.x / (2.0 * g.xmax)) + 1;
    cj = (int)std::floor(g.nyg * pos.y / (2.0 * g.ymax)) + 1;
    ck = (int)std::floor(g.nzg * pos.z / (2.0 * g.zmax)) + 1;
    if (ci > g.nxg || ci < 1) ci = -1;
    if (cj > g.nyg || cj < 1) cj = -1;
    if (ck > g.nzg || ck < 1) ck = -1;
}
// wall_dist :467-521 ; returns <0 where the reference would `error stop`
double wall_dist(const Grid& g, int ci, int cj, int ck, V3 pos, V3 dir, bool ldir[3]) {
    double dx = -999, dy = -999, dz = -999;
    if (dir.x > 0) dx = (g.xface[ci] - pos.x) / dir.x;  // xface(celli+1), 1-based -> [ci]
    else if (dir.x < 0) dx = (g.xface[ci - 1] - pos.x) / dir.x;
    else dx = 100000.0;
    if (dir.y > 0) dy = (g.yface[cj] - pos.y) / dir.y;
    else if (dir.y < 0) dy = (g.yface[cj - 1] - pos.y) / dir.y;
    else dy = 100000.0;
    if (dir.z > 0) dz = (g.zface[ck] - pos.z) / dir.z;
    else if (dir.z < 0) dz = (g.zface[ck - 1] - pos.z) / dir.z;
    else dz = 100000.0;
    double res = std::min(dx, std::min(dy, dz));
    ldir[0] = (res == dx); ldir[1] = (res == dy); ldir[2] = (res == dz);
    return res;
}
// update_pos :524-584
void update_pos(const Grid& g, V3& pos, int& ci, int& cj, int& ck, double dcell, bool wall, V3 dir, const bool ldir[3],
                double delta) {
    if (wall) {
        if (ldir[0]) {
            if (dir.x > 0) pos.x = g.xface[ci] + delta;
            else if (dir.x < 0) pos.x = g.xface[ci - 1] - delta;
            pos.y += dir.y * dcell;
            pos.z += dir.z * dcell;
        } else if (ldir[1]) {
            if (dir.y > 0) pos.y = g.yface[cj] + delta;
            else if (dir.y < 0) pos.y = g.yface[cj - 1] - delta;
            pos.x += dir.x * dcell;
            pos.z += dir.z * dcell;
        } else if (ldir[2]) {
            if (dir.z > 0) pos.z = g.zface[ck] + delta;
            else if (dir.z < 0) pos.z = g.zface[ck - 1] - delta;
            pos.x += dir.x * dcell;
            pos.y += dir.y * dcell;
        }
        update_voxels(g, pos, ci, cj, ck);
    } else {
        pos.x += dir.x * dcell;
        pos.y += dir.y * dcell;
        pos.z += dir.z * dcell;
    }
}
// update_grids :367-465.  `pos` is INOUT in the reference but every caller passes a temporary (oldpos).
void update_grids(const Scene& s, Tally& T, ThreadCounters& C, V3 pos, V3 dir, double d_sdf, Packet& pk) {
    const Grid& g = s.grid;
    V3 old_pos{pos.x + g.xmax, pos.y + g.ymax, pos.z + g.zmax};
    int ci, cj, ck;
    update_voxels(g, old_pos, ci, cj, ck);
    pk.xcell = ci; pk.ycell = cj; pk.zcell = ck;
    if (T.mode & TALLY_PATHLENGTH) {  // #ifdef pathlength :408-445
        const double delta = 1e-8;
        double d = 0.0;
        if (ci == -1 || cj == -1 || ck == -1) {
            pk.tflag = true;
            return;
        }
        for (;;) {
            bool ldir[3] = {false, false, false};
            double dcell = wall_dist(g, ci, cj, ck, old_pos, dir, ldir);
            if (dcell < 0.0) {  // reference: error stop 1 (:510-516)
                C.lost += 1;
                pk.tflag = true;
                break;
            }
            size_t vox = (size_t)(ci - 1) + (size_t)g.nxg * ((size_t)(cj - 1) + (size_t)g.nyg * (size_t)(ck - 1));
            if (d + dcell > d_sdf) {
                dcell = d_sdf - d;
                d = d_sdf;
                pk.phase += dcell;
                if (T.jmean) atomic_add_f(&T.jmean[vox], (float)((double)(float)dcell * pk.weight), T.parallel);
                update_pos(g, old_pos, ci, cj, ck, dcell, false, dir, ldir, delta);
                break;
            } else {
                d += dcell;
                pk.phase += dcell;
                if (T.jmean) atomic_add_f(&T.jmean[vox], (float)((double)(float)dcell * pk.weight), T.parallel);
                update_pos(g, old_pos, ci, cj, ck, dcell, true, dir, ldir, delta);
            }
            if (ci == -1 || cj == -1 || ck == -1) {
                pk.tflag = true;
                break;
            }
        }
        pk.xcell = ci; pk.ycell = cj; pk.zcell = ck;
    } else {  // default build :446-463
        old_pos.x += dir.x * d_sdf;
        old_pos.y += dir.y * d_sdf;
        old_pos.z += dir.z * d_sdf;
        update_voxels(g, old_pos, ci, cj, ck);
        if (ci == -1 || cj == -1 || ck == -1) pk.tflag = true;
        pk.xcell = ci; pk.ycell = cj; pk.zcell = ck;
    }
}

// ---------------------------------------------------------------- Fresnel (src/surfaces.f90)
double fresnel(V3 I, V3 N, double n1, double n2) {  // :86-127
    double costt = std::fabs(dot(I, N));
    if (costt > 1.0) costt = 1.0;
    double sintt = std::sqrt(1.0 - costt * costt);
    double sint2 = n1 / n2 * sintt;
    if (sint2 > 1.0) return 1.0;  // total internal reflection
    if (costt == 1.0) return 0.0;  // Q10: exactly normal incidence -> transmitted
    double cost2 = std::sqrt(1.0 - sint2 * sint2);
    double a = (n1 * costt - n2 * cost2) / (n1 * costt + n2 * cost2);
    double b = (n1 * cost2 - n2 * costt) / (n1 * cost2 + n2 * costt);
    return 0.5 * (a * a + b * b);
}
void reflect(V3& I, V3 N) { I = I - 2.0 * dot(N, I) * N; }  // :42-55
void refract(V3& I, V3 N, double eta) {                      // :57-84
    V3 Nt = N;
    double c1 = dot(Nt, I);
    if (c1 < 0.0) c1 = -c1;
    else Nt = (-1.0) * N;
    double c2 = std::sqrt(1.0 - eta * eta * (1.0 - c1 * c1));
    I = eta * I + (eta * c1 - c2) * Nt;
}
void reflect_refract(V3& I, V3 N, double n1, double n2, double xi, bool& rflag, double& Ri) {  // :14-40
    rflag = false;
    Ri = fresnel(I, N, n1, n2);
    if (xi <= Ri) {
        reflect(I, N);
        rflag = true;
    } else
        refract(I, N, n1 / n2);
}

// ---------------------------------------------------------------- scatter (src/photon.f90:1045-1103)
void scatter(Packet& p, double hgg, double xi_cost, double xi_phi) {
    if (hgg == 0.0) {
        p.cost = 2.0 * xi_cost - 1.0;
    } else {
        double temp = (1.0 - hgg * hgg) / (1.0 - hgg + 2.0 * hgg * xi_cost);
        p.cost = (1.0 + hgg * hgg - temp * temp) / (2.0 * hgg);
    }
    p.sint = std::sqrt(1.0 - p.cost * p.cost);
    p.phi = TWOPI * xi_phi;
    p.cosp = std::cos(p.phi);
    p.sinp = std::sin(p.phi);
    double uxx, uyy, uzz;
    if (p.nzp > 1.0 - 1e-12) {
        uxx = p.sint * p.cosp; uyy = p.sint * p.sinp; uzz = p.cost;
    } else if (p.nzp < -1.0 + 1e-12) {
        uxx = p.sint * p.cosp; uyy = p.sint * p.sinp; uzz = -p.cost;
    } else {
        double temp = std::sqrt(1.0 - p.nzp * p.nzp);
        uxx = p.sint * ((p.nxp * p.nzp * p.cosp - p.nyp * p.sinp) / temp) + p.nxp * p.cost;
        uyy = p.sint * ((p.nyp * p.nzp * p.cosp + p.nxp * p.sinp) / temp) + p.nyp * p.cost;
        uzz = -1.0 * p.sint * p.cosp * temp + p.nzp * p.cost;
    }
    double temp = std::sqrt(uxx * uxx + uyy * uyy + uzz * uzz);
    int guard = 0;
    while (std::fabs(temp - 1.0) > 1e-12 && guard++ < 64) {  // :1091-1097 (guard: the reference loop is unbounded)
        uxx /= temp; uyy /= temp; uzz /= temp;
        temp = std::sqrt(uxx * uxx + uyy * uyy + uzz * uzz);
    }
    p.nxp = uxx; p.nyp = uyy; p.nzp = uzz;
}

// ---------------------------------------------------------------- emitters (src/photon.f90:214-1043)
inline void nudge_face(double& c, double cmax) {  // the 7.9e-7 inset, photon.f90:614-628 etc.
    if (c == -cmax) c += 7.9e-7;
    else if (c == cmax) c -= 7.9e-7;
}
void set_dir_angles(Packet& p) {
    p.phi = std::atan2(p.nyp, p.nxp);
    p.cosp = std::cos(p.phi);
    p.sinp = std::sin(p.phi);
    p.cost = p.nzp;
    p.sint = std::sqrt(1.0 - p.cost * p.cost);
}
// shared tail of focus/annulus: rotate (0,0,-1)->rotation, translate, clip onto the grid box
// (photon.f90:440-562 and :927-1042; the two differ only in the retry cap: counter>4 vs counter>3)
void aim_and_clip(const Scene& s, Packet& pk, V3 dir, int counter_cap) {
    const Grid& g = s.grid;
    const double* sp = s.src.p;
    V3 a = magnitude(V3{0, 0, -1});
    V3 b = magnitude(V3{sp[SP_ROT], sp[SP_ROT + 1], sp[SP_ROT + 2]});
    V3 startPos{-sp[SP_POS], -sp[SP_POS + 1], -sp[SP_POS + 2]};
    M4 t;
    bool anti = veq(vabs(a), vabs(b)) && !veq(a, b);
    if (veq(a, b)) t = m_identity();
    else if (veq(vabs(a), vabs(b))) {
        t = m_identity();
        t.at(3, 3) = -1.0;
    } else
        t = m_rotation_align(a, b);
    // NB `dir .dot. t` is the AFFINE row-vector product (translation row is zero here)
    dir = vec_dot_mat(dir, t);
    dir = magnitude(dir);
    if (anti) t.at(3, 3) = 1.0;
    t = m_matmul(t, m_invert(m_translate(startPos)));
    pk.pos = vec_dot_mat(pk.pos, t);
    pk.nxp = dir.x; pk.nyp = dir.y; pk.nzp = dir.z;
    set_dir_angles(pk);
    bool inX = false, inY = false, inZ = false, triedX = false, triedY = false, triedZ = false;
    int counter = 0;
    while (!inX || !inY || !inZ) {
        double step;
        if (pk.pos.x <= -g.xmax) {
            step = (-g.xmax - pk.pos.x + 9e-7) / pk.nxp;
            pk.pos = pk.pos + dir * step; triedX = true;
        } else if (pk.pos.x >= g.xmax) {
            step = (g.xmax - pk.pos.x - 9e-7) / pk.nxp;
            pk.pos = pk.pos + dir * step; triedX = true;
        } else inX = true;
        if (pk.pos.y <= -g.ymax) {
            step = (-g.ymax - pk.pos.y + 9e-7) / pk.nyp;
            pk.pos = pk.pos + dir * step; triedY = true;
        } else if (pk.pos.y >= g.ymax) {
            step = (g.ymax - pk.pos.y - 9e-7) / pk.nyp;
            pk.pos = pk.pos + dir * step; triedY = true;
        } else inY = true;
        if (pk.pos.z <= -g.zmax) {
            step = (-g.zmax - pk.pos.z + 9e-7) / pk.nzp;
            pk.pos = pk.pos + dir * step; triedZ = true;
        } else if (pk.pos.z >= g.zmax) {
            step = (g.zmax - pk.pos.z - 9e-7) / pk.nzp;
            pk.pos = pk.pos + dir * step; triedZ = true;
        } else inZ = true;
        if ((triedX && triedY && triedZ) || counter > counter_cap) break;
        ++counter;
    }
}
// Emit one packet.  xi[0..2] are the uniforms of the event block (slot meaning per source in DESIGN.md §RNG), xi[3..4] two more
// for the emitters that draw five (dslit) or four (aperture).
// Returns false when a rejection step inside the emitter (gaussian annulus, rang) wants a fresh block.
bool emit(const Scene& s, Packet& pk, const double xi[5]) {
    const Grid& g = s.grid;
    const double* sp = s.src.p;
    V3 opos{sp[SP_POS], sp[SP_POS + 1], sp[SP_POS + 2]};
    V3 odir{sp[SP_DIR], sp[SP_DIR + 1], sp[SP_DIR + 2]};
    pk.phase = 0;
    switch (s.src.kind) {
        case SRC_POINT: {  // :311-359
            pk.pos = opos;
            pk.phi = xi[0] * TWOPI;
            pk.cosp = std::cos(pk.phi);
            pk.sinp = std::sin(pk.phi);
            pk.cost = 2.0 * xi[1] - 1.0;
            pk.sint = std::sqrt(1.0 - pk.cost * pk.cost);
            pk.nxp = pk.sint * pk.cosp;
            pk.nyp = pk.sint * pk.sinp;
            pk.nzp = pk.cost;
            break;
        }
        case SRC_PENCIL: {  // :652-710
            pk.pos = opos;
            nudge_face(pk.pos.x, g.xmax); nudge_face(pk.pos.y, g.ymax); nudge_face(pk.pos.z, g.zmax);
            pk.nxp = odir.x; pk.nyp = odir.y; pk.nzp = odir.z;
            set_dir_angles(pk);
            break;
        }
        case SRC_UNIFORM: {  // :566-649
            pk.nxp = odir.x; pk.nyp = odir.y; pk.nzp = odir.z;
            set_dir_angles(pk);
            double rx = xi[0], ry = xi[1];
            pk.pos.x = sp[SP_P1] + rx * sp[SP_P2] + ry * sp[SP_P3];
            pk.pos.y = sp[SP_P1 + 1] + rx * sp[SP_P2 + 1] + ry * sp[SP_P3 + 1];
            pk.pos.z = sp[SP_P1 + 2] + rx * sp[SP_P2 + 2] + ry * sp[SP_P3 + 2];
            nudge_face(pk.pos.x, g.xmax); nudge_face(pk.pos.y, g.ymax); nudge_face(pk.pos.z, g.zmax);
            break;
        }
        case SRC_CIRCULAR: {  // :214-308
            pk.nxp = odir.x; pk.nyp = odir.y; pk.nzp = odir.z;
            double r = sp[SP_RADIUS] * std::sqrt(xi[0]);
            double theta = xi[1] * TWOPI;
            V3 a = magnitude(V3{1, 0, 0});
            V3 b = magnitude(V3{pk.nxp, pk.nyp, pk.nzp});
            if (veq(vabs(a), vabs(b))) {
                a = magnitude(V3{0, 0, 1});
                pk.pos = V3{r * std::cos(theta), r * std::sin(theta), 0.0};
            } else
                pk.pos = V3{0.0, r * std::cos(theta), r * std::sin(theta)};
            M4 t = m_rotation_align(a, b);
            t = m_matmul(t, m_invert(m_translate(opos)));
            pk.pos = vec_dot_mat(pk.pos, t);
            pk.pos = V3{-pk.pos.x, -pk.pos.y, -pk.pos.z};
            nudge_face(pk.pos.x, g.xmax); nudge_face(pk.pos.y, g.ymax); nudge_face(pk.pos.z, g.zmax);
            set_dir_angles(pk);
            break;
        }
        case SRC_FOCUS: {  // :361-563
            double focal = sp[SP_FOCAL], beam = sp[SP_BEAM];
            if (s.src.subtype == 1) {  // square: ranu(-b,b) twice
                pk.pos = V3{-beam + xi[0] * (beam - (-beam)), -beam + xi[1] * (beam - (-beam)), 0.0};
            } else if (s.src.subtype == 2) {  // circle
                double radius = beam * std::sqrt(xi[0]), phi = TWOPI * xi[1];
                pk.pos = V3{radius * std::cos(phi), radius * std::sin(phi), 0.0};
            } else {  // gaussian (1/e radius)
                double radius = beam * std::sqrt(-std::log(1 - xi[0])), phi = TWOPI * xi[1];
                pk.pos = V3{radius * std::cos(phi), radius * std::sin(phi), 0.0};
            }
            V3 targ{0, 0, -focal};
            double dist = length(pk.pos - targ);
            V3 dir = (-1.0) * (pk.pos - targ) / dist;
            dir = dir * fsign(1.0, focal);
            dir = magnitude(dir);
            aim_and_clip(s, pk, dir, 4);
            break;
        }
        case SRC_DSLIT:       // :712-780   (draw order: slit pick, x1, y1, x2, y2)
        case SRC_APERTURE: {  // :782-848   (draw order: x1, y1, x2, y2)
            // phase experiments with hard-coded geometry in units of the wavelength (sp[SP_RADIUS] carries it: the slot is unused
            // by these two kinds); ranu(a, b) = a + xi (b - a), also with b < a (random_mod.f90:126-135)
            const double lam = sp[SP_RADIUS];
            auto ranu = [](double a, double b, double u) { return a + u * (b - a); };
            double x1, y1, z1, x2, y2, z2;
            if (s.src.kind == SRC_DSLIT) {
                const double a = 60.0 * lam, b = 20.0 * lam;
                if (xi[0] > 0.5) x1 = ranu(a / 2.0, a / 2.0 + b, xi[1]);
                else x1 = ranu(-a / 2.0, -a / 2.0 - b, xi[1]);
                y1 = ranu(-b * 0.5, b * 0.5, xi[2]);
                z2 = 5.0 - (1.e-5 * (2.0 * (5.0 / 400.0)));
                x2 = ranu(-5.0, 5.0, xi[3]);
                y2 = ranu(-5.0, 5.0, xi[4]);
                z1 = (10000.0 * lam) - 5.0;
            } else {
                const double apwid = 200e-6, b = apwid / 2.0, F = 4.95;
                x1 = ranu(-b, b, xi[0]);
                y1 = ranu(-b, b, xi[1]);
                z1 = (1.0 / ((((F / apwid) * (F / apwid)) / 2.0) * lam)) - 0.5;
                x2 = ranu(-0.5, 0.5, xi[2]);
                y2 = ranu(-0.5, 0.5, xi[3]);
                z2 = 0.5 - (1.e-5 * (2.0 * 0.5 / 400.0));
            }
            pk.pos = V3{x2, y2, z2};
            pk.phase = std::sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1));
            pk.nxp = (x2 - x1) / pk.phase;
            pk.nyp = (y2 - y1) / pk.phase;
            pk.nzp = -std::fabs(z2 - z1) / pk.phase;
            pk.cost = pk.nzp;
            pk.sint = std::sqrt(1.0 - pk.cost * pk.cost);
            pk.phi = std::atan2(pk.nyp, pk.nxp);
            pk.cosp = std::cos(pk.phi);
            pk.sinp = std::sin(pk.phi);
            break;
        }
        case SRC_ANNULUS: {  // :850-1043
            double focal = sp[SP_FOCAL], rlo = sp[SP_RLO], rhi = sp[SP_RHI], sigma = sp[SP_SIGMA];
            double radius, mid = (rhi + rlo) / 2.0;
            if (s.src.subtype == 1) radius = std::sqrt(rlo * rlo + (rhi * rhi - rlo * rlo) * xi[0]);
            else if (s.src.subtype == 2) radius = rlo + (rhi - rlo) * xi[0];
            else {  // rang(radius, tmp, mid, sigma), src/random_mod.f90:99-124 — polar Box-Muller with rejection
                double x = -1.0 + xi[0] * 2.0, y = -1.0 + xi[1] * 2.0;
                double sq = y * y + x * x;
                if (sq >= 1.0 || sq == 0.0) return false;  // rejected: caller supplies a fresh block
                radius = mid + sigma * (x * std::sqrt(-2.0 * std::log(sq) / sq));
            }
            double phi = TWOPI * xi[2], cosp = std::cos(phi), sinp = std::sin(phi);
            pk.pos = V3{radius * cosp, radius * sinp, 0.0};
            V3 targ{0, 0, -focal};
            V3 ring{mid * cosp, mid * sinp, 0.0};
            double dist = length(ring - targ);
            V3 dir = (-1.0) * (ring - targ) / dist;
            dir = dir * fsign(1.0, focal);
            dir = magnitude(dir);
            aim_and_clip(s, pk, dir, 3);
            break;
        }
    }
    pk.tflag = false;
    pk.cnts = 0;
    pk.bounces = 0;
    pk.layer = 1;
    pk.weight = 1.0;
    int cell[3];
    g.get_voxel(pk.pos, cell);
    pk.xcell = cell[0]; pk.ycell = cell[1]; pk.zcell = cell[2];
    return true;
}

// ---------------------------------------------------------------- tauint2 (src/inttau2.f90:15-364)
struct Work {
    std::vector<double> ds, dsNew;
};
// returns false if the reference would have hit `error stop` (:276) — packet is marked lost
bool tauint2(const Scene& s, Tally& T, ThreadCounters& C, Packet& pk, Rng& rng, Work& W) {
    const int N = (int)s.top.size();
    double* ds = W.ds.data();
    double* dsNew = W.dsNew.data();
    V3 pos = pk.pos, oldpos = pos, startPos = pos;
    V3 dir{pk.nxp, pk.nyp, pk.nzp};
    const double eps = 1e-8;                      // :56
    // (> 0 strictly, like the engine: the Philox uniform of the tau draw can round to exactly 1; ran2() of the reference cannot)
    const double tau = std::max(-std::log(rng.draw_tau()), 1e-30);  // :58
    double taurun = 0.0, d_sdf, t_sdf;
    auto kappa = [&](int layer) { return s.opt[layer - 1].kappa; };
    auto eval_all = [&](V3 p, double* out) {
        for (int i = 0; i < N; ++i) out[i] = eval_top(s, i, p);
        pk.cnts += N;
    };
    auto min_abs = [&](const double* a) {
        double m = std::fabs(a[0]);
        for (int i = 1; i < N; ++i) m = std::min(m, std::fabs(a[i]));
        return m;
    };
    auto min_val = [&](const double* a) {
        double m = a[0];
        for (int i = 1; i < N; ++i) m = std::min(m, a[i]);
        return m;
    };
    auto detect = [&]() {  // :126-131 etc.
        double sep = std::sqrt((pos.x - startPos.x) * (pos.x - startPos.x) + (pos.y - startPos.y) * (pos.y - startPos.y) +
                               (pos.z - startPos.z) * (pos.z - startPos.z));
        record_segment(s, T, C, startPos, dir, sep, pk.layer, pk.weight);
        startPos = pos;
    };

    while (taurun <= tau) {  // :61
        eval_all(pos, ds);
        d_sdf = min_abs(ds);
        if (d_sdf < eps) {  // on a boundary :73-146
            d_sdf = min_abs(ds) + 2.0 * eps;
            V3 smallStepPos = pos + d_sdf * dir;
            eval_all(smallStepPos, ds);
            int smallStepLayer = maxloc_neg(ds, N);
            if (smallStepLayer == pk.layer) {  // forward :86-102
                oldpos = pos;
                t_sdf = d_sdf * kappa(pk.layer);
                if (taurun + t_sdf < tau) {
                    pos = pos + d_sdf * dir;
                    taurun += t_sdf;
                    update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                } else {
                    d_sdf = (tau - taurun) / kappa(pk.layer);
                    taurun += t_sdf;  // Q1: pos not advanced
                    if (!s.bugcompat) pos = pos + d_sdf * dir;
                    update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                }
            } else {  // backward :104-121
                oldpos = pos;
                t_sdf = d_sdf * kappa(pk.layer);
                if (taurun + t_sdf < tau) {
                    pos = pos - d_sdf * dir;
                    taurun += t_sdf;
                    update_grids(s, T, C, oldpos, dir, d_sdf, pk);  // Q2: deposits along +dir
                } else {
                    d_sdf = (tau - taurun) / kappa(pk.layer);
                    pos = pos - d_sdf * dir;
                    if (!s.bugcompat) taurun = tau;  // Q3: taurun not advanced in the reference
                    update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                }
            }
            detect();
            eval_all(pos, ds);  // :134-140
            d_sdf = min_abs(ds);
            if (min_val(ds) > 0.0) pk.tflag = true;
        }
        if (taurun >= tau || pk.tflag) break;  // :149-152

        while (d_sdf >= eps) {  // :155-192
            t_sdf = d_sdf * kappa(pk.layer);
            if (taurun + t_sdf < tau) {
                taurun += t_sdf;
                oldpos = pos;
                update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                pos = pos + d_sdf * dir;
            } else {
                d_sdf = (tau - taurun) / kappa(pk.layer);
                taurun = tau;
                oldpos = pos;
                pos = pos + d_sdf * dir;
                update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                break;
            }
            eval_all(pos, ds);
            d_sdf = min_abs(ds);
            if (min_val(ds) > 0.0) {
                pk.tflag = true;
                break;
            }
        }
        detect();                              // :196-201
        if (taurun >= tau || pk.tflag) break;  // :204-207

        // boundary crossing :213-337
        d_sdf = min_abs(ds) + 2.0 * eps;
        V3 smallStepPos = pos + d_sdf * dir;
        eval_all(smallStepPos, dsNew);
        int new_layer = maxloc_neg(dsNew, N);
        double glancing = min_abs(dsNew);
        const int old_layer = pk.layer;
        while (new_layer == old_layer && glancing < eps) {  // :225-235
            d_sdf += eps;
            smallStepPos = pos + d_sdf * dir;
            eval_all(smallStepPos, dsNew);
            new_layer = maxloc_neg(dsNew, N);
            glancing = min_abs(dsNew);
        }
        if (new_layer == 0) {  // :237-241
            pk.tflag = true;
            break;
        }
        double n1 = s.opt[pk.layer - 1].n, n2 = s.opt[new_layer - 1].n;
        if (n1 != n2) {  // :248-317
            int layer = -1;
            if (dsNew[new_layer - 1] < 0.0 && ds[new_layer - 1] >= 0.0) layer = new_layer;
            else if (dsNew[old_layer - 1] >= 0.0 && ds[old_layer - 1] < 0.0) layer = old_layer;
            else if (dsNew[new_layer - 1] < 0.0 && dsNew[old_layer - 1] < 0.0) layer = new_layer;
            else if (ds[old_layer - 1] >= 0.0 && dsNew[old_layer - 1] >= 0.0) layer = old_layer;
            else {
                C.lost += 1;
                pk.tflag = true;
                return false;
            }
            V3 Nrm = calc_normal(s, layer - 1, pos);
            bool rflag = false;
            double Ri;
            rng.begin_event();  // one Fresnel event = one block, slot 0
            reflect_refract(dir, Nrm, n1, n2, rng.draw(0), rflag, Ri);
            if (!rflag) {  // transmitted :284-303
                pk.layer = new_layer;
                oldpos = pos;
                update_grids(s, T, C, oldpos, dir, d_sdf, pk);
                t_sdf = d_sdf * kappa(pk.layer);
                taurun += t_sdf;
                pos = smallStepPos;  // Q4: probe used the pre-refraction direction
                detect();
            } else {  // reflected :304-317
                oldpos = pos;
                startPos = pos;
                pk.bounces += 1;
                if (pk.bounces > 1000) {  // Q5: return with no write-back in the reference; here: lost
                    C.lost += 1;
                    pk.tflag = true;
                    return false;
                }
            }
        } else {  // :318-337
            pk.layer = new_layer;
            oldpos = pos;
            update_grids(s, T, C, oldpos, dir, d_sdf, pk);
            t_sdf = d_sdf * kappa(pk.layer);
            taurun += t_sdf;
            pos = smallStepPos;
            detect();
        }
        if (pk.tflag) break;
    }
    pk.pos = pos;  // :341-351
    pk.nxp = dir.x; pk.nyp = dir.y; pk.nzp = dir.z;
    set_dir_angles(pk);
    const Grid& g = s.grid;  // :354-362
    if (std::fabs(pk.pos.x) > g.xmax) pk.tflag = true;
    if (std::fabs(pk.pos.y) > g.ymax) pk.tflag = true;
    if (std::fabs(pk.pos.z) > g.zmax) pk.tflag = true;
    return true;
}

inline size_t voxel_index(const Grid& g, int ci, int cj, int ck) {
    return (size_t)(ci - 1) + (size_t)g.nxg * ((size_t)(cj - 1) + (size_t)g.nyg * (size_t)(ck - 1));
}

struct PacketOut {
    int fate = 0, nscatt = 0, events = 0;
    V3 pos{0, 0, 0};
};

// noBiasPropagation / survivalBiasPropagation (src/kernelsMod.f90:1901-1976, 1979-2067)
void propagate(const Scene& s, Tally& T, ThreadCounters& C, Rng& rng, Work& W, bool survival, double threshold,
               double chance, PacketOut* out) {
    const Grid& g = s.grid;
    const int N = (int)s.top.size();
    Packet pk{};
    auto emit_once = [&]() {
        for (;;) {
            rng.begin_event();
            double xi[5] = {rng.draw(0), rng.draw(1), rng.draw(2), 0.0, 0.0};
            if (s.src.kind >= SRC_DSLIT) rng.draw_extra(xi + 3);
            if (emit(s, pk, xi)) return;
            C.retries += 1;
        }
    };
    emit_once();
    int guard = 0;
    while (pk.xcell < 1 || pk.xcell > g.nxg || pk.ycell < 1 || pk.ycell > g.nyg || pk.zcell < 1 || pk.zcell > g.nzg) {
        C.retries += 1;  // Q6: :1939-1943
        if (++guard > 100000) {
            C.lost += 1;
            if (out) out->fate = 3;
            return;
        }
        emit_once();
    }
    if ((T.mode & TALLY_EMISSION) && T.emission)  // recordEmissionLocation :2184-2200
        atomic_add_f(&T.emission[voxel_index(g, pk.xcell, pk.ycell, pk.zcell)], 1.0f, T.parallel);
    pk.step = 0;
    for (int i = 0; i < N; ++i) W.ds[i] = eval_top(s, i, pk.pos);
    pk.layer = maxloc_neg(W.ds.data(), N, s.launch_mask_le);  // :1949-1952 (test_kernel: <=, :2136)
    int fate = 1;
    int nsc = 0;
    if (pk.layer == 0) {  // the reference would index array(0) (unguarded); treat as lost
        C.lost += 1;
        if (out) { out->fate = 3; out->pos = pk.pos; out->events = (int)rng.event; }
        return;
    }
    bool ok = tauint2(s, T, C, pk, rng, W);
    while (ok && !pk.tflag) {
        rng.begin_event();  // interaction event: slot0 = ran, slot1/2 = scatter, slot3 = next tau
        double ran = rng.draw(0);
        const Optics& o = s.opt[pk.layer - 1];
        if (!survival) {
            if (ran < o.albedo) {
                scatter(pk, o.hgg, rng.draw(1), rng.draw(2));
                C.nscatt += 1; ++nsc;
                pk.step += 1;
            } else {
                pk.tflag = true;
                if ((T.mode & TALLY_ABSORB) && T.absorb)  // recordWeight(packet, 1.0) :2202-2220
                    atomic_add_f(&T.absorb[voxel_index(g, pk.xcell, pk.ycell, pk.zcell)], (float)1.0, T.parallel);
                fate = 0;
                break;
            }
        } else {
            double wabs = pk.weight * (1.0 - o.albedo);
            pk.weight -= wabs;
            if ((T.mode & TALLY_ABSORB) && T.absorb) {
                float* a = &T.absorb[voxel_index(g, pk.xcell, pk.ycell, pk.zcell)];
                // absorb(c) = absorb(c) + weightAbsorbed : real32 + real64 evaluated in real64, stored real32
                if (T.parallel) {
#pragma omp atomic
                    *a += (float)wabs;
                } else
                    *a = (float)((double)*a + wabs);
            }
            if (pk.weight < threshold) {
                if (ran < chance) pk.weight = pk.weight / chance;
                else {
                    pk.tflag = true;
                    fate = 2;
                    break;
                }
            }
            scatter(pk, o.hgg, rng.draw(1), rng.draw(2));
            C.nscatt += 1; ++nsc;
            pk.step += 1;
        }
        ok = tauint2(s, T, C, pk, rng, W);
    }
    if (!ok) fate = 3;
    C.cnts += (double)pk.cnts;
    C.bounces += pk.bounces;
    if (out) {
        out->fate = fate; out->nscatt = nsc; out->pos = pk.pos; out->events = (int)rng.event;
    }
}

void finish_detector(Detector& d) {
    if (d.kind == DET_CAMERA) {
        d.e1 = d.p2 - d.pos;
        d.e2 = d.p3 - d.pos;
        d.width = length(d.e1);
        d.height = length(d.e2);
        d.n = magnitude(cross(d.e2, d.e1));
    }
}

}  // namespace

// ============================================================================================
// C interface (ctypes)
// ============================================================================================
extern "C" {

struct orc_counters {
    double nscatt, sdf_evals, bounces, launched, emit_retries, lost, sweeps, det_hits;
};

void* orc_scene_create(int n_nodes, const int32_t* kind, const int32_t* first_child, const int32_t* n_child,
                       const double* xform, const double* params, int n_top, const int32_t* top_node, const double* mus,
                       const double* mua, const double* hgg, const double* n_ref) {
    Scene* s = new Scene();
    s->nodes.resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        Node& nd = s->nodes[i];
        nd.kind = kind[i];
        nd.first_child = first_child ? first_child[i] : 0;
        nd.n_child = n_child ? n_child[i] : 0;
        std::memcpy(nd.xf.m, xform + 16 * (size_t)i, sizeof(double) * 16);
        std::memcpy(nd.p, params + NODE_P * (size_t)i, sizeof(double) * NODE_P);
    }
    s->top.assign(top_node, top_node + n_top);
    s->opt.resize(n_top);
    for (int t = 0; t < n_top; ++t) s->opt[t].init(mus[t], mua[t], hgg[t], n_ref[t]);
    s->grid.init(200, 200, 200, 1.0, 1.0, 1.0);
    return s;
}
void orc_scene_free(void* h) { delete (Scene*)h; }
void orc_set_grid(void* h, int nx, int ny, int nz, double xm, double ym, double zm) { ((Scene*)h)->grid.init(nx, ny, nz, xm, ym, zm); }
void orc_set_source(void* h, int kind, int subtype, const double* p) {
    Scene* s = (Scene*)h;
    s->src.kind = kind;
    s->src.subtype = subtype;
    std::memcpy(s->src.p, p, sizeof(double) * SP_N);
}
void orc_set_optprops(void* h, int top_index, double mus, double mua, double hgg, double n) {
    ((Scene*)h)->opt[top_index - 1].init(mus, mua, hgg, n);
}
void orc_set_flags(void* h, int bugcompat, int launch_mask_le) {
    ((Scene*)h)->bugcompat = bugcompat != 0;
    ((Scene*)h)->launch_mask_le = launch_mask_le != 0;
}
int64_t orc_set_detectors(void* h, int n, const int32_t* kind, const double* p, const int32_t* nbins) {
    Scene* s = (Scene*)h;
    s->dets.clear();
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        const double* q = p + (size_t)DET_P * i;
        Detector d;
        d.kind = kind[i];
        d.pos = V3{q[0], q[1], q[2]};
        d.dir = V3{q[3], q[4], q[5]};
        int nb = nbins[i];
        switch (d.kind) {
            case DET_CIRCLE:
                d.radius = q[6];
                d.nbins = nb + 1;
                d.bin_wid = nb == 0 ? 1.0 : d.radius / (double)nb;
                break;
            case DET_ANNULUS:
                d.r1 = q[6]; d.r2 = q[7];
                d.nbins = nb + 1;
                d.bin_wid = nb == 0 ? 1.0 : (d.r2 - d.r1) / (double)nb;
                break;
            case DET_FIBRE:
                d.f1 = q[6]; d.f2 = q[7]; d.f1Ap = q[8]; d.f2Ap = q[9]; d.frontOff = q[10]; d.backOff = q[11];
                d.frontToPin = q[12]; d.pinToBack = q[13]; d.pinAp = q[14]; d.acceptAngle = q[15]; d.coreDiameter = q[16];
                d.nbins = nb + 1;
                d.bin_wid = nb == 0 ? 1.0 : d.coreDiameter / 2 / (double)nb;
                break;
            case DET_CAMERA:
                d.p2 = V3{q[3], q[4], q[5]};
                d.p3 = V3{q[6], q[7], q[8]};
                d.nbinsX = nb + 1; d.nbinsY = nb + 1;
                if (nb == 0) d.bin_wid_x = d.bin_wid_y = 1.0;
                else {
                    d.bin_wid_x = q[9] / (double)d.nbinsX;
                    d.bin_wid_y = q[9] / (double)d.nbinsY;
                }
                finish_detector(d);
                break;
        }
        d.offset = off;
        off += d.count();
        s->dets.push_back(d);
    }
    s->det_total = off;
    return off;
}

void orc_sdf_eval(void* h, int top_index, int64_t n, const double* pos, double* out) {
    Scene* s = (Scene*)h;
    const int N = (int)s->top.size();
    for (int64_t i = 0; i < n; ++i) {
        V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        if (top_index > 0) out[i] = eval_top(*s, top_index - 1, p);
        else
            for (int t = 0; t < N; ++t) out[i * N + t] = eval_top(*s, t, p);
    }
}
void orc_sdf_normal(void* h, int top_index, int64_t n, const double* pos, double* out) {
    Scene* s = (Scene*)h;
    for (int64_t i = 0; i < n; ++i) {
        V3 nn = calc_normal(*s, top_index - 1, V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
        out[3 * i] = nn.x; out[3 * i + 1] = nn.y; out[3 * i + 2] = nn.z;
    }
}
int orc_locate_layer(void* h, const double* pos, int le) {
    Scene* s = (Scene*)h;
    const int N = (int)s->top.size();
    std::vector<double> ds(N);
    for (int t = 0; t < N; ++t) ds[t] = eval_top(*s, t, V3{pos[0], pos[1], pos[2]});
    return maxloc_neg(ds.data(), N, le != 0);
}
void orc_fresnel(int64_t n, const double* dir, const double* nrm, const double* n1, const double* n2, const double* xi,
                 double* dir_out, double* R, int32_t* rflag) {
    for (int64_t i = 0; i < n; ++i) {
        V3 I{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}, N{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
        bool rf;
        double Ri;
        reflect_refract(I, N, n1[i], n2[i], xi[i], rf, Ri);
        dir_out[3 * i] = I.x; dir_out[3 * i + 1] = I.y; dir_out[3 * i + 2] = I.z;
        if (R) R[i] = Ri;
        if (rflag) rflag[i] = rf ? 1 : 0;
    }
}
void orc_scatter(int64_t n, const double* dir, const double* hgg, const double* xi2, double* dir_out) {
    for (int64_t i = 0; i < n; ++i) {
        Packet p{};
        p.nxp = dir[3 * i]; p.nyp = dir[3 * i + 1]; p.nzp = dir[3 * i + 2];
        scatter(p, hgg[i], xi2[2 * i], xi2[2 * i + 1]);
        dir_out[3 * i] = p.nxp; dir_out[3 * i + 1] = p.nyp; dir_out[3 * i + 2] = p.nzp;
    }
}
// xi4: 4 uniforms per packet (slot 3 unused here). ok[i]=0 when the emitter asked for a fresh block.
void orc_emit(void* h, int64_t n, const double* xi4, double* pos, double* dir, int32_t* cell, int32_t* ok) {
    Scene* s = (Scene*)h;
    for (int64_t i = 0; i < n; ++i) {
        Packet pk{};
        // probe convention for the five-uniform emitters: uniforms 4 and 5 are xi4[3] and 1 - xi4[3] (smcrt_probe_emit does the same)
        const double xi[5] = {xi4[4 * i], xi4[4 * i + 1], xi4[4 * i + 2], xi4[4 * i + 3], 1.0 - xi4[4 * i + 3]};
        bool r = emit(*s, pk, xi);
        if (ok) ok[i] = r ? 1 : 0;
        pos[3 * i] = pk.pos.x; pos[3 * i + 1] = pk.pos.y; pos[3 * i + 2] = pk.pos.z;
        dir[3 * i] = pk.nxp; dir[3 * i + 1] = pk.nyp; dir[3 * i + 2] = pk.nzp;
        if (cell) { cell[3 * i] = pk.xcell; cell[3 * i + 1] = pk.ycell; cell[3 * i + 2] = pk.zcell; }
    }
}
void orc_detector(void* h, int det_index, int64_t n, const double* start, const double* dir, const double* len,
                  int32_t* hit, int32_t* bin) {
    Scene* s = (Scene*)h;
    const Detector& d = s->dets[det_index - 1];
    for (int64_t i = 0; i < n; ++i) {
        Hit hh{V3{start[3 * i], start[3 * i + 1], start[3 * i + 2]}, V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}, len[i], 0.0, 1.0};
        int64_t b = record_hit_bin(d, hh);
        hit[i] = b > 0;
        bin[i] = (int32_t)b;
    }
}
void orc_get_voxel(void* h, const double* pos, int32_t* cell) {
    int c[3];
    ((Scene*)h)->grid.get_voxel(V3{pos[0], pos[1], pos[2]}, c);
    cell[0] = c[0]; cell[1] = c[1]; cell[2] = c[2];
}
void orc_philox(uint64_t seed, uint64_t id, uint32_t event, uint32_t out[4]) { Philox::block(seed, id, event, out); }
void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        Philox::round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    std::memcpy(out, c, 16);
}
void orc_uniforms(const uint32_t w[4], double out[5]) {
    for (int i = 0; i < 4; ++i) out[i] = u01(w[i]);
    out[4] = u01_open0(w[3]);
}

// matrix / vector helpers for the transcribed KATs (test/SDF/test_SDF.f90:303-676, test/matrix, test/vector)
void orc_rotate_x(double a, double* out) { M4 m = m_rotate_x(a); std::memcpy(out, m.m, 128); }
void orc_rotate_y(double a, double* out) { M4 m = m_rotate_y(a); std::memcpy(out, m.m, 128); }
void orc_rotate_z(double a, double* out) { M4 m = m_rotate_z(a); std::memcpy(out, m.m, 128); }
void orc_rotmat(const double* axis, double a, double* out) { M4 m = m_rotmat(V3{axis[0], axis[1], axis[2]}, a); std::memcpy(out, m.m, 128); }
void orc_rotation_align(const double* a, const double* b, double* out) {
    M4 m = m_rotation_align(V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]});
    std::memcpy(out, m.m, 128);
}
void orc_translate(const double* o, double* out) { M4 m = m_translate(V3{o[0], o[1], o[2]}); std::memcpy(out, m.m, 128); }
void orc_identity(double* out) { M4 m = m_identity(); std::memcpy(out, m.m, 128); }
void orc_skew(const double* a, double* out) { M4 m = m_skew(V3{a[0], a[1], a[2]}); std::memcpy(out, m.m, 128); }
void orc_invert(const double* in, double* out) {
    M4 a; std::memcpy(a.m, in, 128);
    M4 m = m_invert(a); std::memcpy(out, m.m, 128);
}
void orc_matmul(const double* a, const double* b, double* out) {
    M4 x, y; std::memcpy(x.m, a, 128); std::memcpy(y.m, b, 128);
    M4 m = m_matmul(x, y); std::memcpy(out, m.m, 128);
}
void orc_vec_dot_mat(const double* v, const double* m, double* out) {
    M4 x; std::memcpy(x.m, m, 128);
    V3 r = vec_dot_mat(V3{v[0], v[1], v[2]}, x);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_mono(double mus, double mua, double hgg, double n, double* out7) {
    Optics o; o.init(mus, mua, hgg, n);
    out7[0] = o.mus; out7[1] = o.mua; out7[2] = o.hgg; out7[3] = o.g2; out7[4] = o.n; out7[5] = o.kappa; out7[6] = o.albedo;
}

// test_kernel (src/kernelsMod.f90:2069-2182): the serial variant the reference's end-to-end tests run.  Differences to
// noBiasPropagation that matter: launch layer uses mask=(distances<=0) (:2136), no start-voxel rejection loop, no absorb
// deposit, positions summed after scatter orders 1..4 and (end_early) the packet is dropped after the 5th scatter.
// moments: 24 doubles = 10*<r> for orders 1..4 (x,y,z) then 100*<r^2> for orders 1..4, exactly what positions.dat holds.
double orc_test_kernel(void* h, int64_t nphotons, uint64_t seed, int end_early, int nthreads, double* moments) {
    Scene* s = (Scene*)h;
    const int N = (int)s->top.size();
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double sum[24] = {0};
    double nscatt = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double loc[24] = {0};
        double lsc = 0;
        Tally T;
        T.mode = 0;
        T.parallel = nthreads > 1;
        ThreadCounters C;
        Rng rng;
        rng.mode = 0;
        Work W;
        W.ds.resize(N); W.dsNew.resize(N);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t j = 0; j < nphotons; ++j) {
            rng.start_packet(seed, (uint64_t)j);
            Packet pk{};
            for (;;) {
                rng.begin_event();
                double xi[5] = {rng.draw(0), rng.draw(1), rng.draw(2), 0.0, 0.0};
                if (s->src.kind >= SRC_DSLIT) rng.draw_extra(xi + 3);
                if (emit(*s, pk, xi)) break;
            }
            pk.step = 0;
            for (int i = 0; i < N; ++i) W.ds[i] = eval_top(*s, i, pk.pos);
            pk.layer = maxloc_neg(W.ds.data(), N, true);
            if (pk.layer == 0) continue;
            bool ok = tauint2(*s, T, C, pk, rng, W);
            while (ok && !pk.tflag) {
                rng.begin_event();
                const Optics& o = s->opt[pk.layer - 1];
                if (rng.draw(0) < o.albedo) {
                    scatter(pk, o.hgg, rng.draw(1), rng.draw(2));
                    lsc += 1;
                    pk.step += 1;
                    if (pk.step >= 1 && pk.step <= 4) {
                        const int k = pk.step - 1;
                        loc[3 * k] += pk.pos.x; loc[3 * k + 1] += pk.pos.y; loc[3 * k + 2] += pk.pos.z;
                        loc[12 + 3 * k] += pk.pos.x * pk.pos.x; loc[12 + 3 * k + 1] += pk.pos.y * pk.pos.y;
                        loc[12 + 3 * k + 2] += pk.pos.z * pk.pos.z;
                    } else if (end_early)
                        pk.tflag = true;
                } else {
                    pk.tflag = true;
                    break;
                }
                ok = tauint2(*s, T, C, pk, rng, W);
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (int i = 0; i < 24; ++i) sum[i] += loc[i];
            nscatt += lsc;
        }
    }
    for (int i = 0; i < 12; ++i) moments[i] = 10.0 * sum[i] / (double)nphotons;
    for (int i = 12; i < 24; ++i) moments[i] = 100.0 * sum[i] / (double)nphotons;
    return nscatt / (double)nphotons;
}

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// The photon loop of run_MCRT (src/kernelsMod.f90:1861-1888).  rng_mode 0: Philox event blocks keyed by
// (seed, global packet id) — the engine's stream; 1: per-thread sequential xoshiro256** seeded seed+thread id,
// which is how the reference seeds (kernelsMod.f90:1850-1852).  Tallies accumulate into the caller's arrays.
// Returns wall seconds spent in the loop (what the reference times, :1832,1893-1897).
double orc_run(void* h, int64_t nphotons, uint64_t seed, int64_t id_offset, int tally_mode, int survival_bias,
               double threshold, double chance, int nthreads, int rng_mode, float* jmean, float* absorb,
               float* emission, double* det_bins, orc_counters* counters, int32_t* fate, int32_t* nscatt_pp,
               double* final_pos, int32_t* n_events) {
    Scene* s = (Scene*)h;
    if (threshold <= 0) threshold = REF_THRESHOLD;
    if (chance <= 0) chance = REF_CHANCE;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    ThreadCounters total;
    const int N = (int)s->top.size();
    double t0 = 0, t1 = 0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        Tally T;
        T.jmean = jmean; T.absorb = absorb; T.emission = emission; T.det = det_bins;
        T.mode = tally_mode;
        T.parallel = nthreads > 1;
        ThreadCounters C;
        Rng rng;
        rng.mode = rng_mode;
        if (rng_mode == 1) {
            rng.xo.seed(seed + (uint64_t)tid);
            for (int i = 0; i < 100; ++i) rng.xo.uni();  // init_rng(fwd=.true.), random_mod.f90:70-76
        }
        Work W;
        W.ds.resize(N); W.dsNew.resize(N);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t j = 0; j < nphotons; ++j) {
            rng.start_packet(seed, (uint64_t)(id_offset + j));
            PacketOut po;
            propagate(*s, T, C, rng, W, survival_bias != 0, threshold, chance, &po);
            if (fate) fate[j] = po.fate;
            if (nscatt_pp) nscatt_pp[j] = po.nscatt;
            if (n_events) n_events[j] = po.events;
            if (final_pos) { final_pos[3 * j] = po.pos.x; final_pos[3 * j + 1] = po.pos.y; final_pos[3 * j + 2] = po.pos.z; }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.nscatt += C.nscatt; total.cnts += C.cnts; total.bounces += C.bounces; total.retries += C.retries;
            total.lost += C.lost; total.det_hits += C.det_hits;
        }
    }
#ifdef _OPENMP
    t1 = omp_get_wtime();
#endif
    if (counters) {
        counters->nscatt += total.nscatt;
        counters->sdf_evals += total.cnts;
        counters->bounces += total.bounces;
        counters->launched += (double)nphotons;
        counters->emit_retries += total.retries;
        counters->lost += total.lost;
        counters->sweeps += N > 0 ? total.cnts / N : 0;
        counters->det_hits += total.det_hits;
    }
    return t1 - t0;
}

}  // extern "C"
