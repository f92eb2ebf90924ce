// device_scene.cuh — flattened struct-of-records scene as the sm_100a kernel sees it, and the SDF evaluators.
//
// Layout in HBM / shared memory: one blob per context, copied once per kernel into shared
// memory by every CTA (a few KB; all lanes of a warp read the same primitive in the same iteration, so the
// reads are shared-memory broadcasts):
//     DevPrim  prims[n_prims]      96 B each: kind, transform class, 3x4 affine (row-vector convention of
//                                  src/vector_class.f90:292-304 folded into columns), 8 parameters
//     DevTop   tops[n_top]         32 B each: how to evaluate top-level SDF i + its optical properties
//     DevInstr prog[n_instr]       16 B each: postfix program for `model`/modifier trees
//     DevDet   dets[n_det]         96 B each
// The evaluators are templated on the scalar so the same formulas run in FP32 (transport) and FP64
// (4-tap surface normal with the reference's h = 1e-6, src/sdfs/sdf_base.f90:166-190).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace smcrt_dev {

// XF_AFFINE: rigid (orthonormal 3x3 block + translation); XF_NONRIGID: anything else (scale, shear).  The reference accepts any
// 4x4 transform and steps by |d| (src/inttau2.f90:155-192); the closed-form ray bounds and hit geometry of this engine assume that
// the local ray direction is a unit vector, which only a rigid transform guarantees: non-rigid primitives keep plain sphere tracing.
enum : int { XF_IDENTITY = 0, XF_TRANSLATE = 1, XF_AFFINE = 2, XF_NONRIGID = 3 };

template <typename T>
struct PrimT {
    int32_t kind;
    int32_t xf;
    T m[12];  // p'.x = m[0] x + m[1] y + m[2] z + m[3]; p'.y = m[4..7]; p'.z = m[8..11]
    T p[8];
    int32_t pad_[2];
};
using DevPrim = PrimT<float>;    // 96 B
using DevPrimD = PrimT<double>;  // 176 B
static_assert(sizeof(DevPrim) == 96, "DevPrim must be 96 bytes");

struct DevTop {
    int32_t mode;   // 0: single primitive `first`; 1: program prog[first .. first+count)
    int32_t first;
    int32_t count;
    float kappa, albedo, hgg, n, mua;
};
static_assert(sizeof(DevTop) == 32, "DevTop must be 32 bytes");

// postfix program (compiled from the node tree in smcrt_set_scene)
enum : int {
    I_PRIM = 1,        // a = prim index            push d = prim(P)
    I_CSG = 2,         // a = model kind, f[0] = k   d2=pop, d1=pop, push op(d1,d2,k)
    I_PUSH_REV = 3,    // f = o, cx, cy              (cz in the next word via `g`)  push P' (revolution)
    I_PUSH_ELONG = 4,  // f = sx, sy, sz             push P' = max(|P|-s,0)
    I_PUSH_TWIST = 5,  // f[0] = k
    I_PUSH_BEND = 6,   // f[0] = k
    I_POP_P = 7,
    I_EXTRUDE = 8,     // f[0] = h                   d = pop; push extrude(d, P.z)
    I_ONION = 9,       // f[0] = thickness
    I_ELONG_ADD = 10   // f = sx, sy, sz             d += min(max(q),0), q = |P|-s
};
struct DevInstr {
    int32_t op;
    int32_t a;
    float f[3];
    float g;
    int32_t pad_[2];
};
static_assert(sizeof(DevInstr) == 32, "DevInstr must be 32 bytes");
struct DevInstrD {
    int32_t op, a;
    double f[3];
    double g;
};

struct DevDet {
    int32_t kind;
    int32_t nbins;     // stored count (user + 1); cameras: per axis
    int32_t offset;    // into the concatenated bin array
    int32_t pad_;
    float pos[3];      // circle/annulus/fibre: centre (fibre: already pos + dir*frontOffset); camera: p1
    float dir[3];      // plane normal (camera: n = normalised e2 x e1)
    float q[14];       // kind-specific, see engine.cu:pack_detector
};
static_assert(sizeof(DevDet) == 96, "DevDet must be 96 bytes");

// -------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x);
template <> __device__ __forceinline__ float t_abs<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double t_abs<double>(double x) { return fabs(x); }
template <typename T> __device__ __forceinline__ T t_min(T a, T b);
template <> __device__ __forceinline__ float t_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double t_min<double>(double a, double b) { return fmin(a, b); }
template <typename T> __device__ __forceinline__ T t_max(T a, T b);
template <> __device__ __forceinline__ float t_max<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double t_max<double>(double a, double b) { return fmax(a, b); }
template <typename T> __device__ __forceinline__ void t_sincos(T x, T* s, T* c);
template <> __device__ __forceinline__ void t_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void t_sincos<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <typename T> __device__ __forceinline__ T t_clamp01(T v) { return t_min(t_max(v, T(0)), T(1)); }

// One primitive, reference formulas of src/sdfs/sdfs.f90:494-735
template <typename T>
__device__ __forceinline__ T eval_prim(const PrimT<T>& P, T x, T y, T z) {
    T px, py, pz;
    if (P.xf == XF_IDENTITY) {
        px = x; py = y; pz = z;
    } else if (P.xf == XF_TRANSLATE) {
        px = x + P.m[3]; py = y + P.m[7]; pz = z + P.m[11];
    } else {
        px = P.m[0] * x + P.m[1] * y + P.m[2] * z + P.m[3];
        py = P.m[4] * x + P.m[5] * y + P.m[6] * z + P.m[7];
        pz = P.m[8] * x + P.m[9] * y + P.m[10] * z + P.m[11];
    }
    const T* q = P.p;
    switch (P.kind) {
        case 1:  // sphere :494-508
            return t_sqrt(px * px + py * py + pz * pz) - q[0];
        case 2: {  // box :510-525
            T dx = t_abs(px) - q[0], dy = t_abs(py) - q[1], dz = t_abs(pz) - q[2];
            T ox = t_max(dx, T(0)), oy = t_max(dy, T(0)), oz = t_max(dz, T(0));
            return t_sqrt(ox * ox + oy * oy + oz * oz) + t_min(t_max(dx, t_max(dy, dz)), T(0));
        }
        case 3: {  // torus :527-542 (axis = y)
            T tx = t_sqrt(px * px + pz * pz) - q[0];
            return t_sqrt(tx * tx + py * py) - q[1];
        }
        case 4: {  // capped cylinder a->b :544-581
            T bax = q[3] - q[0], bay = q[4] - q[1], baz = q[5] - q[2];
            T pax = px - q[0], pay = py - q[1], paz = pz - q[2];
            T baba = bax * bax + bay * bay + baz * baz;
            T paba = pax * bax + pay * bay + paz * baz;
            T vx = pax * baba - bax * paba, vy = pay * baba - bay * paba, vz = paz * baba - baz * paba;
            T xx = t_sqrt(vx * vx + vy * vy + vz * vz) - q[6] * baba;
            T yy = t_abs(paba - baba * T(0.5)) - baba * T(0.5);
            T x2 = xx * xx, y2 = (yy * yy) * baba, d;
            if (t_max(xx, yy) < T(0)) d = -t_min(x2, y2);
            else d = (xx > T(0) ? x2 : T(0)) + (yy > T(0) ? y2 : T(0));
            T r = t_sqrt(t_abs(d)) / baba;
            return d >= T(0) ? r : -r;
        }
        case 5: {  // triprism :583-597
            T ax = t_abs(px), az = t_abs(pz);
            return t_max(az - q[1], t_max(ax * T(0.866025) + py * T(0.5), -py) - q[0] * T(0.5));
        }
        case 6:    // segment :599-626 (radius 0.1)
        case 7: {  // capsule :628-648
            T pax = px - q[0], pay = py - q[1], paz = pz - q[2];
            T bax = q[3] - q[0], bay = q[4] - q[1], baz = q[5] - q[2];
            T h = t_clamp01((pax * bax + pay * bay + paz * baz) / (bax * bax + bay * bay + baz * baz));
            T ex = pax - bax * h, ey = pay - bay * h, ez = paz - baz * h;
            return t_sqrt(ex * ex + ey * ey + ez * ez) - (P.kind == 6 ? T(0.1) : q[6]);
        }
        case 8: {  // capped cone :650-686
            T ra = q[6], rb = q[7], rba = rb - ra;
            T bax = q[3] - q[0], bay = q[4] - q[1], baz = q[5] - q[2];
            T pax = px - q[0], pay = py - q[1], paz = pz - q[2];
            T baba = bax * bax + bay * bay + baz * baz;
            T paba = (pax * bax + pay * bay + paz * baz) / baba;
            // reference: x = sqrt(pa.pa - baba*paba^2); identical in exact arithmetic to |pa - ba*paba| (the component of
            // pa perpendicular to the axis) but without the catastrophic cancellation in FP32
            T ex = pax - bax * paba, ey = pay - bay * paba, ez = paz - baz * paba;
            T xx = t_sqrt(ex * ex + ey * ey + ez * ez);
            T cax = t_max(T(0), xx - (paba < T(0.5) ? ra : rb));
            T cay = t_abs(paba - T(0.5)) - T(0.5);
            T k = rba * rba + baba;
            T f = t_clamp01((rba * (xx - ra) + paba * baba) / k);
            T cbx = xx - ra - f * rba, cby = paba - f;
            T s = (cbx < T(0) && cay < T(0)) ? T(-1) : T(1);
            return s * t_sqrt(t_min(cax * cax + baba * cay * cay, cbx * cbx + baba * cby * cby));
        }
        case 9: {  // egg :688-718
            T r1 = q[0], r2 = q[1], h = q[2];
            T ax = t_abs(px);
            T r = r1 - r2, hin = h + r;
            T l = (hin * hin - r * r) / (T(2) * r);
            if (py <= T(0)) return t_sqrt(ax * ax + py * py + pz * pz) - r1;
            if ((py - hin) * l > ax * hin) {
                T yy = py - hin;
                return t_sqrt(ax * ax + yy * yy + pz * pz) - ((r1 + l) - t_sqrt(hin * hin + l * l));
            }
            T xx = ax + l;
            return t_sqrt(xx * xx + py * py + pz * pz) - (r1 + l);
        }
        case 10:  // plane :720-735
            return px * q[0] + py * q[1] + pz * q[2];
    }
    return T(1e30);
}

#define SMCRT_BIG 3.0e38f
// Capsule / segment (kinds 7 / 6) with its along-ray hit distance.  A capsule is convex (segment (+) ball), so the ray meets it in
// ONE interval [Tin, Tout] = [min in_i, max out_i] over the non-empty intervals of its three convex parts: the finite cylinder
// around a-b and the two end spheres.  Roots use the cancellation-free forms of the sphere code (rejection-form discriminant,
// {w, C/w} root pair).  The bound is shortened by a few ulp and NOT flagged exact: the next sweep confirms the landing.
__device__ __forceinline__ float2 eval_capsule_ray(const PrimT<float>& P, float px, float py, float pz, float vx, float vy, float vz, float need) {
    const float* q = P.p;
    const float r = P.kind == 6 ? 0.1f : q[6];
    const float pax = px - q[0], pay = py - q[1], paz = pz - q[2];
    const float bax = q[3] - q[0], bay = q[4] - q[1], baz = q[5] - q[2];
    const float baba = bax * bax + bay * bay + baz * baz;
    const float inv = baba > 0.f ? 1.0f / baba : 0.f;  // a == b: a sphere (the reference divides by zero)
    const float baoa = pax * bax + pay * bay + paz * baz, bard = vx * bax + vy * bay + vz * baz;
    // signed distance (src/sdfs/sdfs.f90:628-648)
    const float h = fminf(fmaxf(baoa * inv, 0.f), 1.f);
    const float ex = pax - bax * h, ey = pay - bay * h, ez = paz - baz * h;
    const float d = sqrtf(ex * ex + ey * ey + ez * ez) - r;
    if (fabsf(d) >= need) return make_float2(d, fabsf(d));  // farther than the packet will travel: the plain distance is bound enough
    float tin = SMCRT_BIG, tout = -SMCRT_BIG;
    // end spheres
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float ox = k ? px - q[3] : pax, oy = k ? py - q[4] : pay, oz = k ? pz - q[5] : paz;
        const float b = ox * vx + oy * vy + oz * vz;
        const float rx = ox - b * vx, ry = oy - b * vy, rz = oz - b * vz;
        const float disc = r * r - (rx * rx + ry * ry + rz * rz);
        if (disc >= 0.f) {
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float c = (len - r) * (len + r);
            const float w = -(b + copysignf(sqrtf(disc), b));
            const float t1 = w, t2 = (w != 0.f) ? c / w : w;
            tin = fminf(tin, fminf(t1, t2));
            tout = fmaxf(tout, fmaxf(t1, t2));
        }
    }
    // finite cylinder: (infinite cylinder about the axis) intersected with the slab 0 <= y(t) = baoa + t bard <= baba
    if (baba > 0.f) {
        const float fo = baoa * inv, fv = bard * inv;
        const float opx = pax - bax * fo, opy = pay - bay * fo, opz = paz - baz * fo;   // perpendicular parts
        const float vpx = vx - bax * fv, vpy = vy - bay * fv, vpz = vz - baz * fv;
        const float A = vpx * vpx + vpy * vpy + vpz * vpz;
        const float B = opx * vpx + opy * vpy + opz * vpz;
        const float cx = opy * vpz - opz * vpy, cy = opz * vpx - opx * vpz, cz = opx * vpy - opy * vpx;
        const float disc = A * r * r - (cx * cx + cy * cy + cz * cz);   // A r^2 - |op x vp|^2  (Lagrange identity)
        const float lenop = sqrtf(opx * opx + opy * opy + opz * opz);
        float c1 = SMCRT_BIG, c2 = -SMCRT_BIG;
        if (A > 1e-12f) {
            if (disc >= 0.f) {
                const float C = (lenop - r) * (lenop + r);
                const float w = -(B + copysignf(sqrtf(disc), B));
                const float t1 = w / A, t2 = (w != 0.f) ? C / w : t1;
                c1 = fminf(t1, t2); c2 = fmaxf(t1, t2);
            }
        } else if (lenop <= r) { c1 = -SMCRT_BIG; c2 = SMCRT_BIG; }  // parallel to the axis, inside the tube
        float s1 = -SMCRT_BIG, s2 = SMCRT_BIG;
        if (fabsf(bard) > 1e-20f) {
            const float ta = -baoa / bard, tb = (baba - baoa) / bard;
            s1 = fminf(ta, tb); s2 = fmaxf(ta, tb);
        } else if (baoa < 0.f || baoa > baba) { s1 = SMCRT_BIG; s2 = -SMCRT_BIG; }
        const float lo = fmaxf(c1, s1), hi = fminf(c2, s2);
        if (lo <= hi) { tin = fminf(tin, lo); tout = fmaxf(tout, hi); }
    }
    float t = SMCRT_BIG;
    if (tout > 0.f && tin <= tout) t = tin > 0.f ? tin : tout;
    // a few ulp short of the root: the landing is confirmed (and finished) by the ordinary sphere-trace step
    if (t < SMCRT_BIG) t = t - (2e-6f * t + 4e-7f * (fabsf(px) + fabsf(py) + fabsf(pz)));
    return make_float2(d, fmaxf(fabsf(d), t));  // a miss leaves t = SMCRT_BIG: this body does not limit the step
}

// Distance AND directional step bound of one primitive (FP32 transport path).
// The reference advances by min_i |d_i| (sphere tracing, src/inttau2.f90:155-192): |d_i| is a lower bound of the distance
// to surface i in ANY direction.  Along the packet's own ray a tighter bound exists for the primitives with a closed-form
// ray intersection: the exact distance t_i >= |d_i| at which the ray meets surface i (+inf if it never does).  Stepping
// by min_i max(|d_i|, t_i) visits the same boundary points with the same optical depth (kappa is constant inside a
// layer and deposits are linear along a straight ray), in one step instead of O(log(1/eps)/(1-cos)).
// Kinds without a closed form keep b = |d| (plain sphere tracing).  *exact tells whether b is an exact hit distance.
// `need` = how far the packet can travel before its optical depth runs out (+ margin; SMCRT_BIG when unknown).  A capsule whose
// plain distance |d| is already beyond that cannot limit this move, so its (expensive) ray interval is not computed (bound = |d|):
// in a scattering medium most sweeps end in an interaction long before any wall.  Spheres and boxes are cheap enough that the
// test costs more than it saves (measured: -3..-7 % on every scene without capsules), so they ignore `need`.
//
// sphere (local point p, unit direction v): t^2 + 2 b t + c = 0 with c = |p|^2 - r^2 = d (|p| + r)  (no cancellation near the surface)
__device__ __forceinline__ float sphere_ray(float px, float py, float pz, float vx, float vy, float vz, float r, float& bound) {
    const float len = sqrtf(px * px + py * py + pz * pz);
    const float d = len - r;
    const float b = px * vx + py * vy + pz * vz;
    const float c = d * (len + r);
    // discriminant b^2 - c = r^2 - |p - b v|^2 (v unit): the rejection form does not subtract two O(|p|^2) numbers,
    // which matters for grazing rays (disc -> 0)
    const float rx = px - b * vx, ry = py - b * vy, rz = pz - b * vz;
    const float disc = r * r - (rx * rx + ry * ry + rz * rz);
    float t = SMCRT_BIG;
    if (disc >= 0.f) {
        const float w = -(b + copysignf(sqrtf(disc), b));  // numerically stable root pair {w, c/w}
        const float t1 = w, t2 = (w != 0.f) ? c / w : SMCRT_BIG;
        const float lo = fminf(t1, t2), hi = fmaxf(t1, t2);
        t = lo > 0.f ? lo : (hi > 0.f ? hi : SMCRT_BIG);
    }
    bound = fmaxf(fabsf(d), t);
    return d;
}
// box with half sizes (q0,q1,q2): slab method in the local frame
__device__ __forceinline__ float box_ray(float px, float py, float pz, float vx, float vy, float vz, float q0, float q1, float q2, float& bound) {
    const float dx = fabsf(px) - q0, dy = fabsf(py) - q1, dz = fabsf(pz) - q2;
    const float ox = fmaxf(dx, 0.f), oy = fmaxf(dy, 0.f), oz = fmaxf(dz, 0.f);
    const float d = sqrtf(ox * ox + oy * oy + oz * oz) + fminf(fmaxf(dx, fmaxf(dy, dz)), 0.f);
    const float ix = 1.0f / vx, iy = 1.0f / vy, iz = 1.0f / vz;  // +-inf for an axis-parallel ray: handled by min/max
    const float sx = copysignf(q0, vx), sy = copysignf(q1, vy), sz = copysignf(q2, vz);
    const float n1 = (-sx - px) * ix, f1 = (sx - px) * ix;
    const float n2 = (-sy - py) * iy, f2 = (sy - py) * iy;
    const float n3 = (-sz - pz) * iz, f3 = (sz - pz) * iz;
    const float tn = fmaxf(n1, fmaxf(n2, n3)), tf = fminf(f1, fminf(f2, f3));  // fmaxf/fminf drop a NaN (0*inf) operand
    float t = SMCRT_BIG;
    if (tf >= 0.f && tn <= tf) t = tn > 0.f ? tn : tf;
    bound = fmaxf(fabsf(d), t);
    return d;
}
__device__ __forceinline__ float eval_prim_ray(const PrimT<float>& P, float x, float y, float z, float ux, float uy, float uz, float need,
                                               float& bound, bool& exact) {
    float px, py, pz, vx = ux, vy = uy, vz = uz;
    if (P.xf == XF_IDENTITY) {
        px = x; py = y; pz = z;
    } else if (P.xf == XF_TRANSLATE) {
        px = x + P.m[3]; py = y + P.m[7]; pz = z + P.m[11];
    } else {
        px = P.m[0] * x + P.m[1] * y + P.m[2] * z + P.m[3];
        py = P.m[4] * x + P.m[5] * y + P.m[6] * z + P.m[7];
        pz = P.m[8] * x + P.m[9] * y + P.m[10] * z + P.m[11];
        vx = P.m[0] * ux + P.m[1] * uy + P.m[2] * uz;
        vy = P.m[4] * ux + P.m[5] * uy + P.m[6] * uz;
        vz = P.m[8] * ux + P.m[9] * uy + P.m[10] * uz;
    }
    const float* q = P.p;
    if (P.xf == XF_NONRIGID) {  // |v| != 1: local ray parameters are not world distances
        const float d = eval_prim<float>(P, x, y, z);
        bound = fabsf(d);
        exact = false;
        return d;
    }
    exact = true;
    if (P.kind == 1) return sphere_ray(px, py, pz, vx, vy, vz, q[0], bound);
    if (P.kind == 2) return box_ray(px, py, pz, vx, vy, vz, q[0], q[1], q[2], bound);
    if (P.kind == 10) {  // plane
        const float d = px * q[0] + py * q[1] + pz * q[2];
        const float dn = vx * q[0] + vy * q[1] + vz * q[2];
        const float t = -d / dn;
        bound = fmaxf(fabsf(d), (t > 0.f && t < SMCRT_BIG) ? t : SMCRT_BIG);
        return d;
    }
    exact = false;
    if (P.kind == 7 || P.kind == 6) {
        const float2 r = eval_capsule_ray(P, px, py, pz, vx, vy, vz, need);
        bound = r.y;
        return r.x;
    }
    const float d = eval_prim<float>(P, x, y, z);
    bound = fabsf(d);
    return d;
}

template <typename T>
__device__ __forceinline__ T csg_op(int kind, T d1, T d2, T k) {  // src/sdfs/sdfModifiers.f90:428-491
    switch (kind) {
        case 20: return t_min(d1, d2);
        case 21: {
            T h = t_max(k - t_abs(d1 - d2), T(0)) / k;
            return t_min(d1, d2) - h * h * h * k * T(1.0 / 6.0);
        }
        case 22: return t_max(-d1, d2);
        default: return t_max(d1, d2);
    }
}

// program interpreter for `model` / modifier trees (sdf_base.f90:146-161, sdfModifiers.f90:286-426).
// Stacks are tiny (depth checked at compile time on the host: <= 4 distances, <= 3 saved points).
template <typename T, typename PRIM, typename INSTR>
static __device__ __noinline__ T eval_program(const PRIM* prims, const INSTR* prog, int first, int count, T x, T y, T z) {
    T ds[4];
    T ps[3][3];
    int nd = 0, np = 0;
    for (int ip = first; ip < first + count; ++ip) {
        const INSTR I = prog[ip];
        switch (I.op) {
            case I_PRIM: ds[nd++] = eval_prim<T>(prims[I.a], x, y, z); break;
            case I_CSG: {
                T d2 = ds[--nd], d1 = ds[--nd];
                ds[nd++] = csg_op<T>(I.a, d1, d2, (T)I.f[0]);
                break;
            }
            case I_PUSH_REV: {
                ps[np][0] = x; ps[np][1] = y; ps[np][2] = z; ++np;
                T ix = x - (T)I.f[1], iy = y - (T)I.f[2], iz = z - (T)I.g;
                x = t_sqrt(ix * ix + iz * iz) - (T)I.f[0];
                y = iy;
                z = T(0);
                break;
            }
            case I_PUSH_ELONG: {
                ps[np][0] = x; ps[np][1] = y; ps[np][2] = z; ++np;
                x = t_max(t_abs(x) - (T)I.f[0], T(0));
                y = t_max(t_abs(y) - (T)I.f[1], T(0));
                z = t_max(t_abs(z) - (T)I.f[2], T(0));
                break;
            }
            case I_PUSH_TWIST: {
                ps[np][0] = x; ps[np][1] = y; ps[np][2] = z; ++np;
                T s, c;
                t_sincos<T>((T)I.f[0] * z, &s, &c);
                T nx = c * x - s * y, ny = s * x + c * y;
                x = nx; y = ny;
                break;
            }
            case I_PUSH_BEND: {
                ps[np][0] = x; ps[np][1] = y; ps[np][2] = z; ++np;
                T s, c;
                t_sincos<T>((T)I.f[0] * x, &s, &c);
                T nx = c * x - s * y, ny = s * x + c * y;
                x = nx; y = ny;
                break;
            }
            case I_POP_P: --np; x = ps[np][0]; y = ps[np][1]; z = ps[np][2]; break;
            case I_EXTRUDE: {
                T wx = ds[nd - 1], wy = t_abs(z) - (T)I.f[0];
                T ox = t_max(wx, T(0)), oy = t_max(wy, T(0));
                ds[nd - 1] = t_min(t_max(wx, wy), T(0)) + t_sqrt(ox * ox + oy * oy);
                break;
            }
            case I_ONION: ds[nd - 1] = t_abs(ds[nd - 1]) - (T)I.f[0]; break;
            case I_ELONG_ADD: {
                T qx = t_abs(x) - (T)I.f[0], qy = t_abs(y) - (T)I.f[1], qz = t_abs(z) - (T)I.f[2];
                ds[nd - 1] += t_min(t_max(qx, t_max(qy, qz)), T(0));
                break;
            }
        }
    }
    return ds[0];
}

// The sweep's compact view of a top-level SDF: 2 x float4 = {code, a, b, c}, {p0, p1, p2, -}.
//   HOT_SPHERE / HOT_BOX: a single sphere / box primitive with identity or translation transform, evaluated inline from these
//                         32 bytes: (a,b,c) = translation (local point = world point + t), p = radius / half sizes
//   HOT_PROGRAM:          compound `model`: a = first instruction, b = instruction count (int bits)
//   HOT_GENERAL:          any other single primitive: a = primitive index (int bits) -> eval_prim_ray_general
enum : int { HOT_GENERAL = 0, HOT_SPHERE = 1, HOT_BOX = 2, HOT_PROGRAM = 3 };
struct DevHot {
    int32_t code;
    union { float t[3]; int32_t idx[3]; };
    float p[3];
    float pad_;
};
static_assert(sizeof(DevHot) == 32, "DevHot must be 32 bytes");

// Every single primitive the inline sphere / box code of the sweep does not cover: planes, capsules, transformed primitives, kinds
// without a closed-form ray bound.  Out of line: ONE copy, and the sweep loop stays small (I-cache).
// -> (distance, step bound, exact flag)
static __device__ __noinline__ float3 eval_prim_ray_general(const DevPrim* prim, float x, float y, float z, float ux, float uy, float uz, float need) {
    float b;
    bool ex;
    const float d = eval_prim_ray(*prim, x, y, z, ux, uy, uz, need, b, ex);
    return make_float3(d, b, ex ? 1.f : 0.f);
}

}  // namespace smcrt_dev
