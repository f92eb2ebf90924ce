// kernels.cuh — the persistent photon-packet kernel for sm_100a and its device helpers.
//
// One thread = one packet at a time; a thread that loses its packet immediately launches the next one from a
// global work counter (persistent threads).  The reference's four nested data-dependent loops
// (src/kernelsMod.f90:1958, src/inttau2.f90:61,155,225) are flattened into ONE loop whose body is
//     [cold events: Fresnel | end-of-tauint2 | interaction | emit]  ->  [sweep: evaluate ALL SDFs at one point]
//     ->  [cheap state transition]
// so that every live lane of a warp executes exactly one sweep per iteration (the sweep is the dominant cost
// and its loop over primitives is warp-uniform), see DESIGN.md §4.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "device_scene.cuh"

namespace smcrt_dev {

// ------------------------------------------------------------------------------------------------ params
// What the launch sweeps of a fixed-ray (pencil) source lead to: first_flight_kernel computes it once per run
struct FirstFlight {
    int ok, layer, exact, pad_;
    float kap, d0, s, eps;
};
struct KParams {
    // scene blob (global) and its carve-up; copied to shared memory by every CTA
    const unsigned char* blob;
    int blob_bytes;
    int n_prims, n_top, n_instr, n_det;
    int has_capsule;  // host side: pick the NEED kernels (a bare capsule / segment in the scene, or clear cells in use)
    int simple_scene;  // host side: pick the SIMPLE kernels (translated spheres / boxes only)
    int has_camera;  // a camera detector counts SEGMENTS (detector_base.f90:222-229): segments are then never merged
    int off_tops, off_prog, off_dets, off_hot, off_detp;  // byte offsets inside the blob (prims at 0)
    const DevPrimD* primsD;            // FP64 copies for the surface normal
    const DevInstrD* progD;
    // culling grid: per coarse cell the candidate list of top-level SDFs; null = evaluate all
    const int* cull_start;     // [ncell + 1]
    const int* cull_items;     // top-level SDF indices (0-based), ascending inside a cell
    const float* cull_far;     // [ncell] lower bound of |d| of every SDF NOT in the cell's list, for any point of the cell
    const float* cull_clear;   // [ncell] lower bound of min_i |d_i| over the cell (0 when a surface may cross it); null = not used
    int cull_n[3];
    float cull_lo[3], cull_inv[3];  // cell = floor((x - lo) * inv)
    // voxel grid (src/grid.f90:14-25)
    int nxg, nyg, nzg;
    float gmax[3];     // half extents
    float vox[3];      // voxel edge 2*max/n
    float inv_vox[3];  // n/(2*max)
    float hvox[3];     // half a voxel edge, max/n (centre-origin cell faces of the DDA)
    // source (src/photon.f90), transforms precomputed on the host
    int src_kind, src_sub, src_alt;
    // batched point sources (smcrt_run_sources; the escape-function drivers, kernelsMod.f90:533-642,959-1071): packet id ->
    // source index (id - src_id0) / per_src, position from src_table, detector hits summed per (source, detector)
    const float* src_table;           // n_src x 3, nullptr = the single source of smcrt_set_source
    unsigned long long* src_tot;      // n_src x n_det Q40.24 totals
    unsigned long long src_id0;
    long long per_src;
    float sp[24];
    float Tpos[12];  // local emit position -> world (row-vector affine folded to 3x4)
    float Tdir[9];   // local emit direction -> world (3x3)
    // tallies
    float* jmean;
    float* absorb;
    float* emission;
    unsigned long long* det_bins;  // Q40.24 fixed point
    int det_total;
    int det_in_smem;
    unsigned long long* counters;  // [nscatt, sweeps, bounces, launched, retries, lost, n_top*sweeps(unused), det_hits]
    unsigned long long* next;      // work counter
    unsigned long long* tstamp;    // variant trial only: [0] = %globaltimer at kernel start, [1] = when the last packet id is claimed
    long long nphotons;
    unsigned long long id_offset;
    unsigned long long rec_id0;  // first packet id of the RUN (a run may be several launches): per-packet records are indexed by pid - rec_id0
    uint32_t seed_lo, seed_hi;
    int tally_mode, survival;
    float threshold, chance;
    float eps0, eps_rel;
    int max_steps;
    const FirstFlight* ff;  // nullptr: every packet takes the ordinary launch sweeps
    unsigned long long watchdog_ns;  // trace_queued: how long a warp may find the queues empty / a publication pending (%globaltimer)
    int xchg_off;   // byte offset of the compaction scratch in dynamic shared memory (16-byte aligned)
    int dda_legacy;  // 1: path-length deposits one red.global.add.f32 per voxel crossed (SMCRT_DDA_LEGACY; cross-check of the run walker)
    // run walker (walk_dda_runs): fixed-point difference grids per axis, their scale (2^28 / voxel edge) and "touched" flags
    long long* jdiff[3];
    float jfix[3];
    unsigned int* jdiff_used;
    // -Dpathlength: the trace kernels do not walk voxels, they RECORD the straight segments; deposit_segments_kernel
    // walks them.  Every CTA owns seg_cap records of seg_buf (2 x float4 each: start + length, direction + weight) and counts
    // them in shared memory (byte offset seg_off of its dynamic shared memory); the count goes to seg_count[blockIdx.x] at the end.
    float4* seg_buf;
    unsigned int* seg_count;
    unsigned int seg_cap;
    int seg_off;
    unsigned int* seg_work;         // work counter of deposit_segments_kernel (cleared with the shares)
    float seg_piece;                // the deposit kernel cuts segments into pieces of about this many voxel visits, one piece per lane
    unsigned long long* seg_total;  // all segments produced (recorded or, when a CTA's share was full, walked inline)
    // optional per-packet outputs (smcrt_trace_packets)
    int* out_fate;
    int* out_nscatt;
    int* out_events;
    int* out_sweeps;
    float* out_pos;
    // trackHistory (src/historyStack.f90): the first pass only notes WHICH packets hit a history-tracking detector
    // (hist_ids / hist_det, cursor hist_n, capacity hist_cap); their vertex lists are produced by tracing those few packets again
    // (streams depend on (seed, id) only): id_list = the packet ids of the replay (sorted; per-packet records are indexed by the
    // position in it), out_vert / out_nvert / out_hit = max_vert vertices (x, y, z, scatter index) per packet, how many were
    // written, and the vertex count at the first tracked hit (-1: none).
    unsigned long long* hist_ids;
    int* hist_det;
    unsigned long long* hist_n;
    unsigned long long hist_cap;
    const unsigned long long* id_list;
    long long id_list_n;
    float4* out_vert;
    int* out_nvert;
    int* out_hit;
    int max_vert;
    float* out_dbg;  // 12 floats per packet, written when a packet hits the step cap (engine diagnostics)
    long long dbg_pid;  // engine diagnostics: log the boundary events of this one packet into dbg_log (16 floats each)
    float* dbg_log;
    int dbg_cap;
};

enum : int { C_NSCATT = 0, C_SWEEPS, C_BOUNCES, C_LAUNCHED, C_RETRIES, C_LOST, C_SPARE /* trace_queued watchdog */, C_DETHITS, C_VOXELS, C_REDS, C_COUNT };
enum : int { TALLY_ABSORB = 1, TALLY_PATHLENGTH = 2, TALLY_EMISSION = 4 };
constexpr float TWOPI_F = 6.283185307179586f;
constexpr float DET_FIX = 16777216.0f;  // 2^24: detector bins are Q40.24 fixed point

// ------------------------------------------------------------------------------------------------ RNG
// Philox4x32-10 (Salmon et al. 2011).  counter = (event, id_lo, id_hi, 0), key = (seed_lo, seed_hi):
// one 4-word block per packet "event" (emit attempt / interaction / Fresnel), DESIGN.md §5.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), hi1 = __umulhi(0xCD9E8D57u, c2);
#else
        const uint32_t hi0 = (uint32_t)(((uint64_t)0xD2511F53u * c0) >> 32), hi1 = (uint32_t)(((uint64_t)0xCD9E8D57u * c2) >> 32);
#endif
        const uint32_t lo0 = 0xD2511F53u * c0, lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t w) { return __fmul_rn((float)(w >> 8), 5.9604644775390625e-08f); }  // [0,1)
__device__ __forceinline__ float u01_open0(uint32_t w) {                                                           // (0,1]
    return __fmul_rn(__fadd_rn(__uint2float_rn(w), 1.0f), 2.3283064365386963e-10f);
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------------------------ scene view
struct SceneView {
    const DevPrim* prims;
    const DevTop* tops;
    const DevInstr* prog;
    const DevDet* dets;
    const float4* hot;  // 2 x float4 per top-level SDF: the sweep's view of it (see DevHot in device_scene.cuh)
    const float4* detp; // one float4 per detector: its plane (n, n.p0), all the crossing pre-test needs
    int n_top, n_det;
};
__device__ __forceinline__ SceneView make_view(const unsigned char* base, const KParams& P) {
    SceneView sc;
    sc.prims = reinterpret_cast<const DevPrim*>(base);
    sc.tops = reinterpret_cast<const DevTop*>(base + P.off_tops);
    sc.prog = reinterpret_cast<const DevInstr*>(base + P.off_prog);
    sc.dets = reinterpret_cast<const DevDet*>(base + P.off_dets);
    sc.hot = reinterpret_cast<const float4*>(base + P.off_hot);
    sc.detp = reinterpret_cast<const float4*>(base + P.off_detp);
    sc.n_top = P.n_top;
    sc.n_det = P.n_det;
    return sc;
}

__device__ __forceinline__ float eval_top_f(const SceneView& sc, int t, float x, float y, float z) {
    const DevTop T = sc.tops[t];
    if (T.mode == 0) return eval_prim<float>(sc.prims[T.first], x, y, z);
    return eval_program<float, DevPrim, DevInstr>(sc.prims, sc.prog, T.first, T.count, x, y, z);
}
__device__ __forceinline__ double eval_top_d(const KParams& P, const SceneView& sc, int t, double x, double y, double z) {
    const DevTop T = sc.tops[t];
    if (T.mode == 0) return eval_prim<double>(P.primsD[T.first], x, y, z);
    return eval_program<double, DevPrimD, DevInstrD>(P.primsD, P.progD, T.first, T.count, x, y, z);
}
// calcNormal (src/sdfs/sdf_base.f90:166-190): tetrahedral 4-tap gradient with h = 1e-6, in FP64 like the
// reference (h is far below FP32 resolution); only executed at refractive-index-mismatch crossings.
__device__ __forceinline__ double3 surface_normal_fd(const KParams& P, const SceneView& sc, int t, double X, double Y, double Z) {
    const double h = 1e-6;
    const double f1 = eval_top_d(P, sc, t, X + h, Y - h, Z - h);
    const double f2 = eval_top_d(P, sc, t, X - h, Y - h, Z + h);
    const double f3 = eval_top_d(P, sc, t, X - h, Y + h, Z - h);
    const double f4 = eval_top_d(P, sc, t, X + h, Y + h, Z + h);
    double nx = f1 - f2 - f3 + f4, ny = -f1 - f2 + f3 + f4, nz = -f1 + f2 - f3 + f4;
    const double il = rsqrt(nx * nx + ny * ny + nz * nz);
    return make_double3(nx * il, ny * il, nz * il);
}

// FP64 "polish" of a boundary hit.  The FP32 march stops within eps ~ 4 ulp(|pos|) of the surface, the
// reference within 1e-8.  Reflecting/refracting a few 1e-7 off the surface perturbs the impact parameter of every bounce,
// which pumps packets into whispering-gallery orbits of (unions of) spheres ~100x faster than the reference does.  One
// Newton step of the FP64 distance along the ray puts the packet on the surface to ~1e-12 before the normal is taken.
// Returns the signed distance moved along the ray (|t| <= tmax).
__device__ __forceinline__ double polish_hit_newton(const KParams& P, const SceneView& sc, int t, double X, double Y, double Z, double ux, double uy,
                                                    double uz, double tmax) {
    const double h = 1e-6;
    double moved = 0.0;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {  // two Newton steps: the second removes the curvature term of the first
        const double f0 = eval_top_d(P, sc, t, X, Y, Z);
        const double f1 = eval_top_d(P, sc, t, X + h * ux, Y + h * uy, Z + h * uz);
        const double g = (f1 - f0) / h;
        if (!(fabs(g) > 1e-3)) break;  // grazing: the along-ray root is ill-conditioned, keep the landing point
        const double dt = fmin(fmax(-f0 / g, -tmax - moved), tmax - moved);
        X += dt * ux; Y += dt * uy; Z += dt * uz;
        moved += dt;
    }
    return moved;
}

// Boundary hit geometry of a Fresnel event: the along-ray move that puts the packet ON surface `t` (polish_hit) and the outward
// surface normal there (surface_normal), in one call.  Single untransformed / translated spheres and boxes -- every surface of
// the sphere, slab, skin and jacques scenes -- have both in closed form: the ray/sphere root and o/|o|; the face plane of a box and
// its axis.  The closed forms are what the Newton polish and the h = 1e-6 four-tap gradient converge to (difference O(h^2/r^2)
// resp. rounding), at ~1/6 of the FP64 work: Fresnel events run with ~2 of 32 lanes active, so their instruction count is what
// a refractive scene pays for them.  Everything else (and a box hit closer to an edge than the probe reach) takes the generic path.
// Returned in registers (<= 3 doubles each: a larger struct would come back through the caller's stack frame), hence two calls:
// polish_hit() = the along-ray move, surface_normal() = the normal at the moved point; same signatures as the generic pair.
__device__ __forceinline__ bool closed_form_prim(const KParams& P, const SceneView& sc, int t, const DevPrimD*& Q) {
    const DevTop T = sc.tops[t];
    if (T.mode != 0) return false;
    Q = P.primsD + T.first;
    return Q->xf <= XF_TRANSLATE && (Q->kind == 1 || Q->kind == 2);
}
static __device__ __noinline__ double polish_hit(const KParams& P, const SceneView& sc, int t, double X, double Y, double Z, double ux, double uy,
                                          double uz, double tmax) {
    const DevPrimD* Q;
    if (closed_form_prim(P, sc, t, Q)) {
        double ox = X, oy = Y, oz = Z;
        if (Q->xf == XF_TRANSLATE) { ox += Q->m[3]; oy += Q->m[7]; oz += Q->m[11]; }
        if (Q->kind == 1) {  // sphere: t^2 + 2 b t + c = 0, the root next to the packet
            const double R = Q->p[0];
            const double len2 = ox * ox + oy * oy + oz * oz, b = ox * ux + oy * uy + oz * uz;
            const double disc = R * R - (len2 - b * b);
            if (!(b * b > 1e-6 * len2 && disc > 0.0)) return 0.0;  // |df/dt| = |b|/|o| <= 1e-3: same grazing guard as polish_hit
            const double s = sqrt(disc);
            return fmin(fmax(b > 0.0 ? s - b : -b - s, -tmax), tmax);
        }
        // box: planar while the two other face distances stay negative over the probe reach and the normal's taps
        const double dx = fabs(ox) - Q->p[0], dy = fabs(oy) - Q->p[1], dz = fabs(oz) - Q->p[2];
        const bool fx = dx >= dy && dx >= dz, fy = !fx && dy >= dz;
        const double da = fx ? dx : (fy ? dy : dz), db = fx ? dy : dx, dc = (fx || fy) ? dz : dy;
        const double reach = fmin(da, 0.0) - (tmax + 4e-6);
        if (db < reach && dc < reach) {
            const double oa = fx ? ox : (fy ? oy : oz), ua = fx ? ux : (fy ? uy : uz);
            const double g = oa < 0.0 ? -ua : ua;
            return fabs(g) > 1e-3 ? fmin(fmax(-da / g, -tmax), tmax) : 0.0;
        }
    }
    return polish_hit_newton(P, sc, t, X, Y, Z, ux, uy, uz, tmax);
}
static __device__ __noinline__ double3 surface_normal(const KParams& P, const SceneView& sc, int t, double X, double Y, double Z) {
    const DevPrimD* Q;
    if (closed_form_prim(P, sc, t, Q)) {
        double ox = X, oy = Y, oz = Z;
        if (Q->xf == XF_TRANSLATE) { ox += Q->m[3]; oy += Q->m[7]; oz += Q->m[11]; }
        if (Q->kind == 1) {
            // The reference's stencil k_i = (+,-,-), (-,-,+), (-,+,-), (+,+,+) is not central: with f = |o| - r,
            //   sum_i k_i f(o + h k_i) = 4h [ n - (h/|o|) (ny nz, nx nz, nx ny) ] + O(h^3),  n = o/|o|,
            // i.e. calcNormal is tilted by ~h/r against the true normal (1e-3 for the smallest spheres of sphere.toml).  The tilt
            // is part of the reference's deterministic output, so it is reproduced (residual O((h/|o|)^2)).
            const double il = rsqrt(ox * ox + oy * oy + oz * oz);
            const double a = ox * il, b = oy * il, c = oz * il, e = 1e-6 * il;
            const double nx = a - e * b * c, ny = b - e * a * c, nz = c - e * a * b;
            const double in = rsqrt(nx * nx + ny * ny + nz * nz);
            return make_double3(nx * in, ny * in, nz * in);
        }
        const double dx = fabs(ox) - Q->p[0], dy = fabs(oy) - Q->p[1], dz = fabs(oz) - Q->p[2];
        const bool fx = dx >= dy && dx >= dz, fy = !fx && dy >= dz;
        const double db = fx ? dy : dx, dc = (fx || fy) ? dz : dy;
        const double lim = fmin(fx ? dx : (fy ? dy : dz), 0.0) - 4e-6;
        if (db < lim && dc < lim) {  // every tap of the four-tap gradient sees the same face (outside: no edge region; inside: same argmax)
            const double sg = (fx ? ox : (fy ? oy : oz)) < 0.0 ? -1.0 : 1.0;
            return make_double3(fx ? sg : 0.0, fy ? sg : 0.0, (fx || fy) ? 0.0 : sg);
        }
    }
    return surface_normal_fd(P, sc, t, X, Y, Z);
}

// Evaluate ALL top-level SDFs at (x,y,z): min|d|, min d, argmax of the negatives (the reference's
// maxloc(ds, mask=ds<0): innermost surface wins, ties -> lowest index, none -> 0), value there, and the value of
// SDF `layer` (1-based).   src/inttau2.f90:63-68,80-84,135-139,179-183,216-221,229-234
struct Sweep {
    float amin, smin, dL, bmin;
    int L;
    bool bexact;
};
// distance, directional step bound and exact flag of top-level SDF i.  SIMPLE: the scene holds translated spheres and boxes only (the
// host checked): the general evaluators are not even compiled into the loop (4 % on the slab: the hot loop is I-cache sensitive).
template <bool SIMPLE>
__device__ __forceinline__ void top_ray(const SceneView& sc, int i, float x, float y, float z, float ux, float uy, float uz, float need, float& d,
                                        float& b, bool& ex) {
    // two 16-byte shared loads bring everything a translated sphere or box needs (no tops[] -> prims[] indirection, no
    // transform-class / kind ladder); every other kind goes through the out-of-line general evaluator
    const float4 h0 = sc.hot[2 * i], h1 = sc.hot[2 * i + 1];
    const int code = __float_as_int(h0.x);
    if (code == HOT_SPHERE) {
        d = sphere_ray(x + h0.y, y + h0.z, z + h0.w, ux, uy, uz, h1.x, b);
        ex = true;
    } else if (code == HOT_BOX) {
        d = box_ray(x + h0.y, y + h0.z, z + h0.w, ux, uy, uz, h1.x, h1.y, h1.z, b);
        ex = true;
    } else if (SIMPLE) {
        d = b = SMCRT_BIG; ex = false;  // not reached
    } else if (code == HOT_PROGRAM) {  // compound `model`: plain sphere tracing
        d = eval_program<float, DevPrim, DevInstr>(sc.prims, sc.prog, __float_as_int(h0.y), __float_as_int(h0.z), x, y, z);
        b = fabsf(d);
        ex = false;
    } else {
        const float3 r = eval_prim_ray_general(sc.prims + __float_as_int(h0.y), x, y, z, ux, uy, uz, need);
        d = r.x; b = r.y; ex = r.z != 0.f;
    }
}
template <bool ANY_ORDER, bool SIMPLE>
__device__ __forceinline__ void sweep_one(const SceneView& sc, int i, float x, float y, float z, float ux, float uy, float uz, float need, Sweep& s) {
    float d, b;
    bool ex;
    top_ray<SIMPLE>(sc, i, x, y, z, ux, uy, uz, need, d, b, ex);
    s.amin = fminf(s.amin, fabsf(d));
    s.smin = fminf(s.smin, d);
    if (b < s.bmin) { s.bmin = b; s.bexact = ex; }
    // innermost negative; ties -> lowest index (maxloc).  Culled lists are sorted by code path, not by index: explicit compare there
    if (ANY_ORDER) { if (d < 0.f && (d > s.dL || (d == s.dL && i + 1 < s.L))) { s.dL = d; s.L = i + 1; } }
    else if (d < 0.f && d > s.dL) { s.dL = d; s.L = i + 1; }
}
// Culled sweep: only the SDFs that can matter for a point of this coarse cell are evaluated.  The host builds, per cell, the
// list of SDFs that can attain min|d| or be the innermost negative one somewhere in the cell (interval bounds from the value at
// the cell centre and 1-Lipschitz continuity), plus `far`: a lower bound of |d| of all the others.  The min / argmax over the
// list equals the min / argmax over all SDFs for every point of the cell, and a step is additionally capped by `far`.
template <bool SIMPLE>
__device__ __forceinline__ Sweep sweep_all(const KParams& P, const SceneView& sc, float x, float y, float z, float ux, float uy, float uz, float need) {
    Sweep s;
    s.amin = SMCRT_BIG; s.smin = SMCRT_BIG; s.dL = -SMCRT_BIG; s.L = 0; s.bmin = SMCRT_BIG; s.bexact = false;
    if (P.cull_start) {
        const int cx = (int)floorf((x - P.cull_lo[0]) * P.cull_inv[0]), cy = (int)floorf((y - P.cull_lo[1]) * P.cull_inv[1]),
                  cz = (int)floorf((z - P.cull_lo[2]) * P.cull_inv[2]);
        if (cx >= 0 && cx < P.cull_n[0] && cy >= 0 && cy < P.cull_n[1] && cz >= 0 && cz < P.cull_n[2]) {
            const int c = cx + P.cull_n[0] * (cy + P.cull_n[1] * cz);
            if (need < SMCRT_BIG) {
                // Clear cell: no surface comes closer than `cl` to any point of it.  A packet whose optical depth runs out within
                // `need` < cl interacts before it can reach one: no SDF is evaluated at all (a scattering medium between sparse
                // bodies -- the dermis around the vessels -- spends most of its sweeps here).  The step decision sees a surface
                // at distance cl and takes the interaction branch, exactly as it would with the true distances.
                const float cl = __ldg(P.cull_clear + c);
                if (need < cl) { s.amin = cl; s.smin = -cl; s.bmin = cl; return s; }
            }
            const int i0 = __ldg(P.cull_start + c), i1 = __ldg(P.cull_start + c + 1);
            for (int k = i0; k < i1; ++k) sweep_one<true, SIMPLE>(sc, __ldg(P.cull_items + k), x, y, z, ux, uy, uz, need, s);
            const float far = __ldg(P.cull_far + c);
            if (far < s.bmin) { s.bmin = fmaxf(far, s.amin); s.bexact = false; }  // never step past an unlisted surface
            return s;
        }
    }
    const int n = sc.n_top;
    for (int i = 0; i < n; ++i) sweep_one<false, SIMPLE>(sc, i, x, y, z, ux, uy, uz, need, s);
    return s;
}

// ------------------------------------------------------------------------------------------------ Fresnel
// src/surfaces.f90:14-127 (fresnel :86-127, reflect :42-55, refract :57-84).  Returns the coefficient; sets rflag.
// Evaluated in FP64: it runs once per index-mismatch crossing (cold), and R(theta) is ill-conditioned next to the
// total-internal-reflection knee, where FP32 rounding of I.N alone moves R by ~1e-5.
struct Refl {
    float x, y, z, R;
    bool reflected;
};
static __device__ __noinline__ Refl reflect_refract(float ux, float uy, float uz, double3 N, float n1f, float n2f, float xi) {
    const double n1 = n1f, n2 = n2f;
    const double dx = ux, dy = uy, dz = uz;
    const double idn = dx * N.x + dy * N.y + dz * N.z;
    const double costt = fmin(fabs(idn), 1.0);
    const double eta = n1 / n2;
    // sin(theta_t)^2 = eta^2 (1 - cos^2): the TIR test and cos(theta_t) need no sin(theta_i) (one FP64 sqrt instead of three)
    const double sin2t = eta * eta * (1.0 - costt * costt);
    double R, cost2 = 0.0;
    if (sin2t > 1.0) R = 1.0;            // total internal reflection
    else if (costt == 1.0) R = 0.0;      // exactly normal incidence: transmitted (reference quirk Q10)
    else {
        cost2 = sqrt(1.0 - sin2t);
        const double an = n1 * costt - n2 * cost2, ad = n1 * costt + n2 * cost2;
        const double bn = n1 * cost2 - n2 * costt, bd = n1 * cost2 + n2 * costt;
        const double inv = 1.0 / (ad * bd);  // one division for both amplitude ratios
        const double a = an * bd * inv, b = bn * ad * inv;
        R = 0.5 * (a * a + b * b);
    }
    Refl o;
    o.R = (float)R;
    if ((double)xi <= R) {  // reflect: I - 2 (N.I) N
        o.reflected = true;
        const double k = 2.0 * idn;
        o.x = (float)(dx - k * N.x); o.y = (float)(dy - k * N.y); o.z = (float)(dz - k * N.z);
    } else {  // refract with the normal flipped to oppose I
        o.reflected = false;
        double c1 = idn, sg = 1.0;
        if (c1 < 0.0) c1 = -c1;
        else sg = -1.0;
        // cos(theta_t): the value of the coefficient branch unless |I.N| was clamped or the incidence is exactly normal
        const double c2 = (cost2 > 0.0 && c1 <= 1.0) ? cost2 : sqrt(1.0 - eta * eta * (1.0 - c1 * c1));
        const double k = (eta * c1 - c2) * sg;
        o.x = (float)(eta * dx + k * N.x); o.y = (float)(eta * dy + k * N.y); o.z = (float)(eta * dz + k * N.z);
    }
    return o;
}

// ------------------------------------------------------------------------------------------------ scatter
// photon%scatter, src/photon.f90:1045-1103 (mcxyz direction update). xi_c: cos(theta) draw, xi_p: phi draw.
__device__ __forceinline__ void hg_scatter(float& dx_, float& dy_, float& dz_, float hgg, float xi_c, float xi_p) {
    // FP32-first algebra: the reference computes cos(theta) and then sin = sqrt(1 - cos^2), which cancels for the
    // forward-peaked angles HG favours.  Here 1 - cos(theta) is formed without cancellation:
    //   g = 0 : 1 - cos = 2 (1 - xi)
    //   g != 0: t = (1-g^2)/(1-g+2 g xi), cos = (1+g^2-t^2)/(2g)  =>  1 - cos = (1-g)(1-xi)(t+1-g)/(1-g+2 g xi)
    //   and symmetrically          1 + cos = (1+g) xi (t+1+g)/(1-g+2 g xi)          (back-scattering, g < 0)
    float omc, opc;
    if (hgg == 0.0f) {
        omc = 2.0f * (1.0f - xi_c);
        opc = 2.0f * xi_c;
    } else {
        const float den = 1.0f - hgg + 2.0f * hgg * xi_c;
        const float t = (1.0f - hgg * hgg) / den;
        omc = (1.0f - hgg) * (1.0f - xi_c) * (t + 1.0f - hgg) / den;
        opc = (1.0f + hgg) * xi_c * (t + 1.0f + hgg) / den;
    }
    omc = fminf(fmaxf(omc, 0.0f), 2.0f);
    opc = fminf(fmaxf(opc, 0.0f), 2.0f);
    const float cost = omc < 1.0f ? 1.0f - omc : opc - 1.0f;
    const float sint = sqrtf(omc * opc);
    float sinp, cosp;
    sincospif(2.0f * xi_p, &sinp, &cosp);  // exact range reduction, no slow path
    float ux, uy, uz;
    const float nx = dx_, ny = dy_, nz = dz_;
    // sqrt(1 - nz^2) of the reference == sqrt(nx^2 + ny^2) for a unit vector, without the cancellation near the poles.
    // The reference switches to the polar form at |nz| > 1 - 1e-12, i.e. (in FP32) when nx = ny = 0 to rounding.
    const float t2 = nx * nx + ny * ny;
    if (t2 < 1e-12f) {
        ux = sint * cosp; uy = sint * sinp; uz = nz > 0.f ? cost : -cost;
    } else {
        const float it = rsqrtf(t2);
        const float t = t2 * it;
        ux = sint * ((nx * nz * cosp - ny * sinp) * it) + nx * cost;
        uy = sint * ((ny * nz * cosp + nx * sinp) * it) + ny * cost;
        uz = -sint * cosp * t + nz * cost;
    }
    // :1091-1097 renormalises until |len-1| <= 1e-12; one FP32 normalisation is the same operation at FP32 resolution
    const float l2 = ux * ux + uy * uy + uz * uz;
    const float il = rsqrtf(l2);
    const float il2 = il * (1.5f - 0.5f * l2 * il * il);  // one Newton step: rsqrt.approx is only ~2 ulp
    dx_ = ux * il2; dy_ = uy * il2; dz_ = uz * il2;
}

// ------------------------------------------------------------------------------------------------ voxels
__device__ __forceinline__ bool in_grid(const KParams& P, float x, float y, float z) {
    // update_voxels (src/inttau2.f90:587-614): cell = floor(n (x + max) / (2 max)) + 1 is valid  <=>  -max <= x < max.
    // Compared on the coordinate itself: the scaled form rounds up to n in FP32 for x within ~n ulp of the upper face.
    // CLOSED at the upper face too: a medium surface that coincides with the grid face (the side walls of the skin stack, the slab
    // scenes) puts a packet that reflects there ON the face -- x == max in FP32, a set of measure zero at the reference's FP64 --
    // and the half-open test killed it on the +x, +y, +z faces only (2 % of the packets of skin_b200.toml, none on the minus
    // faces).  voxel_of / dda_start clamp the index of such a point to the last cell.
    return fabsf(x) <= P.gmax[0] && fabsf(y) <= P.gmax[1] && fabsf(z) <= P.gmax[2];
}
__device__ __forceinline__ long long voxel_of(const KParams& P, float x, float y, float z) {
    int i = (int)floorf((x + P.gmax[0]) * P.inv_vox[0]);
    int j = (int)floorf((y + P.gmax[1]) * P.inv_vox[1]);
    int k = (int)floorf((z + P.gmax[2]) * P.inv_vox[2]);
    i = min(max(i, 0), P.nxg - 1); j = min(max(j, 0), P.nyg - 1); k = min(max(k, 0), P.nzg - 1);
    return (long long)i + (long long)P.nxg * ((long long)j + (long long)P.nyg * (long long)k);
}
// The deposits are reductions into GLOBAL memory whose result nobody reads: red.global.  Said in PTX, because in the walkers (not
// inlined, the grids reached through a KParams reference) the compiler cannot prove the address space, and atomicAdd() then becomes
// a generic ATOM -- a shared-window test with a CAS loop beside it, and an atomic WITH a reply, which the warp's scoreboard waits
// for (long-scoreboard stall 5.9 per issued instruction in the deposit kernel on skin_b200.toml; lts op_atom, not op_red).
__device__ __forceinline__ void red_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void red_i64(long long* p, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
}
// Warp-aggregated deposit: lanes of the converged group that hit the same voxel are summed with shuffles and
// ONE red.global.add.f32 is issued per distinct voxel.
// NB must stay inlined: __activemask() inside a called (noinline) function does not name the lanes that called it together,
// and the aggregation then loses deposits (measured: wrong emission totals) besides being 25 % slower.
__device__ __forceinline__ void deposit(float* grid, long long vox, float w) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, (unsigned long long)(grid + vox));  // the ADDRESS: lanes of different grids never merge
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    float sum = w;
    if (peers != (1u << lane)) {
        sum = 0.f;
        unsigned rem = peers;
        while (rem) {  // every peer executes the same number of shuffles
            const int src = __ffs(rem) - 1;
            sum += __shfl_sync(peers, w, src);
            rem &= rem - 1;
        }
    }
    if (lane == leader) red_f32(grid + vox, sum);
}

// update_grids in -Dpathlength mode (src/inttau2.f90:408-445): deposits (segment length * weight) into every voxel the straight
// piece crosses.  Returns true when the walk starts or ends outside the grid (packet dies).
// (Default build: only the end-of-step voxel matters, see the WALK macro of the kernel.)  Arguments by value: a pointer
// argument would force the caller's state into local memory.
//
// Cell faces are taken centre-origin, face(i) = (2i - n) * (max / n), so their rounding error scales with the face's own
// coordinate and not with max: with the corner-origin form (x + max) a packet on the beam axis of a wide grid
// (|x| << max) sees face - x wrong by ulp(max), and dividing that by a small direction cosine gives an arbitrarily
// large NEGATIVE first crossing distance, which the loop below would deposit as path that was never travelled.
struct DdaStart {
    int c[3];
    float t[3], dt[3];
};
__device__ __forceinline__ DdaStart dda_start(const KParams& P, float fx, float fy, float fz, float dx, float dy, float dz) {
    const float f3[3] = {fx, fy, fz}, d3[3] = {dx, dy, dz};
    const int n3[3] = {P.nxg, P.nyg, P.nzg};
    const float BIG = 3.0e38f;
    DdaStart s;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float hv = P.hvox[a];
        int c = (int)floorf(fmaf(f3[a], P.inv_vox[a], 0.5f * (float)n3[a]));
        c = min(max(c, 0), n3[a] - 1);
        if (c > 0 && f3[a] < (float)(2 * c - n3[a]) * hv) --c;                       // make the cell agree with its own faces
        else if (c < n3[a] - 1 && f3[a] >= (float)(2 * c + 2 - n3[a]) * hv) ++c;
        s.c[a] = c;
        const float face = (float)(2 * (c + (d3[a] > 0.f)) - n3[a]) * hv;
        s.t[a] = d3[a] != 0.f ? fmaxf((face - f3[a]) / d3[a], 0.f) : BIG;
        s.dt[a] = d3[a] != 0.f ? P.vox[a] / fabsf(d3[a]) : BIG;
    }
    return s;
}
// The deposits of one straight segment (start f, unit direction d, length len), one call per SEGMENT (SEGMENT_END of the kernel),
// not per piece: a packet of sphere.toml moves in 15 pieces but changes direction three times.
//
// Voxel walk (the reference's loop, :417-441): an Amanatides-Woo DDA with one red.global.add.f32 per voxel crossed.
//
// Run walk.  One red per voxel crossed cannot go faster than the L2 atomic units (~2e11 red/s when the lines stay
// in L2, 6e10 into one column), and a packet of the slab scene crosses ~600 voxels.  But a straight segment is a sequence of RUNS:
// maximal stretches inside one voxel column along the ray's dominant axis `a` (the axis whose faces come fastest).  A run that
// enters voxel k_in, crosses m faces and ends in voxel k_out deposits
//        p_in, c, c, ..., c, p_out            c = vox_a / |u_a| (the full chord), p_in / p_out the partial end pieces
// -- a range update.  It is issued as FOUR integer atomics into a fixed-point DIFFERENCE grid of axis a,
//        D[lo] += p_lo;  D[lo+1] += c - p_lo;  D[hi] += p_hi - c;  D[hi+1] -= p_hi,
// whose prefix sum along a (jdiff_scan_kernel, run once before the grid is read) is exactly that deposit sequence: integer adds
// are associative, so the telescoping is exact whatever the order of the atomics.  The loop iterates over COLUMN CHANGES, not
// voxels: a segment of the slab's pencil beam (166 voxels) costs 4 atomics.  Runs of fewer than 4 voxels are deposited directly
// (f32 red into jmean), and a segment whose runs are expected to be that short -- oblique to the grid, or only a few voxels long:
// the short free paths of a turbid medium -- takes the plain voxel walk, which has less set-up.  Same voxels, same lengths as the
// voxel walk up to the rounding of a face time (ta + j dta by one fma instead of j additions).
// Fixed point: 2^28 units per voxel edge of the run's axis (3.7e-9 relative), weight folded in; |D| < 2^63 holds for 2^32
// full-chord deposits into one entry, and the engine scans the difference grids at least every 2^32 packets.
// CTA-private accumulator of the deposit kernel for HOT difference-grid entries ("per-CTA shared-memory tile for the
// beam column", generalised).  A pencil beam sends 1e9 range updates per 1e8 packets into the ~1300 entries of four voxel columns,
// and L2 takes same-entry atomics one by one (2.5e10/s measured: the ceiling of r01_red_peaks.json for one column).  An open-
// addressing table in shared memory -- key = the entry's address, first come first kept, at most HOT_PROBES probes -- takes the
// updates of the entries it holds (two native 32-bit shared adds with carry) and is flushed once per CTA; an entry that found no
// slot goes to L2 as before, so a diffuse scene (millions of distinct entries) loses only the probes.
// 2048 slots = 32 KB: room for the slab's ~1300 hot entries; 4096 slots cost the other scenes 2-3 % (the shared memory comes out of
// the L1 that serves the kernel's spills and records) and gained the slab 0.6 %
constexpr int HOT_BITS = 11, HOT_SLOTS = 1 << HOT_BITS, HOT_PROBES = 3;
constexpr int WARPQ_CAP = 32;  // records in each of a warp's two queues of the deposit kernel (2 x float4 per record)
struct HotTable {
    unsigned long long key[HOT_SLOTS];
    unsigned int lo[HOT_SLOTS], hi[HOT_SLOTS];
};
__device__ __forceinline__ bool red_i64_hot(HotTable* T, long long* p, long long v) {  // true: the table took it
    const unsigned long long k = (unsigned long long)p;
    unsigned int h = (unsigned int)((k >> 3) * 0x9E3779B97F4A7C15ull >> (64 - HOT_BITS));
#pragma unroll
    for (int probe = 0; probe < HOT_PROBES; ++probe) {
        unsigned long long cur = T->key[h];
        if (cur == 0ull) cur = atomicCAS(&T->key[h], 0ull, k), cur = cur ? cur : k;
        if (cur == k) {
            const unsigned int vl = (unsigned int)(unsigned long long)v, vh = (unsigned int)((unsigned long long)v >> 32);
            const unsigned int old = atomicAdd(&T->lo[h], vl);
            atomicAdd(&T->hi[h], vh + ((old + vl < old) ? 1u : 0u));
            return true;
        }
        h = (h + 1u) & (HOT_SLOTS - 1);
    }
    red_i64(p, v);
    return false;
}
// Two functions (the deposit kernel knows which one a segment takes and calls it directly: each gets its own register allocation).
// walk_voxels: the voxel walk.  Returns (voxels visited, atomics issued).
__device__ __forceinline__ uint2 walk_voxels_inl(const KParams& P, float fx, float fy, float fz, float dx, float dy, float dz, float len, float weight) {
    if (!in_grid(P, fx, fy, fz)) return make_uint2(0u, 0u);  // :411-415
    const DdaStart S = dda_start(P, fx, fy, fz, dx, dy, dz);
    const long long vy = (long long)P.nxg, vz = (long long)P.nxg * (long long)P.nyg;
    const long long off = (long long)S.c[0] + vy * (long long)S.c[1] + vz * (long long)S.c[2];
    // faces that can still be crossed along each axis before the walk leaves the grid
    int rx = dx > 0.f ? P.nxg - 1 - S.c[0] : S.c[0], ry = dy > 0.f ? P.nyg - 1 - S.c[1] : S.c[1], rz = dz > 0.f ? P.nzg - 1 - S.c[2] : S.c[2];
    float tx = S.t[0], ty = S.t[1], tz = S.t[2];
    const float dtx = S.dt[0], dty = S.dt[1], dtz = S.dt[2];
    float t = 0.f;
    // the voxel's flat index is carried along (one 64-bit add per crossing instead of two 64-bit multiply-adds)
    const long long svx = dx > 0.f ? 1ll : -1ll, svy = dy > 0.f ? vy : -vy, svz = dz > 0.f ? vz : -vz;
    float* cell = P.jmean + off;
    unsigned int nvox = 0u;
    for (;;) {
        const float tn = fminf(tx, fminf(ty, tz));
        ++nvox;
        if (tn >= len) {
            red_f32(cell, fmaxf(len - t, 0.f) * weight);
            break;
        }
        red_f32(cell, fmaxf(tn - t, 0.f) * weight);
        t = tn;
        // which face: x before y before z on a tie.  Selects, not branches: the lanes of a warp walk different rays
        const bool stx = tx <= ty && tx <= tz, sty = !stx && ty <= tz, stz = !stx && !sty;
        cell += stx ? svx : (sty ? svy : svz);
        tx += stx ? dtx : 0.f; ty += sty ? dty : 0.f; tz += stz ? dtz : 0.f;
        rx -= stx ? 1 : 0; ry -= sty ? 1 : 0; rz -= stz ? 1 : 0;
        if ((rx | ry | rz) < 0) break;  // :437-440
    }
    return make_uint2(nvox, nvox);
}
// (called where the walk is not the hot path; the deposit kernel's voxel-walk queue inlines the body: its KParams reads are then
// constant-bank loads and no registers are saved around a call)
static __device__ __noinline__ uint2 walk_voxels(const KParams& P, float fx, float fy, float fz, float dx, float dy, float dz, float len, float weight) {
    return walk_voxels_inl(P, fx, fy, fz, dx, dy, dz, len, weight);
}
// walk_runs: the run walk.  `hot`: the deposit kernel's table (nullptr: every range update goes straight to L2)
// Returns (voxels visited, atomics issued, range-update entries offered to the table, entries it took).
static __device__ __noinline__ uint4 walk_runs(const KParams& P, float fx, float fy, float fz, float dx, float dy, float dz, float len, float weight,
                                               HotTable* hot) {
    if (!in_grid(P, fx, fy, fz)) return make_uint4(0u, 0u, 0u, 0u);  // :411-415
    const DdaStart S = dda_start(P, fx, fy, fz, dx, dy, dz);
    const long long vy = (long long)P.nxg, vz = (long long)P.nxg * (long long)P.nyg;
    long long off = (long long)S.c[0] + vy * (long long)S.c[1] + vz * (long long)S.c[2];
    // dominant axis a; the two others b, c in index order
    const int a = S.dt[0] <= S.dt[1] ? (S.dt[0] <= S.dt[2] ? 0 : 2) : (S.dt[1] <= S.dt[2] ? 1 : 2);
    const bool a0 = a == 0, a1 = a == 1, a2 = a == 2;
    const float dta = a0 ? S.dt[0] : (a1 ? S.dt[1] : S.dt[2]);
    const float dtb = a0 ? S.dt[1] : S.dt[0], dtc = a2 ? S.dt[1] : S.dt[2];
    float ta = a0 ? S.t[0] : (a1 ? S.t[1] : S.t[2]);
    int ia = a0 ? S.c[0] : (a1 ? S.c[1] : S.c[2]);
    const int na = a0 ? P.nxg : (a1 ? P.nyg : P.nzg);
    const float da = a0 ? dx : (a1 ? dy : dz);
    const long long stra = a0 ? 1ll : (a1 ? vy : vz);
    float tb = a0 ? S.t[1] : S.t[0], tc = a2 ? S.t[1] : S.t[2];
    int ib = a0 ? S.c[1] : S.c[0], ic = a2 ? S.c[1] : S.c[2];
    const int nb = a0 ? P.nyg : P.nxg, nc = a2 ? P.nyg : P.nzg;
    const float db = a0 ? dy : dx, dc = a2 ? dy : dz;
    const long long strb = a0 ? vy : 1ll, strc = a2 ? vy : vz;
    const bool fwd = da > 0.f;
    const long long sva = fwd ? stra : -stra, svb = db > 0.f ? strb : -strb, svc = dc > 0.f ? strc : -strc;
    const int sb = db > 0.f ? 1 : -1, sc_ = dc > 0.f ? 1 : -1;
    long long* const D = a0 ? P.jdiff[0] : (a1 ? P.jdiff[1] : P.jdiff[2]);
    const float fix = (a0 ? P.jfix[0] : (a1 ? P.jfix[1] : P.jfix[2])) * weight;
    float t = 0.f;
    bool used = false;
    unsigned int nvox = 0u, nred = 0u, n_try = 0u, n_hit = 0u;
    for (;;) {
        const float tcol = fminf(tb, tc);        // the ray leaves this column (or never: BIG)
        const float tend = fminf(tcol, len);
        // faces of axis a crossed before tend: at ta + j dta, j = 0 .. m-1
        int m = 0;
        if (ta < tend) {
            m = (int)fminf((tend - ta) / dta, 1.0e9f) + 1;
            if (fmaf((float)(m - 1), dta, ta) >= tend) --m;          // rounding of the quotient, one face either way
            else if (fmaf((float)m, dta, ta) < tend) ++m;
        }
        const int room = fwd ? na - 1 - ia : ia;  // faces that can be crossed without leaving the grid
        const bool leave = m > room;
        if (leave) m = room;
        nvox += (unsigned int)m + 1u;
        if (m == 0) {
            red_f32(P.jmean + off, fmaxf((leave ? ta : tend) - t, 0.f) * weight);
            ++nred;
        } else {
            const float p_in = fmaxf(ta - t, 0.f);
            // the last voxel of a run that leaves the grid is crossed whole
            const float p_out = leave ? dta : fmaxf(tend - fmaf((float)(m - 1), dta, ta), 0.f);
            if (m < 3) {  // 2 or 3 voxels: direct deposits are no more atomics than the range update
                red_f32(P.jmean + off, p_in * weight);
                if (m == 2) red_f32(P.jmean + off + sva, dta * weight);
                red_f32(P.jmean + off + (long long)m * sva, p_out * weight);
                nred += (unsigned int)m + 1u;
            } else {
                const long long qc = __float2ll_rn(dta * fix);
                const long long q_in = __float2ll_rn(p_in * fix), q_out = __float2ll_rn(p_out * fix);
                const long long q_lo = fwd ? q_in : q_out, q_hi = fwd ? q_out : q_in;
                long long* lo = D + (fwd ? off : off - (long long)m * stra);
                const int hi_idx = fwd ? ia + m : ia;
                if (hot) {
                    n_hit += red_i64_hot(hot, lo, q_lo) ? 1u : 0u;
                    n_hit += red_i64_hot(hot, lo + stra, qc - q_lo) ? 1u : 0u;
                    n_hit += red_i64_hot(hot, lo + (long long)m * stra, q_hi - qc) ? 1u : 0u;
                    if (hi_idx + 1 < na) n_hit += red_i64_hot(hot, lo + (long long)(m + 1) * stra, -q_hi) ? 1u : 0u;
                    n_try += 4u;
                } else {
                    red_i64(lo, q_lo);
                    red_i64(lo + stra, qc - q_lo);
                    red_i64(lo + (long long)m * stra, q_hi - qc);
                    if (hi_idx + 1 < na) red_i64(lo + (long long)(m + 1) * stra, -q_hi);
                }
                nred += 4u;
                used = true;
            }
            ia += fwd ? m : -m;
            off += (long long)m * sva;
            ta = fmaf((float)m, dta, ta);
        }
        if (leave || tcol >= len) break;  // :437-440 / end of the segment
        t = tend;
        bool out;
        if (tb <= tc) { ib += sb; off += svb; tb += dtb; out = (unsigned)ib >= (unsigned)nb; }
        else          { ic += sc_; off += svc; tc += dtc; out = (unsigned)ic >= (unsigned)nc; }
        if (out) break;
    }
    if (used) P.jdiff_used[a] = 1u;
    return make_uint4(nvox, nred, n_try, n_hit);
}
// Does the run walk pay for this segment?  Only if it is at least 6 voxels long along its dominant axis and steep enough that a
// column lasts 4 voxels (faces crossed per unit length along each axis: r = |u| / vox).
__device__ __forceinline__ bool takes_run_walk(const KParams& P, float dx, float dy, float dz, float len, float& work) {
    const float rx = fabsf(dx) * P.inv_vox[0], ry = fabsf(dy) * P.inv_vox[1], rz = fabsf(dz) * P.inv_vox[2];
    const float rmax = fmaxf(rx, fmaxf(ry, rz)), rsum = rx + ry + rz;
    const float rmid = fmaxf(fminf(rx, ry), fminf(fmaxf(rx, ry), rz));
    const bool runs = !P.dda_legacy && len * rmax > 6.0f && rmax > 4.0f * rmid;
    work = 1.0f + len * (runs ? 2.0f * (rsum - rmax) : rsum);  // a segment the run walker takes only works per COLUMN change
    return runs;
}
static __device__ __forceinline__ uint2 walk_segment(const KParams& P, float fx, float fy, float fz, float dx, float dy, float dz, float len, float weight,
                                                     HotTable* hot = nullptr) {
    float work;
    if (takes_run_walk(P, dx, dy, dz, len, work)) {
        const uint4 w = walk_runs(P, fx, fy, fz, dx, dy, dz, len, weight, hot);
        return make_uint2(w.x, w.y);
    }
    return walk_voxels(P, fx, fy, fz, dx, dy, dz, len, weight);
}

// Prefix sum of one difference grid along its axis, added to jmean; the grid is cleared on the way.
//   AXIS 0 (x, contiguous): one warp per row, 32 entries per step (shuffle scan + carry)
//   AXIS 1 / 2: one thread per (x, z) / (x, y) line, neighbouring threads on neighbouring x: coalesced
template <int AXIS>
__global__ void jdiff_scan_kernel(long long* __restrict__ D, float* __restrict__ jmean, const unsigned int* used, int nx, int ny, int nz, double unit) {
    if (!used[AXIS]) return;
    if (AXIS == 0) {
        const int lane = threadIdx.x & 31;
        const long long rows = (long long)ny * nz, warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5,
                        nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long r = warp0; r < rows; r += nwarp) {
            long long carry = 0;
            for (int i0 = 0; i0 < nx; i0 += 32) {
                const int i = i0 + lane;
                long long v = i < nx ? D[r * nx + i] : 0ll;
                const bool nz_any = __any_sync(0xffffffffu, v != 0ll);
                if (nz_any && i < nx) D[r * nx + i] = 0ll;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long u = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += u;
                }
                v += carry;
                if (i < nx && v != 0ll) jmean[r * nx + i] += (float)((double)v * unit);
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
    } else {
        const long long lines = (long long)nx * (AXIS == 1 ? nz : ny), n = AXIS == 1 ? ny : nz;
        const long long stride = AXIS == 1 ? (long long)nx : (long long)nx * ny;
        for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < lines; l += (long long)gridDim.x * blockDim.x) {
            const long long base = AXIS == 1 ? (l % nx) + (l / nx) * (long long)nx * ny : l;
            long long run = 0;
            for (long long k = 0; k < n; ++k) {
                const long long at = base + k * stride;
                const long long v = D[at];
                if (v != 0ll) { D[at] = 0ll; run += v; }
                if (run != 0ll) jmean[at] += (float)((double)run * unit);
            }
        }
    }
}

// The trace kernels' side of -Dpathlength: append the segment to the CTA's share of the segment buffer; if the share is full (the
// host sizes the packets per launch from the scene's measured segments per packet, with a margin) walk it here and now.
__device__ __forceinline__ void record_segment(const KParams& P, unsigned int* seg_cnt, float sx, float sy, float sz, float ux, float uy, float uz,
                                               float px, float py, float pz, float weight, unsigned int& c_vox, unsigned int& c_red) {
    const float lx = px - sx, ly = py - sy, lz = pz - sz;
    const float l2 = lx * lx + ly * ly + lz * lz;
    if (!(l2 > 0.f)) return;
    const float len = sqrtf(l2);
    const unsigned int at = atomicAdd(seg_cnt, 1u);
    if (at < P.seg_cap) {
        float4* r = P.seg_buf + 2ull * ((unsigned long long)blockIdx.x * P.seg_cap + at);
        r[0] = make_float4(sx, sy, sz, len);
        r[1] = make_float4(ux, uy, uz, weight);
    } else {
        const uint2 w = walk_segment(P, sx, sy, sz, ux, uy, uz, len, weight);
        c_vox += w.x; c_red += w.y;
    }
}

#ifndef SMCRT_TRACE_TU  // (engine.cu only)
// The deposit kernel: walks the recorded segments.  One warp takes 32 records of a CTA's share at a time (two coalesced 16-byte
// loads per lane) and sorts them by the work they are:
//   * a segment that stays inside ONE voxel -- half the free paths of a turbid medium -- is one atomic, issued at once;
//   * a short segment for the voxel walk, and one for the run walker (the beam's first flight, a long free path along an axis),
//     go to the warp's two queues in shared memory and are walked when 32 of a kind have gathered, one per lane: every lane of
//     the walk is busy.  (Walked where they turned up, the run walk ran on 1.3 lanes of 32 and was 28 % of the kernel's
//     instructions on skin_b200.toml; the voxel walk's set-up, ~130 instructions, ran for whatever lanes were not single-voxel.)
//   * a long segment (the 300-voxel flight of an escaping packet, a refracted ray through the 200^3 grid of sphere.toml) is cut
//     into pieces of seg_piece voxel visits that are dealt out to the lanes.
// A small kernel with a small loop: the voxel walk does not compete with the transport code for the instruction cache (in one
// kernel the hot code was 39 KB, beyond the 32 KB L1.5 I-cache: 6 stall cycles per issue waiting for instructions).
//
// A warp's queue: 32 records (2 x float4).  Lanes with `pred` append theirs; when that makes 32, the full queue is taken (one
// record per lane, returned in qa / qb with `full`), and the records that did not fit start the next filling.
__device__ __forceinline__ bool warp_queue_push(float4* q, int& n, bool pred, const float4& a, const float4& b, int lane, float4& qa, float4& qb) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return false;
    const int cnt = __popc(m), rank = __popc(m & ((1u << lane) - 1u)), room = 32 - n;
    if (pred && rank < room) { q[2 * (n + rank)] = a; q[2 * (n + rank) + 1] = b; }
    __syncwarp();
    if (n + cnt < 32) { n += cnt; return false; }
    qa = q[2 * lane]; qb = q[2 * lane + 1];
    __syncwarp();
    if (pred && rank >= room) { q[2 * (rank - room)] = a; q[2 * (rank - room) + 1] = b; }
    n += cnt - 32;
    __syncwarp();
    return true;
}
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) deposit_segments_kernel(const __grid_constant__ KParams P, int n_shares) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned int c_vox = 0u, c_red = 0u;  // (per thread: < 2^32 in any launch)
    extern __shared__ __align__(16) unsigned char dsm[];
    HotTable* hot = reinterpret_cast<HotTable*>(dsm);
    for (int i = threadIdx.x; i < HOT_SLOTS; i += blockDim.x) { hot->key[i] = 0ull; hot->lo[i] = 0u; hot->hi[i] = 0u; }
    __syncthreads();
    float4* const runq = reinterpret_cast<float4*>(dsm + sizeof(HotTable)) + warp * (4 * WARPQ_CAP);  // the warp's two queues
    float4* const voxq = runq + 2 * WARPQ_CAP;
    int n_runq = 0, n_voxq = 0;  // (warp-uniform)
    bool use_hot = true;         // (warp-uniform) the table still takes a fair share of what this warp offers it
    unsigned int hot_try = 0u, hot_hit = 0u;
    // work items = (share, eighth of the share), handed out by a global counter: the CTAs of this launch are persistent (as many as
    // fit beside their 32-KB tables) and stay busy until the last record.  (Handed out to single warps, a 64th of a share at a
    // time and no barrier, the kernel was 1-3 % slower: the warps of a CTA no longer read neighbouring records.)
    __shared__ unsigned int item_s;
    constexpr unsigned int PARTS = 8u;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item_s = atomicAdd(P.seg_work, 1u);
        __syncthreads();
        const unsigned int item = item_s;
        if (item >= (unsigned int)n_shares * PARTS) break;
        const int sh = (int)(item / PARTS);
        const unsigned int part = item % PARTS;
        const unsigned int n_all = min(P.seg_count[sh], P.seg_cap);
        const unsigned int r0 = (unsigned int)(((unsigned long long)n_all * part / PARTS + 31ull) & ~31ull),
                           r1 = part + 1u == PARTS ? n_all : (unsigned int)(((unsigned long long)n_all * (part + 1u) / PARTS + 31ull) & ~31ull);
        const unsigned int n = min(r1, n_all);
        const float4* recs = P.seg_buf + 2ull * (unsigned long long)sh * P.seg_cap;
        for (unsigned int base = r0 + warp * 32u; base < n; base += nwarp * 32u) {
            const unsigned int i = base + lane;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 1.f, 0.f);
            float work = 0.f;
            bool runs = false;
            if (i < n) {
                a = recs[2u * i]; b = recs[2u * i + 1u];
                runs = takes_run_walk(P, b.x, b.y, b.z, a.w, work);
            }
            // ---- inside one voxel?  Start and end in voxel units, both well inside the same cell (the margin covers the rounding
            // of these few operations: the walk's own cell and face arithmetic is not needed to know that no face is crossed).
            // The walk would deposit len * weight into the start cell; so does this.
            if (work > 0.f && work < 4.0f && !P.dda_legacy) {
                const float s0 = fmaf(a.x, P.inv_vox[0], 0.5f * (float)P.nxg), s1 = fmaf(a.y, P.inv_vox[1], 0.5f * (float)P.nyg),
                            s2 = fmaf(a.z, P.inv_vox[2], 0.5f * (float)P.nzg);
                const float e0 = fmaf(a.w * b.x, P.inv_vox[0], s0), e1 = fmaf(a.w * b.y, P.inv_vox[1], s1), e2 = fmaf(a.w * b.z, P.inv_vox[2], s2);
                const float c0 = floorf(s0), c1 = floorf(s1), c2 = floorf(s2);
                constexpr float MG = 2.0e-3f;
                const float lo = fminf(fminf(fminf(s0 - c0, e0 - c0), fminf(s1 - c1, e1 - c1)), fminf(s2 - c2, e2 - c2));
                const float hi = fmaxf(fmaxf(fmaxf(s0 - c0, e0 - c0), fmaxf(s1 - c1, e1 - c1)), fmaxf(s2 - c2, e2 - c2));
                if (lo > MG && hi < 1.0f - MG && c0 >= 0.f && c1 >= 0.f && c2 >= 0.f && c0 < (float)P.nxg && c1 < (float)P.nyg && c2 < (float)P.nzg) {
                    red_f32(P.jmean + ((long long)c0 + (long long)P.nxg * ((long long)c1 + (long long)P.nyg * (long long)c2)), a.w * b.w);
                    ++c_vox; ++c_red;
                    work = 0.f;  // done
                }
            }
            // ---- long segments: cut into PIECES of about seg_piece voxel visits, numbered through the batch (a warp prefix sum);
            // lane j of pass p walks piece 32 p + j of whichever segment it belongs to (its record fetched by shuffle): equal work
            // per lane.  A voxel that holds a cut gets its length in two deposits.
            const bool brief = work > 0.f && work <= P.seg_piece;
            const int np = (work > 0.f && !brief) ? (int)fminf(ceilf(work / P.seg_piece), 65536.f) : 0;
            int incl = np;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int g0 = 0; g0 < total; g0 += 32) {
                const int g = min(g0 + lane, total - 1);
                int own = 0;  // the segment piece g belongs to: the number of lanes whose pieces all come before it
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, own + st - 1);
                    if (v <= g) own += st;
                }
                const int npo = __shfl_sync(0xffffffffu, np, own), first = __shfl_sync(0xffffffffu, incl, own) - npo;
                const float L = __shfl_sync(0xffffffffu, a.w, own), w = __shfl_sync(0xffffffffu, b.w, own);
                const float ax = __shfl_sync(0xffffffffu, a.x, own), ay = __shfl_sync(0xffffffffu, a.y, own), az = __shfl_sync(0xffffffffu, a.z, own);
                const float vx = __shfl_sync(0xffffffffu, b.x, own), vy = __shfl_sync(0xffffffffu, b.y, own), vz = __shfl_sync(0xffffffffu, b.z, own);
                if (g0 + lane < total) {
                    const int q = g - first;
                    const float inv = 1.0f / (float)npo;
                    const float t0 = L * ((float)q * inv), t1 = q + 1 == npo ? L : L * ((float)(q + 1) * inv);
                    const uint2 wk = walk_segment(P, fmaf(t0, vx, ax), fmaf(t0, vy, ay), fmaf(t0, vz, az), vx, vy, vz, t1 - t0, w, use_hot ? hot : nullptr);
                    c_vox += wk.x; c_red += wk.y;
                }
            }
            // ---- short segments: into the queue of their walker; a full queue is walked, one record per lane
            float4 qa, qb;
            if (warp_queue_push(voxq, n_voxq, brief && !runs, a, b, lane, qa, qb)) {
                const uint2 w = walk_voxels_inl(P, qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, qa.w, qb.w);
                c_vox += w.x; c_red += w.y;
            }
            if (warp_queue_push(runq, n_runq, brief && runs, a, b, lane, qa, qb)) {
                const uint4 w = walk_runs(P, qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, qa.w, qb.w, use_hot ? hot : nullptr);
                c_vox += w.x; c_red += w.y;
                // Does the table earn its probes?  On a pencil beam nearly every entry is one of the ~1300 it holds; on a diffuse scene
                // (sphere.toml) it fills with entries that never come back and every update pays three probes for nothing -- a third
                // of this kernel's instructions (profiles/r02_sphere_deposit_kernel.txt).  The warp counts, and stops offering.
                if (use_hot) {
                    hot_try += __reduce_add_sync(0xffffffffu, w.z); hot_hit += __reduce_add_sync(0xffffffffu, w.w);
                    if (hot_try >= 1024u) {
                        if (4u * hot_hit < hot_try) use_hot = false;
                        hot_try = 0u; hot_hit = 0u;
                    }
                }
            }
        }
    }
    if (lane < n_voxq) {  // what is left in the warp's queues
        const float4 qa = voxq[2 * lane], qb = voxq[2 * lane + 1];
        const uint2 w = walk_voxels(P, qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, qa.w, qb.w);
        c_vox += w.x; c_red += w.y;
    }
    if (lane < n_runq) {
        const float4 qa = runq[2 * lane], qb = runq[2 * lane + 1];
        const uint4 w = walk_runs(P, qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, qa.w, qb.w, use_hot ? hot : nullptr);
        c_vox += w.x; c_red += w.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HOT_SLOTS; i += blockDim.x)  // flush the table: one L2 atomic per entry held
        if (hot->key[i]) {
            const long long v = (long long)(((unsigned long long)hot->hi[i] << 32) | (unsigned long long)hot->lo[i]);
            if (v) red_i64(reinterpret_cast<long long*>(hot->key[i]), v);
        }
    unsigned long long t_vox = c_vox, t_red = c_red;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t_vox += __shfl_xor_sync(0xffffffffu, t_vox, o); t_red += __shfl_xor_sync(0xffffffffu, t_red, o); }
    if (lane == 0 && t_vox) { atomicAdd(&P.counters[C_VOXELS], t_vox); atomicAdd(&P.counters[C_REDS], t_red); }
}
// (the shares are cleared for the next launch by a second tiny kernel: a CTA of the deposit kernel may still be reading a count)
__global__ void clear_segment_counts_kernel(unsigned int* cnt, int n, unsigned int* work) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cnt[i] = 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) *work = 0u;
}

#endif  // SMCRT_TRACE_TU

// ------------------------------------------------------------------------------------------------ detectors
// One straight segment (start, dir, length) against one detector: record_hit_1D/2D + check_hit_*
// (src/detectors/detector_base.f90:137-163,206-235; src/detectors/detectors.f90:147-469;
//  intersectPlane/Circle src/geometryMod.f90:217-270).  Returns the 1-based flat bin, 0 on miss.
// FP32 note: the reference accepts a hit when 0 < t <= pointSep with t = ((p0-l0).n)/(n.l).  In FP32
// the rounding of t against pointSep (both ~1e-7 relative) is comparable to the sub-eps gaps between consecutive
// segments, so crossings would be dropped or double counted.  The same condition is therefore evaluated as a sign
// change of the plane side function g(x) = (p0 - x).n between the segment's two stored end points: a crossing is
// recorded iff g(start) >= 0 and g(end) < 0.  Consecutive segments share their end point bit-for-bit, so every
// crossing is seen exactly once ("watertight"), which is what the reference's FP64 arithmetic achieves implicitly.
__device__ __forceinline__ int nint_pos(float v) { return (int)floorf(v + 0.5f); }  // Fortran nint for v >= 0
// Plane detectors (circle, annulus, fibre): crossing test + radial distance of the hit point from the detector centre.
// The side function is evaluated in plane form g(x) = n.p0 - n.x from ONE 16-byte record (pl = (n, n.p0)): a segment that does
// not cross -- nearly all of them -- costs one shared load, six FMAs and two compares per detector.  Any fixed g keeps the test
// watertight: consecutive segments share their end point bit for bit, so g(end of k) IS g(start of k+1).
// `denom` = n.l is returned for the fibre's acceptance chain.
__device__ __forceinline__ bool det_plane_cross(const float4 pl, float s0, float s1, float s2, float e0, float e1, float e2, float& gs) {
    gs = pl.w - (pl.x * s0 + pl.y * s1 + pl.z * s2);
    const float ge = pl.w - (pl.x * e0 + pl.y * e1 + pl.z * e2);
    return gs >= 0.f && ge < 0.f;
}
__device__ __forceinline__ bool det_plane_hit(const DevDet& D, const float4 pl, float gs, float s0, float s1, float s2, float d0, float d1, float d2,
                                              float& r, float& denom) {
    denom = pl.x * d0 + pl.y * d1 + pl.z * d2;
    if (!(denom > 1e-6f)) return false;  // intersectPlane: src/geometryMod.f90:234
    const float t = gs / denom;
    const float vx = s0 + d0 * t - D.pos[0], vy = s1 + d1 * t - D.pos[1], vz = s2 + d2 * t - D.pos[2];
    r = sqrtf(vx * vx + vy * vy + vz * vz);
    return true;
}
__device__ __forceinline__ int det_bin_circle(const DevDet& D, float r) {  // :147-164
    return r <= D.q[0] ? min(nint_pos(r / D.q[1]) + 1, D.nbins) : 0;
}
static __device__ __noinline__ int det_bin_annulus_fibre(const DevDet* Dp, float r, float denom) {
    const DevDet& D = *Dp;
    if (D.kind == 2) {  // annulus :212-244: not inside r1, inside r2
        if (r <= D.q[0] || !(r <= D.q[1])) return 0;
        return max(min(nint_pos((r - D.q[0]) / D.q[2]) + 1, D.nbins), 1);
    }
    // fibre: 4f relay in the thin-lens approximation :323-393
    if (!(r <= D.q[0])) return 0;
    const float costt = fminf(denom, 1.0f);
    const float sintt = sqrtf(1.0f - costt * costt);
    float gradient = sintt / costt;
    float radius = r;
    gradient = -radius / D.q[1] + gradient;
    radius = radius + gradient * D.q[3];
    if (radius > D.q[4]) return 0;
    radius = radius + gradient * D.q[5];
    if (radius > D.q[6]) return 0;
    gradient = -radius / D.q[2] + gradient;
    radius = radius + gradient * D.q[7];
    const float angle = fabsf(atanf(gradient)) * (360.0f / TWOPI_F);
    if (angle > D.q[8] || radius > D.q[9]) return 0;
    return min(nint_pos(fabsf(radius) / D.q[10]) + 1, D.nbins);
}
// camera :447-469 + record_hit_2D_sub (no pointSep test, bins the segment START, adds 1)
static __device__ __noinline__ int det_bin_camera(const DevDet* Dp, float s0, float s1, float s2, float d0, float d1, float d2) {
    const DevDet& D = *Dp;
    const float dn = d0 * D.dir[0] + d1 * D.dir[1] + d2 * D.dir[2];
    const float tt = ((D.pos[0] - s0) * D.dir[0] + (D.pos[1] - s1) * D.dir[1] + (D.pos[2] - s2) * D.dir[2]) / dn;
    if (!(tt >= 0.f)) return 0;
    const float vx = s0 + tt * d0 - D.pos[0], vy = s1 + tt * d1 - D.pos[1], vz = s2 + tt * d2 - D.pos[2];
    const float p1 = (vx * D.q[0] + vy * D.q[1] + vz * D.q[2]) / D.q[6];
    const float p2 = (vx * D.q[3] + vy * D.q[4] + vz * D.q[5]) / D.q[7];
    if (!(p1 < D.q[6] && p1 > 0.f && p2 < D.q[7] && p2 > 0.f)) return 0;
    const float bx = s2 + D.q[10], by = s1 + D.q[11];  // sic: hitpoint%pos%z + this%pos%x
    int ix = min((int)(bx / D.q[8]) + 1, D.nbins), iy = min((int)(by / D.q[9]) + 1, D.nbins);
    if (ix < 1) ix = D.nbins;
    if (iy < 1) iy = D.nbins;
    return ix + (iy - 1) * D.nbins;
}
// One straight segment against one detector -> 1-based flat bin, 0 on miss (the kernel's DETECT site and the probe kernel)
__device__ __forceinline__ int detector_bin(const DevDet* Dp, const float4 pl, float s0, float s1, float s2, float d0, float d1, float d2, float e0,
                                            float e1, float e2) {
    if (Dp->kind == 4) return det_bin_camera(Dp, s0, s1, s2, d0, d1, d2);
    float r, denom, gs;
    if (!det_plane_cross(pl, s0, s1, s2, e0, e1, e2, gs)) return 0;
    if (!det_plane_hit(*Dp, pl, gs, s0, s1, s2, d0, d1, d2, r, denom)) return 0;
    return Dp->kind == 1 ? det_bin_circle(*Dp, r) : det_bin_annulus_fibre(Dp, r, denom);
}

// ------------------------------------------------------------------------------------------------ emitters
// src/photon.f90:214-1043.  u0..u2: uniforms of the event block.  Returns false when the emitter's own rejection
// step (gaussian annulus, `rang` src/random_mod.f90:99-124) wants a fresh block.
__device__ __forceinline__ void nudge_face(float& c, float cmax) {
    // 7.9e-7 inset when exactly on a grid face (photon.f90:614-628); in FP32 the inset must survive rounding
    const float inset = fmaxf(7.9e-7f, 4.0f * 1.1920929e-7f * cmax);
    if (c == -cmax) c += inset;
    else if (c == cmax) c -= inset;
}
__device__ __forceinline__ void clip_to_grid(const KParams& P, float pos[3], const float dir[3], int cap) {
    // photon.f90:495-553 / :988-1036: project onto the grid box along the ray, at most `cap`+2 passes
    bool in[3] = {false, false, false}, tried[3] = {false, false, false};
    int counter = 0;
    while (!in[0] || !in[1] || !in[2]) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float inset = fmaxf(9e-7f, 4.0f * 1.1920929e-7f * P.gmax[a]);
            if (pos[a] <= -P.gmax[a]) {
                const float st = (-P.gmax[a] - pos[a] + inset) / dir[a];
                pos[0] += dir[0] * st; pos[1] += dir[1] * st; pos[2] += dir[2] * st;
                tried[a] = true;
            } else if (pos[a] >= P.gmax[a]) {
                const float st = (P.gmax[a] - pos[a] - inset) / dir[a];
                pos[0] += dir[0] * st; pos[1] += dir[1] * st; pos[2] += dir[2] * st;
                tried[a] = true;
            } else
                in[a] = true;
        }
        if ((tried[0] && tried[1] && tried[2]) || counter > cap) break;
        ++counter;
    }
}
__device__ __forceinline__ void xform_pos(const float T[12], const float l[3], float o[3]) {
    o[0] = T[0] * l[0] + T[1] * l[1] + T[2] * l[2] + T[3];
    o[1] = T[4] * l[0] + T[5] * l[1] + T[6] * l[2] + T[7];
    o[2] = T[8] * l[0] + T[9] * l[1] + T[10] * l[2] + T[11];
}
__device__ __forceinline__ bool emit_packet_v(const KParams& P, float u0, float u1, float u2, float u3, float u4, float pos[3], float dir[3]);
struct Emitted {
    float x, y, z, dx, dy, dz;
    bool ok;
};
// (pid, event): dslit / aperture draw five / four uniforms; the two beyond the event's block come from its SECOND Philox block
// (counter word 3 = 1), words 0 and 1
static __device__ __noinline__ Emitted emit_packet(const KParams& P, float u0, float u1, float u2, unsigned long long pid, uint32_t event) {  // out of line: all source kinds live here
    float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f};
    Emitted e;
    float u3 = 0.f, u4 = 0.f;
    if (P.src_kind >= 7) {
        uint32_t x[4];
        philox4x32_10(event, (uint32_t)pid, (uint32_t)(pid >> 32), 1u, P.seed_lo, P.seed_hi, x);
        u3 = u01(x[0]); u4 = u01(x[1]);
    }
    e.ok = emit_packet_v(P, u0, u1, u2, u3, u4, pos, dir);
    e.x = pos[0]; e.y = pos[1]; e.z = pos[2]; e.dx = dir[0]; e.dy = dir[1]; e.dz = dir[2];
    return e;
}
__device__ __forceinline__ bool emit_packet_v(const KParams& P, float u0, float u1, float u2, float u3, float u4, float pos[3], float dir[3]) {
    const float* sp = P.sp;
    switch (P.src_kind) {
        case 7:    // dslit :712-780 (draws: slit pick, x1, y1, x2, y2)
        case 8: {  // aperture :782-848 (draws: x1, y1, x2, y2)
            // hard-coded geometry in units of the wavelength (sp[15]); the source point (x1, y1, z1) is 1e3..1e7 grid units away, so
            // the direction is formed in FP64: (x2 - x1) / |.| with |x2 - x1| << |z2 - z1| would lose its digits in FP32
            const double lam = (double)sp[15];
            double x1, y1, z1, x2, y2, z2;
            if (P.src_kind == 7) {
                const double a = 60.0 * lam, b = 20.0 * lam;
                x1 = u0 > 0.5f ? a * 0.5 + (double)u1 * b : -a * 0.5 - (double)u1 * b;   // ranu(lo, hi) = lo + xi (hi - lo), hi < lo allowed
                y1 = -b * 0.5 + (double)u2 * b;
                z2 = 5.0 - (1.e-5 * (2.0 * (5.0 / 400.0)));
                x2 = -5.0 + 10.0 * (double)u3;
                y2 = -5.0 + 10.0 * (double)u4;
                z1 = 10000.0 * lam - 5.0;
            } else {
                const double apwid = 200e-6, b = apwid * 0.5, F = 4.95;
                x1 = -b + (double)u0 * apwid;
                y1 = -b + (double)u1 * apwid;
                z1 = 1.0 / ((((F / apwid) * (F / apwid)) * 0.5) * lam) - 0.5;
                x2 = -0.5 + (double)u2;
                y2 = -0.5 + (double)u3;
                z2 = 0.5 - (1.e-5 * (2.0 * 0.5 / 400.0));
            }
            const double ex = x2 - x1, ey = y2 - y1, ez = z2 - z1;
            const double il = rsqrt(ex * ex + ey * ey + ez * ez);
            pos[0] = (float)x2; pos[1] = (float)y2; pos[2] = (float)z2;
            // z2 = zmax - 2.5e-7 (resp. - 2.5e-8) rounds ONTO the grid face in FP32: same inset as the other emitters
            nudge_face(pos[0], P.gmax[0]); nudge_face(pos[1], P.gmax[1]); nudge_face(pos[2], P.gmax[2]);
            dir[0] = (float)(ex * il); dir[1] = (float)(ey * il); dir[2] = (float)(-fabs(ez) * il);
            const float nl = rsqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);  // unit in FP32 as well
            dir[0] *= nl; dir[1] *= nl; dir[2] *= nl;
            return true;
        }
        case 1: {  // point :311-359
            pos[0] = sp[0]; pos[1] = sp[1]; pos[2] = sp[2];
            float sinp, cosp;
            sincospif(2.0f * u0, &sinp, &cosp);
            const float cost = 2.0f * u1 - 1.0f;
            const float sint = sqrtf(1.0f - cost * cost);
            dir[0] = sint * cosp; dir[1] = sint * sinp; dir[2] = cost;
            return true;
        }
        case 2: {  // pencil :652-710
            pos[0] = sp[0]; pos[1] = sp[1]; pos[2] = sp[2];
            nudge_face(pos[0], P.gmax[0]); nudge_face(pos[1], P.gmax[1]); nudge_face(pos[2], P.gmax[2]);
            dir[0] = sp[3]; dir[1] = sp[4]; dir[2] = sp[5];
            return true;
        }
        case 3: {  // uniform :566-649
            pos[0] = sp[6] + u0 * sp[9] + u1 * sp[12];
            pos[1] = sp[7] + u0 * sp[10] + u1 * sp[13];
            pos[2] = sp[8] + u0 * sp[11] + u1 * sp[14];
            nudge_face(pos[0], P.gmax[0]); nudge_face(pos[1], P.gmax[1]); nudge_face(pos[2], P.gmax[2]);
            dir[0] = sp[3]; dir[1] = sp[4]; dir[2] = sp[5];
            return true;
        }
        case 4: {  // circular :214-308
            const float r = sp[15] * sqrtf(u0);
            float s, c;
            sincospif(2.0f * u1, &s, &c);
            float l[3];
            if (P.src_alt) { l[0] = r * c; l[1] = r * s; l[2] = 0.f; }
            else           { l[0] = 0.f;   l[1] = r * c; l[2] = r * s; }
            float w[3];
            xform_pos(P.Tpos, l, w);
            pos[0] = -w[0]; pos[1] = -w[1]; pos[2] = -w[2];
            nudge_face(pos[0], P.gmax[0]); nudge_face(pos[1], P.gmax[1]); nudge_face(pos[2], P.gmax[2]);
            dir[0] = sp[3]; dir[1] = sp[4]; dir[2] = sp[5];
            return true;
        }
        case 5:    // focus :361-563
        case 6: {  // annulus :850-1043
            float l[3] = {0.f, 0.f, 0.f}, a[3];  // local position, local aim point
            const float focal = sp[16];
            if (P.src_kind == 5) {
                const float beam = sp[17];
                if (P.src_sub == 1) {  // square
                    l[0] = -beam + u0 * (2.0f * beam);
                    l[1] = -beam + u1 * (2.0f * beam);
                } else {
                    const float rad = P.src_sub == 2 ? beam * sqrtf(u0) : beam * sqrtf(-logf(1.0f - u0));
                    float s, c;
                    sincospif(2.0f * u1, &s, &c);
                    l[0] = rad * c; l[1] = rad * s;
                }
                a[0] = l[0]; a[1] = l[1]; a[2] = 0.f;
            } else {
                const float rlo = sp[18], rhi = sp[19], mid = 0.5f * (rhi + rlo);
                float rad;
                if (P.src_sub == 1) rad = sqrtf(rlo * rlo + (rhi * rhi - rlo * rlo) * u0);
                else if (P.src_sub == 2) rad = rlo + (rhi - rlo) * u0;
                else {
                    const float gx = -1.0f + 2.0f * u0, gy = -1.0f + 2.0f * u1;
                    const float sq = gx * gx + gy * gy;
                    if (sq >= 1.0f || sq == 0.0f) return false;
                    rad = mid + sp[20] * (gx * sqrtf(-2.0f * logf(sq) / sq));
                }
                float s, c;
                sincospif(2.0f * u2, &s, &c);
                l[0] = rad * c; l[1] = rad * s;
                a[0] = mid * c; a[1] = mid * s; a[2] = 0.f;  // all rays aim from the ring-mid radius
            }
            // dir = -(a - targ)/|a - targ| * sign(1, focal), targ = (0,0,-focal)
            float dl[3] = {-a[0], -a[1], -(a[2] + focal)};
            const float sgn = focal >= 0.f ? 1.0f : -1.0f;
            float il = sgn * rsqrtf(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
            dl[0] *= il; dl[1] *= il; dl[2] *= il;
            float dw[3] = {P.Tdir[0] * dl[0] + P.Tdir[1] * dl[1] + P.Tdir[2] * dl[2],
                           P.Tdir[3] * dl[0] + P.Tdir[4] * dl[1] + P.Tdir[5] * dl[2],
                           P.Tdir[6] * dl[0] + P.Tdir[7] * dl[1] + P.Tdir[8] * dl[2]};
            il = rsqrtf(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
            dir[0] = dw[0] * il; dir[1] = dw[1] * il; dir[2] = dw[2] * il;
            xform_pos(P.Tpos, l, pos);
            clip_to_grid(P, pos, dir, P.src_kind == 5 ? 4 : 3);
            return true;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ the kernel
// Per-lane state machine.  Hot states own one sweep (evaluation of ALL SDFs at one point) per loop iteration:
//   ST_MARCH     sweep at pos: the reference's top-of-loop / re-evaluation / sphere-trace evaluations (inttau2.f90:63,134,179)
//   ST_BND_PROBE sweep at pos + (d+2eps) dir: on-boundary nudge probe (:77-84)
//   ST_CROSS     sweep at pos + dstep dir: boundary-crossing probe and its creep (:213-235)
// Cold states are resolved at the top of an iteration; each consumes exactly one Philox block, generated at ONE site:
//   ST_FRESNEL   index-mismatch crossing (:248-317)        ST_INTERACT  scatter / absorb (kernelsMod.f90:1958-1974)
//   ST_EMIT      launch a new packet (kernelsMod.f90:1937-1952)
enum : int { ST_MARCH = 0, ST_BND_PROBE, ST_CROSS, ST_FRESNEL, ST_INTERACT, ST_EMIT, ST_DONE, ST_HOLD /* within one iteration: waiting for FINISH */ };
enum : int { FATE_ABSORBED = 0, FATE_ESCAPED = 1, FATE_ROULETTE = 2, FATE_LOST = 3 };
enum : int { POST_NONE = 0, POST_FINISH, POST_AFTER_TRACE, POST_NEXT_LOOP };
enum : int { LOST_STEPS = 1, LOST_NO_SURFACE = 2, LOST_BOUNCES = 3, LOST_NO_LAYER = 4, LOST_EMIT = 5 };

// tau = -ln(xi): MUFU.LG2 * ln2 (abs. error 2^-21.4 near 1, <= 2 ulp elsewhere): a 4e-7 perturbation of a free path.
#ifdef SMCRT_PRECISE_LOG
#define SMCRT_LOG logf
#else
#define SMCRT_LOG __logf
#endif
// optional features of a run (per-packet records, diagnostics, batched sources, survival biasing): tested through SMCRT_OPT so
// that the LEAN kernels -- what a plain smcrt_run of a sphere/box scene uses -- do not carry their branches (5 % on the slab)
#define SMCRT_OPT(X) (!LEAN && (X))
#ifndef SMCRT_BLOCK
#define SMCRT_BLOCK 256      // threads per CTA
#endif
#ifndef SMCRT_MINBLOCKS
#define SMCRT_MINBLOCKS 3    // resident CTAs per SM the register allocation is tuned for (80 registers; measured +5 % over 2)
#endif
// optional per-packet record of smcrt_trace_packets (out of line: cold)
// index of a packet's per-packet records: its offset in the run, or its position in the (sorted) id list of a replay
__device__ __forceinline__ long long rec_index(const KParams& P, unsigned long long pid) {
    if (!P.id_list) return (long long)(pid - P.rec_id0);
    long long lo = 0, hi = P.id_list_n - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (P.id_list[mid] < pid) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// trackHistory: history%push(vec4(packet%pos, packet%step)) (kernelsMod.f90:1954,1959) for a replayed packet
static __device__ __noinline__ void push_vertex(const KParams& P, unsigned long long pid, float x, float y, float z, float step) {
    const long long k = rec_index(P, pid);
    const int n = P.out_nvert[k];
    if (n < P.max_vert) P.out_vert[k * (long long)P.max_vert + n] = make_float4(x, y, z, step);
    P.out_nvert[k] = n + 1;  // (beyond max_vert: counted, not stored)
}
static __device__ __noinline__ void record_packet(const KParams& P, unsigned long long pid, int fate, int why, uint32_t ev, int steps,
                                           float x, float y, float z) {
    const long long k = rec_index(P, pid);
    P.out_fate[k] = fate;
    if (P.out_events) P.out_events[k] = fate == 3 ? -why : (int)ev;
    if (P.out_sweeps) P.out_sweeps[k] = steps;
    if (P.out_pos) { P.out_pos[3 * k] = x; P.out_pos[3 * k + 1] = y; P.out_pos[3 * k + 2] = z; }
}

constexpr int XCHG_WORDS = 22;  // 32-bit words of packet state exchanged by the compaction step (at most)

// MINBLOCKS = resident CTAs per SM the register allocation is made for (2: 128 registers, no spills; 3: 80; 4: 64).  Which one
// wins depends on the scene (slab with detectors: 4, +8 %; long histories inside one body: 2, +17 %), so the engine times
// the three on the first large run of a scene and keeps the fastest (engine.cu: run_on_device).
// NEED: the sweep is told how far the packet can still travel (capsule scenes, clear cells of the culling grid): compiled in only
// for the scenes that use it -- carrying the value through the sweep costs 4 % on the ones that do not.
template <bool PATHLEN, bool HASDET, bool COMPACT, int MINBLOCKS, bool NEED>
__global__ void __launch_bounds__(SMCRT_BLOCK, (MINBLOCKS * 256) / SMCRT_BLOCK) trace_persistent(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    {  // stage the scene in shared memory (16-byte vector copies)
        const int4* src = reinterpret_cast<const int4*>(P.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = threadIdx.x; i < P.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    unsigned long long* sbins = reinterpret_cast<unsigned long long*>(smem + P.blob_bytes);
    unsigned int* seg_cnt = reinterpret_cast<unsigned int*>(smem + P.seg_off);  // PATHLEN: segments recorded by this CTA
    if (PATHLEN && threadIdx.x == 0) *seg_cnt = 0u;
    // compaction scratch (COMPACT only): per-state totals (double buffered) and one slot of XCHG_WORDS words per thread
    uint32_t* xtot = reinterpret_cast<uint32_t*>(smem + P.xchg_off);
    uint32_t* xbuf = xtot + 16;
    if (COMPACT && threadIdx.x < 16) xtot[threadIdx.x] = 0u;
    uint32_t xiter = 0;
    constexpr bool SIMPLE = false;  // (sphere/box-only specialisation: queued kernels only)
    constexpr bool LEAN = false;
    // COMPACT only: xiter == XTAIL = compaction switched off for the rest of the run (CTA-uniform; no register of its own)
    constexpr uint32_t XTAIL = 0xffffffffu;
    if (HASDET && P.det_in_smem)
        for (int i = threadIdx.x; i < P.det_total; i += blockDim.x) sbins[i] = 0ull;
    __shared__ FirstFlight ff;
    if (threadIdx.x == 0) {
        ff.ok = 0;
        if (P.ff) ff = *P.ff;
    }
    __syncthreads();
    const SceneView sc = make_view(smem, P);
    const int lane = threadIdx.x & 31;
    if (P.tstamp && blockIdx.x == 0 && threadIdx.x == 0) P.tstamp[0] = globaltimer_ns();

    // ---- packet state (scalars only: nothing here may be address-taken, or it lands in local memory)
    // Position in FP64, mirrored to FP32 for everything evaluated per sweep: with an FP32 position a step smaller than
    // half an ulp of a coordinate is lost and rays grazing a wall stop converging.
    double pxd = 0, pyd = 0, pzd = 0;
    float px = 0, py = 0, pz = 0, ux = 0, uy = 0, uz = 1, sx = 0, sy = 0, sz = 0;  // position, direction, segment start
    float tau = 0.f, taurun = 0.f, dstep = 0.f, qs = 0.f, dlast = 0.f, weight = 1.f;
    int layer = 0, new_layer = 0, state = ST_EMIT, bounces = 0, steps = 0;
    bool tflag = false, launch = false, have_pid = false;
    int phase = 0;  // of ST_MARCH: 0 = top of the tauint2 loop, 1 = re-evaluation after a boundary nudge, 2 = inside the sphere-trace loop
    unsigned long long pid = 0;
    uint32_t ev = 0;
    // per-thread event counters (< 2^32 each); the rare ones (bounces, emit retries, lost) go straight to the global counters
    unsigned int c_nscatt = 0, c_sweeps = 0, c_dethits = 0, c_vox = 0, c_red = 0;

#include "step_macros.inc"
    for (;;) {
        // every lane of the warp executes this vote in every iteration (lanes leave the loop together or not at all): full mask, so
        // that a warp the compiler has not reconverged here cannot lose a subset of its lanes
#define STEP_EXIT_CHECK if ((!COMPACT || xiter == XTAIL) && __all_sync(0xffffffffu, state == ST_DONE)) break;
#include "step_body.inc"
#undef STEP_EXIT_CHECK

        if (COMPACT && xiter != XTAIL) {
            // ============================ event compaction ============================
            // Counting sort of the CTA's packets by state through shared memory: afterwards the lanes of a warp are (mostly)
            // in the same state, so the divergent cold blocks / transitions run with full warps.  Order inside a bucket is
            // irrelevant.  Two barriers per iteration; the per-state totals are double buffered so that zeroing never races.
            const unsigned full = 0xffffffffu;
            uint32_t* tot = xtot + 8 * (xiter & 1u);
            const unsigned same = __match_any_sync(full, state);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int lead = __ffs(same) - 1;
            uint32_t woff = 0;
            if (lane == lead) woff = atomicAdd(&tot[state], (uint32_t)__popc(same));
            woff = __shfl_sync(full, woff, lead);
            __syncthreads();  // (A) totals complete; everybody has consumed the previous exchange
            if (threadIdx.x < 8) xtot[8 * ((xiter + 1u) & 1u) + threadIdx.x] = 0u;
            uint32_t base = 0;
#pragma unroll
            for (int k = 0; k < ST_DONE; ++k) base += (k < state) ? tot[k] : 0u;
            const bool all_done = tot[ST_DONE] == (uint32_t)blockDim.x;
            // The pool is empty and half of the CTA has nothing left to do: this is the last exchange.  The survivors end up packed
            // in the first warps and finish without the two barriers per iteration, each warp leaving on its own; the run's tail
            // -- the longest history, alone -- advances at one warp's latency instead of the CTA's.
            const bool go_tail = 2u * tot[ST_DONE] >= (uint32_t)blockDim.x;
            uint32_t* w = xbuf + (base + woff + (uint32_t)rank);
            const int B = blockDim.x;
            // packet state: 18 words + 3 (segment start, detector scenes) + 1 (weight, survival biasing).  qs is implied by the
            // state, the event index shares a word with the flags (ev <= 100000 < 2^17), steps (< 2^21) one with bounces (<= 1001)
            w[0 * B] = (uint32_t)__double2loint(pxd); w[1 * B] = (uint32_t)__double2hiint(pxd);
            w[2 * B] = (uint32_t)__double2loint(pyd); w[3 * B] = (uint32_t)__double2hiint(pyd);
            w[4 * B] = (uint32_t)__double2loint(pzd); w[5 * B] = (uint32_t)__double2hiint(pzd);
            w[6 * B] = __float_as_uint(ux); w[7 * B] = __float_as_uint(uy); w[8 * B] = __float_as_uint(uz);
            w[9 * B] = __float_as_uint(tau); w[10 * B] = __float_as_uint(taurun); w[11 * B] = __float_as_uint(dstep);
            w[12 * B] = __float_as_uint(dlast);
            w[13 * B] = (uint32_t)layer | ((uint32_t)new_layer << 16);
            w[14 * B] = (uint32_t)state | ((uint32_t)phase << 4) | (tflag ? 256u : 0u) | (launch ? 512u : 0u) | (have_pid ? 1024u : 0u) | (ev << 11);
            w[15 * B] = (uint32_t)steps | ((uint32_t)bounces << 21);
            w[16 * B] = (uint32_t)pid; w[17 * B] = (uint32_t)(pid >> 32);
            if (HASDET || PATHLEN) { w[18 * B] = __float_as_uint(sx); w[19 * B] = __float_as_uint(sy); w[20 * B] = __float_as_uint(sz); }
            if (P.survival) w[21 * B] = __float_as_uint(weight);
            __syncthreads();  // (B) all slots written
            if (all_done) break;
            const uint32_t* r = xbuf + threadIdx.x;
            pxd = __hiloint2double((int)r[1 * B], (int)r[0 * B]);
            pyd = __hiloint2double((int)r[3 * B], (int)r[2 * B]);
            pzd = __hiloint2double((int)r[5 * B], (int)r[4 * B]);
            px = (float)pxd; py = (float)pyd; pz = (float)pzd;
            ux = __uint_as_float(r[6 * B]); uy = __uint_as_float(r[7 * B]); uz = __uint_as_float(r[8 * B]);
            tau = __uint_as_float(r[9 * B]); taurun = __uint_as_float(r[10 * B]); dstep = __uint_as_float(r[11 * B]);
            dlast = __uint_as_float(r[12 * B]);
            layer = (int)(r[13 * B] & 0xffffu); new_layer = (int)(r[13 * B] >> 16);
            const uint32_t fl = r[14 * B];
            state = (int)(fl & 15u); phase = (int)((fl >> 4) & 15u); tflag = (fl & 256u) != 0; launch = (fl & 512u) != 0; have_pid = (fl & 1024u) != 0;
            ev = fl >> 11;
            steps = (int)(r[15 * B] & 0x1fffffu); bounces = (int)(r[15 * B] >> 21);
            pid = (unsigned long long)r[16 * B] | ((unsigned long long)r[17 * B] << 32);
            if (HASDET || PATHLEN) { sx = __uint_as_float(r[18 * B]); sy = __uint_as_float(r[19 * B]); sz = __uint_as_float(r[20 * B]); }
            if (P.survival) weight = __uint_as_float(r[21 * B]);
            qs = (state == ST_BND_PROBE || state == ST_CROSS) ? dstep : 0.f;
            xiter = go_tail ? XTAIL : xiter + 1u;
        }
    }

#include "step_macros_undef.inc"

    // ---- epilogue: flush CTA-private detector bins and per-thread counters
    __syncthreads();
    if (PATHLEN && threadIdx.x == 0) {
        P.seg_count[blockIdx.x] = *seg_cnt;
        atomicAdd(P.seg_total, (unsigned long long)*seg_cnt);
    }
    if (HASDET && P.det_in_smem)
        for (int i = threadIdx.x; i < P.det_total; i += blockDim.x)
            if (sbins[i]) atomicAdd(&P.det_bins[i], sbins[i]);
    // every id below nphotons was claimed exactly once
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.counters[C_LAUNCHED], (unsigned long long)P.nphotons);
    unsigned long long cs[C_COUNT] = {c_nscatt, c_sweeps, 0ull, 0ull, 0ull, 0ull, 0ull, c_dethits, c_vox, c_red};
#pragma unroll
    for (int c = 0; c < C_COUNT; ++c) {
        unsigned long long v = cs[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&P.counters[c], v);
    }
}

// ------------------------------------------------------------------------------------------------ the queued kernel
// Same per-packet state machine (step_body.inc), different scheduling.  trace_persistent keeps one packet per THREAD and, in its
// COMPACT variants, re-sorts the CTA's packets by state every iteration behind two CTA-wide barriers: every warp then waits for
// the slowest one (a warp of FP64 Fresnel events takes ~4x a warp of sweeps; ncu: 44 % of the samples of the 41-sphere scene sit
// at that barrier).  Here packets live in SLOTS in shared memory (QSLOTS_PER_THREAD per thread) and four lock-free ring queues
// hold the slot numbers by what the packet needs next:
//     Q_SWEEP (march / boundary probe / crossing probe)   Q_FRESNEL   Q_INTERACT   Q_EMIT (free slot: claim the next packet id)
// A WARP takes up to 32 slots from the fullest queue, loads them (6 x 16 bytes per lane), runs ONE step -- its lanes all start in
// the same state, so the cold event code runs with full warps, FP64 Fresnel included -- stores them and appends each to the queue of
// its new state.  No barrier anywhere in the loop; a warp never waits for another warp's step.  Results are those of the other
// kernels bit for bit where the tallies are integers (streams depend on (seed, packet id, event index) only).
enum : int { Q_SWEEP = 0, Q_FRESNEL, Q_INTERACT, Q_EMIT, Q_COUNT };
#ifndef SMCRT_Q_SLOTS
#define SMCRT_Q_SLOTS 2
#endif
#ifndef SMCRT_Q_KEEP_MIN
#define SMCRT_Q_KEEP_MIN 20
#endif
#ifndef SMCRT_Q_REFILL_MIN
#define SMCRT_Q_REFILL_MIN 1
#endif
constexpr int QSLOTS_PER_THREAD = SMCRT_Q_SLOTS;  // packets in flight per thread
constexpr int QSLOT_WORDS = 24;  // 96 bytes: the 22 words of the compaction exchange, padded to 6 x 16 bytes
struct QueueCtl {
    unsigned int head[Q_COUNT];      // next entry to take
    unsigned int reserved[Q_COUNT];  // producers: entries handed out for writing
    unsigned int published[Q_COUNT]; // producers: entries written (published in reservation order)
    unsigned int retired;            // slots whose EMIT step found the packet pool empty
    unsigned int pad_[3];
};
__host__ __device__ constexpr int queued_smem_bytes(int threads) {
    return (int)sizeof(QueueCtl) + Q_COUNT * threads * QSLOTS_PER_THREAD * 2 + threads * QSLOTS_PER_THREAD * QSLOT_WORDS * 4;
}

template <bool PATHLEN, bool HASDET, int MINBLOCKS, bool SIMPLE, bool LEAN>
__global__ void __launch_bounds__(SMCRT_BLOCK, (MINBLOCKS * 256) / SMCRT_BLOCK) trace_queued(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    {  // stage the scene in shared memory (16-byte vector copies)
        const int4* src = reinterpret_cast<const int4*>(P.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = threadIdx.x; i < P.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    unsigned long long* sbins = reinterpret_cast<unsigned long long*>(smem + P.blob_bytes);
    unsigned int* seg_cnt = reinterpret_cast<unsigned int*>(smem + P.seg_off);  // PATHLEN: segments recorded by this CTA
    if (PATHLEN && threadIdx.x == 0) *seg_cnt = 0u;
    static_assert((SMCRT_BLOCK & (SMCRT_BLOCK - 1)) == 0 && SMCRT_BLOCK * QSLOTS_PER_THREAD <= 65536, "ring indices are masked 16-bit slot numbers");
    const int M = blockDim.x * QSLOTS_PER_THREAD;  // packets in flight per CTA (power of two: the rings wrap with a mask)
    QueueCtl* qc = reinterpret_cast<QueueCtl*>(smem + P.xchg_off);
    unsigned short* ring = reinterpret_cast<unsigned short*>(qc + 1);                  // [Q_COUNT][M]
    // [6][M]: word k of every slot in one plane, so that the lanes of a warp (different slots, same word) spread over the banks --
    // slot-major (96-byte stride) put them on four of the eight 16-byte columns: 3.5 wavefronts per access where 1 is ideal
    uint4* slots = reinterpret_cast<uint4*>(ring + Q_COUNT * M);
    volatile QueueCtl* vq = qc;
    if (threadIdx.x < Q_COUNT) {
        const unsigned int n0 = threadIdx.x == Q_EMIT ? (unsigned int)M : 0u;          // every slot starts free
        qc->head[threadIdx.x] = 0u; qc->reserved[threadIdx.x] = n0; qc->published[threadIdx.x] = n0;
    }
    if (threadIdx.x == 0) qc->retired = 0u;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        ring[Q_EMIT * M + i] = (unsigned short)i;
        slots[3 * M + i] = make_uint4(0u, 0u, (uint32_t)ST_EMIT, 0u);                    // w[14]: state EMIT, no packet id yet
    }
    if (HASDET && P.det_in_smem)
        for (int i = threadIdx.x; i < P.det_total; i += blockDim.x) sbins[i] = 0ull;
    __shared__ FirstFlight ff;
    if (threadIdx.x == 0) {
        ff.ok = 0;
        if (P.ff) ff = *P.ff;
    }
    __syncthreads();
    const SceneView sc = make_view(smem, P);
    const int lane = threadIdx.x & 31;
    if (P.tstamp && blockIdx.x == 0 && threadIdx.x == 0) P.tstamp[0] = globaltimer_ns();
    const unsigned full = 0xffffffffu;
    constexpr bool NEED = false;  // (the plain kernels win on the scenes that use it)
    const unsigned int mask = (unsigned int)M - 1u;
    unsigned int c_nscatt = 0, c_sweeps = 0, c_dethits = 0, c_vox = 0, c_red = 0;

#include "step_macros.inc"
    // The warp keeps the packets that stay in its current class in registers from one iteration to the next and only moves the
    // others: packets that changed class are stored and queued, the lanes they leave (and the lanes of finished packets) are
    // refilled from the queue of the class the warp works on.  A packet is loaded / stored once per class change, not once per
    // step, and a class that only a few lanes of this warp are in (Fresnel events, typically) is never run on those few lanes:
    // they are queued until a whole warp's worth has gathered somewhere in the CTA.
    constexpr int KEEP_MIN = SMCRT_Q_KEEP_MIN;  // fewer held packets than this in every class: queue them all and take a batch of the fullest class
    double pxd = 0, pyd = 0, pzd = 0;
    float px = 0, py = 0, pz = 0, ux = 0, uy = 0, uz = 1, sx = 0, sy = 0, sz = 0;
    float tau = 0.f, taurun = 0.f, dstep = 0.f, qs = 0.f, dlast = 0.f, weight = 1.f;
    int layer = 0, new_layer = 0, state = ST_DONE, bounces = 0, steps = 0;
    bool tflag = false, launch = false, have_pid = false;
    int phase = 0;
    unsigned long long pid = 0;
    uint32_t ev = 0;
    unsigned int slot = 0;
    bool has = false;  // this lane holds a packet (slot `slot`) in registers
    int wc = -1;       // class of the warp's previous iteration (warp-uniform)
    bool draining = false;  // a slot of this warp has found the packet pool empty (warp-uniform)
    for (;;) {
        // ---- which class does the warp work on next
        const int cls = !has ? -1 : (state <= ST_CROSS ? Q_SWEEP : (state == ST_FRESNEL ? Q_FRESNEL : (state == ST_INTERACT ? Q_INTERACT : Q_EMIT)));
        int c = -1;
        unsigned keep = 0u;
        // usual case: enough lanes are still in the class the warp worked on last time (one ballot instead of four)
        const unsigned stay = wc >= 0 ? __ballot_sync(full, cls == wc) : 0u;
        if (__popc(stay) >= (draining ? 1 : KEEP_MIN)) { c = wc; keep = stay; }
        else {
            int bestn = 0;
#pragma unroll
            for (int q = 0; q < Q_COUNT; ++q) {
                const unsigned m = __ballot_sync(full, cls == q);
                const int k = __popc(m);
                if (k > bestn) { bestn = k; c = q; keep = m; }
            }
            // Few packets held: they are all queued and the warp takes a batch of the fullest class, which gathers the stragglers
            // of the CTA's warps.  Not at the END of a launch (one packet is enough once a slot of this warp has found the packet pool
            // empty): there is less and less to gather then, and the last histories (a packet on its 1000 reflections inside a
            // sphere of sphere.toml: 10 ms alone) would be stored and reloaded at every step.  They stay in registers.
            if (bestn < (draining ? 1 : KEEP_MIN)) { c = -1; keep = 0u; }
        }
        // ---- store and queue the packets that leave (all of them when no class is kept)
        const bool push = has && cls != c;
        const unsigned pm = __ballot_sync(full, push);
        if (push) {
            uint4* w = slots + slot;
            w[0] = make_uint4((uint32_t)__double2loint(pxd), (uint32_t)__double2hiint(pxd), (uint32_t)__double2loint(pyd), (uint32_t)__double2hiint(pyd));
            w[M] = make_uint4((uint32_t)__double2loint(pzd), (uint32_t)__double2hiint(pzd), __float_as_uint(ux), __float_as_uint(uy));
            w[2 * M] = make_uint4(__float_as_uint(uz), __float_as_uint(tau), __float_as_uint(taurun), __float_as_uint(dstep));
            w[3 * M] = make_uint4(__float_as_uint(dlast), (uint32_t)layer | ((uint32_t)new_layer << 16),
                                  (uint32_t)state | ((uint32_t)phase << 4) | (tflag ? 256u : 0u) | (launch ? 512u : 0u) | (have_pid ? 1024u : 0u) | (ev << 11),
                                  (uint32_t)steps | ((uint32_t)bounces << 21));
            w[4 * M] = make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), __float_as_uint(sx), __float_as_uint(sy));
            w[5 * M] = make_uint4(__float_as_uint(sz), __float_as_uint(weight), 0u, 0u);
            const unsigned grp = __match_any_sync(pm, cls);
            const int cnt = __popc(grp), lead = __ffs(grp) - 1;
            unsigned int start = 0;
            if (lane == lead) start = atomicAdd(&qc->reserved[cls], (unsigned int)cnt);
            start = __shfl_sync(grp, start, lead);
            ring[cls * M + ((start + (unsigned int)__popc(grp & ((1u << lane) - 1u))) & mask)] = (unsigned short)slot;
            __syncwarp(grp);
            if (lane == lead) {
                __threadfence_block();  // slot contents and ring entries before the publication
                // publish in reservation order.  Bounded by the watchdog period (see below): the wait is for another warp's
                // few-instruction publication, so only a lost queue entry can make it long
                unsigned int spins = 0u;
                unsigned long long t_wd = 0ull;
                while (atomicCAS(&qc->published[cls], start, start + (unsigned int)cnt) != start) {
                    if ((++spins & 0xfffu) == 0u) {
                        const unsigned long long now = globaltimer_ns();
                        if (!t_wd) t_wd = now;
                        else if (now - t_wd > P.watchdog_ns) { atomicAdd(&P.counters[C_SPARE], 1ull); break; }
                    }
                }
            }
            has = false;
            state = ST_DONE;
        }
        __syncwarp(full);
        // ---- (re)fill the free lanes from the queue of the kept class, or take a batch of the fullest queue
        const unsigned freem = ~keep;
        const int want = 32 - __popc(keep);
        if (c < 0 || want >= SMCRT_Q_REFILL_MIN) {  // a nearly full warp does not bother
            unsigned int h0 = 0, n = 0;
            int q = c;
            if (lane == 0) {
                if (q < 0) {  // fullest queue (head first: the difference can only over-estimate)
                    unsigned int bestn = 0;
#pragma unroll
                    for (int k = 0; k < Q_COUNT; ++k) {
                        const unsigned int h = vq->head[k];
                        const unsigned int a = vq->published[k] - h;
                        if (a > bestn) { bestn = a; q = k; }
                    }
                }
                if (q >= 0) {
                    h0 = vq->head[q];
                    n = min(vq->published[q] - h0, (unsigned int)want);
                    if (n && atomicCAS(&qc->head[q], h0, h0 + n) != h0) n = 0;  // another warp was faster: next iteration looks again
                    if (n) __threadfence_block();  // the producers' slot / ring stores (fenced before their publication) before our loads
                }
            }
            n = __shfl_sync(full, n, 0);
            if (n) {
                h0 = __shfl_sync(full, h0, 0);
                q = __shfl_sync(full, q, 0);
                c = q;
                // The entries [h0, h0 + n) cannot be overwritten before they are read here: a ring holds M entries and at most
                // M - n slots can be queued anywhere while this warp holds n of them.
                const unsigned int rank = (unsigned int)__popc(freem & ((1u << lane) - 1u));
                if (((freem >> lane) & 1u) && rank < n) {
                    slot = ring[q * M + ((h0 + rank) & mask)];
                    const uint4* r = slots + slot;
                    const uint4 a = r[0], b = r[M], cc = r[2 * M], d = r[3 * M], e = r[4 * M], f = r[5 * M];
                    pxd = __hiloint2double((int)a.y, (int)a.x); pyd = __hiloint2double((int)a.w, (int)a.z); pzd = __hiloint2double((int)b.y, (int)b.x);
                    px = (float)pxd; py = (float)pyd; pz = (float)pzd;
                    ux = __uint_as_float(b.z); uy = __uint_as_float(b.w); uz = __uint_as_float(cc.x);
                    tau = __uint_as_float(cc.y); taurun = __uint_as_float(cc.z); dstep = __uint_as_float(cc.w);
                    dlast = __uint_as_float(d.x);
                    layer = (int)(d.y & 0xffffu); new_layer = (int)(d.y >> 16);
                    const uint32_t fl = d.z;
                    state = (int)(fl & 15u); phase = (int)((fl >> 4) & 15u); tflag = (fl & 256u) != 0; launch = (fl & 512u) != 0; have_pid = (fl & 1024u) != 0;
                    ev = fl >> 11;
                    steps = (int)(d.w & 0x1fffffu); bounces = (int)(d.w >> 21);
                    pid = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
                    sx = __uint_as_float(e.z); sy = __uint_as_float(e.w); sz = __uint_as_float(f.x);
                    weight = __uint_as_float(f.y);
                    qs = (state == ST_BND_PROBE || state == ST_CROSS) ? dstep : 0.f;
                    has = true;
                }
            }
        }
        wc = c;
        if (!__any_sync(full, has)) {  // nothing held, nothing to take
            // Warp-wide control flow is decided by ONE lane's reads of the shared counters, broadcast: per-lane volatile reads could
            // see different values and split a warp whose later votes and shuffles all use the full mask.
            // watchdog: waiting is normal only while other warps finish the last histories (milliseconds; seconds for a long tail
            // with a raised step cap).  P.watchdog_ns of %globaltimer without any queue entry appearing means the queues have
            // lost a slot: leave with an error flag (smcrt_wait reports it) rather than hang the device.
            // (the wait has its own loop: its state is not live state of the hot loop)
            int verdict = 0;  // 1: work appeared, 2: every slot retired -> the CTA is done, 3: watchdog
            unsigned long long t_wd = 0ull;
            for (unsigned int spins = 0u; !verdict; ++spins) {
                if (lane == 0) {
                    unsigned int a = 0u;
#pragma unroll
                    for (int k = 0; k < Q_COUNT; ++k) a |= vq->published[k] - vq->head[k];
                    if (vq->retired == (unsigned int)M) verdict = 2;
                    else if (a) verdict = 1;
                    else if ((spins & 0xfu) == 0u) {  // (%globaltimer is a slow read: every 16th look)
                        const unsigned long long now = globaltimer_ns();
                        if (!t_wd) t_wd = now;
                        else if (now - t_wd > P.watchdog_ns) verdict = 3;
                    }
                }
                verdict = __shfl_sync(full, verdict, 0);
                if (!verdict) __nanosleep(40);
            }
            if (verdict == 3 && lane == 0) atomicAdd(&P.counters[C_SPARE], 1ull);
            if (verdict >= 2) break;
            continue;
        }

        // ---- one step (lanes without a packet are in ST_DONE and sit it out)
#define STEP_EXIT_CHECK
#include "step_body.inc"
#undef STEP_EXIT_CHECK

        // ---- slots whose EMIT step found the pool empty are retired
        const unsigned gone = __ballot_sync(full, has && state == ST_DONE);
        if (gone) {
            if (lane == 0) atomicAdd(&qc->retired, (unsigned int)__popc(gone));
            if (state == ST_DONE) has = false;
            draining = true;
        }
    }
#include "step_macros_undef.inc"

    // ---- epilogue: flush CTA-private detector bins and per-thread counters
    __syncthreads();
    if (PATHLEN && threadIdx.x == 0) {
        P.seg_count[blockIdx.x] = *seg_cnt;
        atomicAdd(P.seg_total, (unsigned long long)*seg_cnt);
    }
    if (HASDET && P.det_in_smem)
        for (int i = threadIdx.x; i < P.det_total; i += blockDim.x)
            if (sbins[i]) atomicAdd(&P.det_bins[i], sbins[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.counters[C_LAUNCHED], (unsigned long long)P.nphotons);
    unsigned long long cs[C_COUNT] = {c_nscatt, c_sweeps, 0ull, 0ull, 0ull, 0ull, 0ull, c_dethits, c_vox, c_red};
#pragma unroll
    for (int c = 0; c < C_COUNT; ++c) {
        unsigned long long v = cs[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&P.counters[c], v);
    }
}

#ifndef SMCRT_TRACE_TU  // everything below is compiled into engine.cu only (trace_inst.cu holds the trace-kernel instantiations)
// ------------------------------------------------------------------------------------------------ culling-grid set-up
// First flight of a fixed-ray source.  A pencil source launches EVERY packet from the same point in the same
// direction (src/photon.f90:652-710), so the sweeps that start a history -- the launch sweep and, for a source that sits on a
// surface (validation1.toml: on the slab's face), the boundary probe behind it -- evaluate the same points for every packet: 2 of
// the slab scene's 5.9 sweeps per packet.  ONE thread does them once per run, with the trace kernels' own sweep code, and leaves
// what the transitions need: the layer the flight happens in, the probe length already behind the packet (d0), the step to the next
// surface (s) and whether it is exact.  The EMIT step (step_body.inc) then plays the first move directly -- the expressions of the
// ST_MARCH / ST_CROSS transitions on these constants and the packet's own tau -- and the packet enters the loop where its history
// starts to differ from the others'.  Anything unusual at the launch point (Fresnel surface, creep, forward nudge, camera, batched
// sources) leaves ok = 0 and the packets take the ordinary path.  A kernel of its own: the same constants for every kernel variant.
__global__ void first_flight_kernel(const __grid_constant__ KParams P, FirstFlight* out) {
    if (blockIdx.x || threadIdx.x) return;
    FirstFlight f;
    f.ok = 0; f.layer = 0; f.exact = 0; f.pad_ = 0; f.kap = 0.f; f.d0 = 0.f; f.s = 0.f; f.eps = 0.f;
    const SceneView sc = make_view(P.blob, P);
    if (P.src_kind == 2 && !P.src_table && !P.has_camera) {
        const Emitted em = emit_packet(P, 0.f, 0.f, 0.f, 0ull, 0u);
        if (em.ok && in_grid(P, em.x, em.y, em.z)) {
            const float e_ = fmaxf(P.eps0, P.eps_rel * fmaxf(fabsf(em.x), fmaxf(fabsf(em.y), fabsf(em.z))));
            const Sweep S1 = sweep_all<false>(P, sc, em.x, em.y, em.z, em.dx, em.dy, em.dz, SMCRT_BIG);
            f.eps = e_;
            if (S1.L != 0) {
                if (S1.amin < e_) {  // on a boundary: the probe of :77-84, taken as the crossing probe it is (step_body.inc, ST_CROSS)
                    const float d0 = S1.amin + 2.0f * e_;
                    const float qx = (float)((double)em.x + (double)d0 * (double)em.dx), qy = (float)((double)em.y + (double)d0 * (double)em.dy),
                                qz = (float)((double)em.z + (double)d0 * (double)em.dz);
                    const Sweep S2 = sweep_all<false>(P, sc, qx, qy, qz, em.dx, em.dy, em.dz, SMCRT_BIG);
                    if (S2.L != S1.L && S2.L != 0 && S2.amin >= e_ && sc.tops[S1.L - 1].n == sc.tops[S2.L - 1].n) {
                        f.layer = S2.L;
                        f.kap = sc.tops[S2.L - 1].kappa;
                        f.d0 = d0;
                        f.s = S2.bmin < SMCRT_BIG ? fmaxf(S2.amin, S2.bmin - (0.25f * e_ + 2.4e-7f * S2.bmin)) : S2.amin;
                        f.exact = S2.bexact ? 1 : 0;
                        f.ok = 1;
                    }
                } else {            // inside a layer: the first sphere-trace step (:155-176)
                    f.layer = S1.L;
                    f.kap = sc.tops[S1.L - 1].kappa;
                    f.d0 = 0.f;
                    f.s = S1.bmin < SMCRT_BIG ? fmaxf(S1.amin, S1.bmin - (0.25f * e_ + 2.4e-7f * S1.bmin)) : S1.amin;
                    f.exact = S1.bexact ? 1 : 0;
                    f.ok = 1;
                }
            }
        }
    }
    *out = f;
}

// One thread per (cell, top-level SDF): FP64 distance at the cell centre.  The host turns the matrix into candidate lists.
__global__ void cull_eval_kernel(const __grid_constant__ KParams P, long long n_pairs, double lox, double loy, double loz, double dx, double dy,
                                 double dz, int nx, int ny, float* out) {
    const SceneView sc = make_view(P.blob, P);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % P.n_top);
        const long long c = i / P.n_top;
        const int cx = (int)(c % nx), cy = (int)((c / nx) % ny), cz = (int)(c / ((long long)nx * ny));
        out[i] = (float)eval_top_d(P, sc, t, lox + (cx + 0.5) * dx, loy + (cy + 0.5) * dy, loz + (cz + 0.5) * dz);
    }
}

// ------------------------------------------------------------------------------------------------ probes
// Deterministic-component kernels for the parity tests: same device functions as above.
__global__ void probe_sdf_kernel(const __grid_constant__ KParams P, int top_index, long long n, const float* pos, float* dist,
                                 float* normal) {
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < P.blob_bytes / 16; i += blockDim.x)
        reinterpret_cast<int4*>(smem)[i] = reinterpret_cast<const int4*>(P.blob)[i];
    __syncthreads();
    const SceneView sc = make_view(smem, P);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        if (top_index > 0) {
            dist[i] = eval_top_f(sc, top_index - 1, x, y, z);
            if (normal) {
                const double3 nn = surface_normal(P, sc, top_index - 1, x, y, z);
                normal[3 * i] = (float)nn.x; normal[3 * i + 1] = (float)nn.y; normal[3 * i + 2] = (float)nn.z;
            }
        } else
            for (int t = 0; t < P.n_top; ++t) dist[i * P.n_top + t] = eval_top_f(sc, t, x, y, z);
    }
}
// directional step bound of one top-level SDF: what sweep_one feeds into the step decision (distance, bound, exact flag)
__global__ void probe_ray_kernel(const __grid_constant__ KParams P, int top_index, long long n, const float* pos, const float* dir,
                                 float* dist, float* bound, int* exact) {
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < P.blob_bytes / 16; i += blockDim.x)
        reinterpret_cast<int4*>(smem)[i] = reinterpret_cast<const int4*>(P.blob)[i];
    __syncthreads();
    const SceneView sc = make_view(smem, P);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d, b;
        bool ex;
        top_ray<false>(sc, top_index - 1, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], SMCRT_BIG, d, b, ex);
        dist[i] = d; bound[i] = b; exact[i] = ex ? 1 : 0;
    }
}
__global__ void probe_fresnel_kernel(long long n, const float* dir, const float* nrm, const float* n1, const float* n2,
                                     const float* xi, float* dir_out, float* R, int* rflag) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double3 N = make_double3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        const Refl o = reflect_refract(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], N, n1[i], n2[i], xi[i]);
        dir_out[3 * i] = o.x; dir_out[3 * i + 1] = o.y; dir_out[3 * i + 2] = o.z;
        R[i] = o.R;
        rflag[i] = o.reflected ? 1 : 0;
    }
}
__global__ void probe_scatter_kernel(long long n, const float* dir, const float* hgg, const float* xi, float* dir_out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d0 = dir[3 * i], d1 = dir[3 * i + 1], d2 = dir[3 * i + 2];
        hg_scatter(d0, d1, d2, hgg[i], xi[2 * i], xi[2 * i + 1]);
        dir_out[3 * i] = d0; dir_out[3 * i + 1] = d1; dir_out[3 * i + 2] = d2;
    }
}
__global__ void probe_emit_kernel(const __grid_constant__ KParams P, long long n, const float* xi4, float* pos, float* dir, int* cell) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float p[3] = {0, 0, 0}, d[3] = {0, 0, 0};
        // (probe convention for the five-uniform emitters: uniforms 4 and 5 are xi4[3] and 1 - xi4[3])
        emit_packet_v(P, xi4[4 * i], xi4[4 * i + 1], xi4[4 * i + 2], xi4[4 * i + 3], 1.0f - xi4[4 * i + 3], p, d);
        for (int a = 0; a < 3; ++a) { pos[3 * i + a] = p[a]; dir[3 * i + a] = d[a]; }
        // get_voxel_cart (src/grid.f90:51-78)
        const int dims[3] = {P.nxg, P.nyg, P.nzg};
        for (int a = 0; a < 3; ++a) {
            int c = (int)floorf((p[a] + P.gmax[a]) * P.inv_vox[a]) + 1;
            if (c == dims[a] + 1 && p[a] < P.gmax[a]) c = dims[a];  // the scaled form rounds up within ~n ulp of the upper face (cf. in_grid / voxel_of)
            if (c < 1 || c > dims[a]) c = -1;
            cell[3 * i + a] = c;
        }
    }
}
__global__ void probe_detector_kernel(const __grid_constant__ KParams P, int det_index, long long n, const float* start,
                                      const float* dir, const float* len, int* hit, int* bin) {
    const DevDet* D = reinterpret_cast<const DevDet*>(P.blob + P.off_dets) + (det_index - 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s[3] = {start[3 * i], start[3 * i + 1], start[3 * i + 2]};
        const float d[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        const float e[3] = {s[0] + d[0] * len[i], s[1] + d[1] * len[i], s[2] + d[2] * len[i]};
        const int b = detector_bin(D, reinterpret_cast<const float4*>(P.blob + P.off_detp)[det_index - 1], s[0], s[1], s[2], d[0], d[1], d[2], e[0], e[1], e[2]);
        hit[i] = b > 0;
        bin[i] = b;
    }
}

// Sparse read-back of a tally grid (smcrt_fetch): (index, value) pairs of the non-zero voxels, unordered.  A pencil-beam slab run
// touches ~2000 of 1.25e8 voxels; the scan reads the grid once at HBM speed and the host copy shrinks from 500 MB to a few KB.
// Gives up (cursor > cap) as soon as the grid turns out to be dense; the caller then copies it whole.
__global__ void nnz_pack_kernel(const float* __restrict__ g, long long n, unsigned int* __restrict__ idx, float* __restrict__ val,
                                unsigned long long* cursor, unsigned long long cap) {
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
            if (*reinterpret_cast<volatile unsigned long long*>(cursor) > cap) return;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (e[k] != 0.f) {
                    const unsigned long long at = atomicAdd(cursor, 1ull);
                    if (at < cap) { idx[at] = (unsigned int)(4 * i + k); val[at] = e[k]; }
                }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 << 2; i < n; ++i)
            if (g[i] != 0.f) {
                const unsigned long long at = atomicAdd(cursor, 1ull);
                if (at < cap) { idx[at] = (unsigned int)i; val[at] = g[i]; }
            }
}

// sparse reduce (engine.cu: reduce_grid): the root adds the (index, value) pairs received from the other ranks to its grid
__global__ void scatter_add_kernel(float* __restrict__ g, const unsigned int* __restrict__ idx, const float* __restrict__ val, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) atomicAdd(g + idx[i], val[i]);
}

// red.global.add.f32 throughput microbenchmark (the secondary bound of path-length mode).
//   pattern 0: every thread walks its own pseudo-random voxel sequence over the whole grid (L2/HBM scatter)
//   pattern 1: every thread walks the SAME column of `span` voxels (stride nx*ny: the beam axis of a pencil source)
//   pattern 2: DDA-like: each thread walks `span` consecutive voxels along x from a random start (sector-local runs)
__global__ void red_bench_kernel(float* grid, long long nvox, long long stride, int span, int pattern, int ops_per_thread) {
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long h = (tid + 1) * 0x9E3779B97F4A7C15ull;
    long long v = 0;
    for (int i = 0; i < ops_per_thread; ++i) {
        if (pattern == 0) {
            h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
            v = (long long)(h % (unsigned long long)nvox);
        } else if (pattern == 1) {
            v = (nvox / 2 / stride) % stride + (long long)((i + (int)(tid & 1023)) % span) * stride;
        } else {
            if (i % span == 0) { h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31; v = (long long)(h % (unsigned long long)(nvox - span)); }
            else ++v;
        }
        atomicAdd(grid + v, 1.0f);
    }
}

#endif  // SMCRT_TRACE_TU

// Kernel variants: scheduling (one packet per thread / + compaction behind CTA barriers / slot queues) x register
// budget (2, 3 or 4 resident CTAs per SM = 128, 80 or 64 registers).  The instantiations live in four translation units
// (trace_inst.cu compiled once per <PATHLEN, HASDET> pair, in parallel); engine.cu gets the kernels through these pickers.
enum : int { SCHED_PLAIN = 0, SCHED_COMPACT = 1, SCHED_QUEUED = 2 };
typedef void (*trace_kernel_t)(const KParams);
trace_kernel_t pick_kernel_pl0_hd0(int sched, int mb, bool need, bool simple, bool lean);
trace_kernel_t pick_kernel_pl0_hd1(int sched, int mb, bool need, bool simple, bool lean);
trace_kernel_t pick_kernel_pl1_hd0(int sched, int mb, bool need, bool simple, bool lean);
trace_kernel_t pick_kernel_pl1_hd1(int sched, int mb, bool need, bool simple, bool lean);

}  // namespace smcrt_dev
