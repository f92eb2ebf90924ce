// engine.cu — host half of libsmcrt_gpu.so: the C ABI of include/smcrt.h on top of kernels.cuh.
// One smcrt_ctx drives 1..N GPUs from one host thread (the reference's OpenMP threads become GPUs,
// src/kernelsMod.f90:1833-1836); photon ids are split into contiguous ranges per GPU and every GPU owns private
// tallies that are summed with NCCL at fetch time (the intent of the dead mpi_reduce block, :2351-2357).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smcrt.h"
#include "host_math.hpp"
#include "kernels.cuh"

using namespace smcrt_dev;
using smcrt_math::M44;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int set_err(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}
extern "C" void smcrt_set_error_(const char* msg) { g_err = msg ? msg : ""; }
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------ NCCL (lazy)
// Bound at run time so that single-GPU use never needs NCCL and the library has no link-time dependency on a
// particular libnccl build (torch ships its own libnccl.so.2; the system has another).
namespace nccl {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclUint64 = 5, ncclUint32 = 3, ncclSum = 0 };
static void* lib = nullptr;
static ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
static ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
static ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
static ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
static ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*GroupStart)() = nullptr;
static ncclResult_t (*GroupEnd)() = nullptr;
static const char* (*GetErrorString)(ncclResult_t) = nullptr;
static int load() {
    if (lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return set_err("NCCL is required for multi-GPU reduction but libnccl.so.2 could not be loaded: %s", dlerror());
#define SYM(var, name)                                                  \
    *(void**)(&var) = dlsym(lib, name);                                 \
    if (!var) return set_err("libnccl lacks symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Reduce, "ncclReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}
}  // namespace nccl
#define NC(call)                                                                                  \
    do {                                                                                          \
        int r_ = (call);                                                                          \
        if (r_ != 0) return set_err("%s failed: %s", #call, nccl::GetErrorString ? nccl::GetErrorString(r_) : "?"); \
    } while (0)

// ------------------------------------------------------------------------------------------------ context
static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

constexpr int SEG_MAX_SHARES = 148 * 8;  // CTAs of a trace launch (<= SMs x resident CTAs): shares of the segment buffer
constexpr int NVAR_MAX = 8;  // kernel variants a trial can cover (events / time stamps are sized for it)
static bool create_events(cudaEvent_t* ev, int n) {
    for (int i = 0; i < n; ++i) if (cudaEventCreate(&ev[i]) != cudaSuccess) return false;
    return true;
}
struct DeviceState {
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tune_ev[NVAR_MAX + 1] = {};  // brackets of the kernel-variant trial launches
    const float* src_table = nullptr;  // set for the duration of smcrt_run_sources
    unsigned long long* src_tot = nullptr;
    unsigned long long src_id0 = 0;
    long long per_src = 0;
    int tuning = -1;  // >= 0: trial launches in flight for variant key `tuning` (read back in smcrt_wait)
    unsigned char* blob = nullptr;
    DevPrimD* primsD = nullptr;
    DevInstrD* progD = nullptr;
    float *jmean = nullptr, *absorb = nullptr, *emission = nullptr;
    // path-length mode: fixed-point difference grids per axis (allocated by the first path-length run), their
    // "touched" flags, whether deposits are waiting in them, and how many packets have gone in since the last scan
    long long* jdiff[3] = {nullptr, nullptr, nullptr};
    unsigned int* jdiff_used = nullptr;
    bool jdiff_dirty = false;
    long long jdiff_packets = 0;
    // trackHistory: (packet id, detector) of every hit on a history-tracking detector since the last reset
    FirstFlight* ff = nullptr;  // first flight of a pencil source (first_flight_kernel)
    unsigned long long* hist_ids = nullptr;
    int* hist_det = nullptr;
    unsigned long long* hist_n = nullptr;
    // A path-length run of several chunks alternates between two lanes: a stream, a half of the segment buffer and a packet counter
    // each, so that a chunk's trace kernel starts while the last histories of the chunk before are still running (see run_on_device)
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_lane[2] = {nullptr, nullptr};
    // segment buffer of the path-length mode: seg_records records of 32 bytes, split evenly between the CTAs of a trace launch
    float4* seg_buf = nullptr;
    unsigned int* seg_count = nullptr;   // [SEG_MAX_SHARES]
    unsigned long long* seg_total = nullptr;
    size_t seg_records = 0;
    unsigned long long* det_bins = nullptr;
    unsigned long long* counters = nullptr;  // C_COUNT + 1 (work counter last) + 2 x 8 time stamps of the variant trial
    // sparse read-back scratch (smcrt_fetch): device pair list + cursor, pinned host mirror
    unsigned int* nz_idx = nullptr;
    float* nz_val = nullptr;
    unsigned long long* nz_cursor = nullptr;
    unsigned int* nz_idx_h = nullptr;
    float* nz_val_h = nullptr;
    size_t nz_cap = 0;
    // sparse reduce (reduce_buffers): every rank's pair count, and on the root the pairs received from the others
    unsigned long long* nz_counts = nullptr;
    unsigned int* rx_idx = nullptr;
    float* rx_val = nullptr;
    size_t rx_cap = 0;
    int* cull_start = nullptr;
    int* cull_items = nullptr;
    float* cull_far = nullptr;
    float* cull_clear = nullptr;
    nccl::ncclComm_t comm = nullptr;
    int sm_count = 148;
    bool ran = false;
};

struct HostDet {
    int kind, nbins_user, stored;
    long long count, offset;
};

struct smcrt_ctx {
    std::vector<DeviceState> devs;
    // grid
    int nxg = 0, nyg = 0, nzg = 0;
    double gmax[3] = {1, 1, 1};
    // scene (host copies)
    std::vector<DevPrim> prims;
    std::vector<DevPrimD> primsD;
    std::vector<DevTop> tops;
    std::vector<DevInstr> prog;
    std::vector<DevInstrD> progD;
    std::vector<DevDet> dets;
    std::vector<HostDet> hdets;
    long long det_total = 0;
    std::vector<double> opt_mus, opt_mua, opt_hgg, opt_n;
    bool scene_dirty = true;
    int off_tops = 0, off_prog = 0, off_dets = 0, off_hot = 0, off_detp = 0, blob_bytes = 0;
    // source
    int src_kind = SMCRT_SRC_POINT, src_sub = 0, src_alt = 0;
    float sp[24] = {0};
    float Tpos[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    float Tdir[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    // knobs
    double eps0 = 1e-8, eps_rel = 4.76837158203125e-07 /* 2^-21 = 4 ulp(1.0f) */;
    long long max_steps = 200000;
    // run state
    bool comm_all = false;    // in-process communicator over devs
    bool comm_rank = false;   // one-rank-per-process communicator (devs.size()==1)
    int nranks = 1, rank = 0;
    double last_ms = 0;
    long long launches = 0;
    bool pending = false;
    // culling grid (built at upload time for scenes with many top-level SDFs)
    bool cull_on = false, cull_allowed = true, scene_lipschitz = true, compact_allowed = false;
    std::vector<int32_t> track;  // per detector: trackHistory
    bool any_track = false;
    int seg_inline = -1;        // path-length mode: 1 = the trace kernels walk their segments themselves, 0 = recorded + deposit kernel, -1 = not timed yet
    double seg_per_packet = 0;  // measured straight segments per packet of the current scene (path-length mode); 0 = not measured
    bool dda_legacy = false;  // SMCRT_DDA_LEGACY at smcrt_create: path-length deposits one red per voxel crossed (A/B switch, tests)
    // kernel-variant choice per tally configuration [pathlength][detectors]: 0 = not timed yet, else 1 + index into VARIANTS
    int tuned_mb[2][2] = {{0, 0}, {0, 0}};
    uint64_t last_fetch_bytes = 0;  // device -> host bytes moved by the last smcrt_fetch
    uint64_t scene_hash = 0, det_hash = 0;  // tuned_mb is kept while the scene and detectors stay bit-identical
    int cull_n[3] = {0, 0, 0};
    double cull_lo[3] = {0, 0, 0}, cull_cell[3] = {1, 1, 1};
    double cull_mean_list = 0;
    uint64_t cull_key = 0;  // what the current culling grid was built for (scene hash, grid box)
    bool cull_key_valid = false;
    int touched_modes = 0;   // OR of the tally modes run since the last reset: only those grids are reduced
    int dirty_modes = 7;     // grids that may hold non-zero voxels on SOME device (cleared by smcrt_reset_tallies)
    const unsigned long long* replay_ids = nullptr;  // set for the duration of smcrt_history_replay (device pointers)
    float4* replay_vert = nullptr;
    int *replay_nvert = nullptr, *replay_hit = nullptr;
    unsigned long long* replay_bins = nullptr;
    int replay_max_vert = 0;
    long long dbg_pid = -1;
    float* dbg_log = nullptr;
    int dbg_cap = 0;
};

static int n_voxels(const smcrt_ctx* c, size_t* out) {
    *out = (size_t)c->nxg * c->nyg * c->nzg;
    return 0;
}

extern "C" const char* smcrt_last_error(void) { return g_err.c_str(); }
extern "C" const char* smcrt_version(void) { return "smcrt-b200 0.1 (sm_100a)"; }

extern "C" int smcrt_create(smcrt_ctx** out, int n_gpus, const int* device_ids) {
    if (!out) return set_err("smcrt_create: out is null");
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail == 0)
        return set_err("smcrt_create: no CUDA device available (%s); this engine has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (n_gpus == 0) n_gpus = avail;
    if (n_gpus < 0 || n_gpus > avail) return set_err("smcrt_create: %d GPUs requested, %d visible", n_gpus, avail);
    smcrt_ctx* c = new smcrt_ctx();
    c->cull_allowed = getenv("SMCRT_NO_CULL") == nullptr;  // A/B switch for the culling grid (tests, profiling)
    // Event compaction is OPT-IN: measured slower than the plain persistent loop on every shipped scene
    c->compact_allowed = getenv("SMCRT_COMPACT") != nullptr;
    c->dda_legacy = getenv("SMCRT_DDA_LEGACY") != nullptr;
    c->devs.resize(n_gpus);
    for (int g = 0; g < n_gpus; ++g) {
        DeviceState& D = c->devs[g];
        D.dev = device_ids ? device_ids[g] : g;
        if (D.dev < 0 || D.dev >= avail) {
            delete c;
            return set_err("smcrt_create: device id %d out of range", D.dev);
        }
        cudaDeviceProp prop;
        if (cudaSetDevice(D.dev) != cudaSuccess || cudaGetDeviceProperties(&prop, D.dev) != cudaSuccess) {
            delete c;
            return set_err("smcrt_create: cannot select device %d", D.dev);
        }
        if (prop.major < 10) {
            delete c;
            return set_err("smcrt_create: device %d is sm_%d%d; this library contains sm_100a code only", D.dev, prop.major, prop.minor);
        }
        D.sm_count = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&D.ev0) != cudaSuccess || cudaEventCreate(&D.ev1) != cudaSuccess ||
            !create_events(D.tune_ev, NVAR_MAX + 1) ||
            cudaMalloc(&D.counters, sizeof(unsigned long long) * (C_COUNT + 1 + 16 + 2)) != cudaSuccess ||  // counters, next, 16 time stamps, the lanes' next
            cudaMemset(D.counters, 0, sizeof(unsigned long long) * (C_COUNT + 1 + 16 + 2)) != cudaSuccess) {
            delete c;
            return set_err("smcrt_create: resource allocation failed on device %d: %s", D.dev, cudaGetErrorString(cudaGetLastError()));
        }
    }
    *out = c;
    return 0;
}

static void free_grids(DeviceState& D) {
    cudaSetDevice(D.dev);
    cudaFree(D.jmean); cudaFree(D.absorb); cudaFree(D.emission);
    D.jmean = D.absorb = D.emission = nullptr;
    for (int a = 0; a < 3; ++a) { cudaFree(D.jdiff[a]); D.jdiff[a] = nullptr; }
    D.jdiff_dirty = false; D.jdiff_packets = 0;
}

extern "C" void smcrt_destroy(smcrt_ctx* c) {
    if (!c) return;
    for (DeviceState& D : c->devs) {
        cudaSetDevice(D.dev);
        if (D.stream) cudaStreamSynchronize(D.stream);
        if (D.comm && nccl::CommDestroy) nccl::CommDestroy(D.comm);
        free_grids(D);
        cudaFree(D.blob); cudaFree(D.primsD); cudaFree(D.progD); cudaFree(D.det_bins); cudaFree(D.counters); cudaFree(D.jdiff_used); cudaFree(D.ff); cudaFree(D.hist_ids); cudaFree(D.hist_det); cudaFree(D.hist_n); cudaFree(D.seg_buf); cudaFree(D.seg_count); cudaFree(D.seg_total);
        cudaFree(D.cull_start); cudaFree(D.cull_items); cudaFree(D.cull_far); cudaFree(D.cull_clear);
        cudaFree(D.nz_idx); cudaFree(D.nz_val); cudaFree(D.nz_cursor); cudaFreeHost(D.nz_idx_h); cudaFreeHost(D.nz_val_h);
        cudaFree(D.nz_counts); cudaFree(D.rx_idx); cudaFree(D.rx_val);
        if (D.ev0) cudaEventDestroy(D.ev0);
        if (D.ev1) cudaEventDestroy(D.ev1);
        for (cudaEvent_t e : D.tune_ev) if (e) cudaEventDestroy(e);
        for (cudaStream_t s : D.lane_stream) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
        for (cudaEvent_t ev : {D.ev_fork, D.ev_lane[0], D.ev_lane[1]}) if (ev) cudaEventDestroy(ev);
        if (D.stream) cudaStreamDestroy(D.stream);
    }
    delete c;
}

extern "C" int smcrt_set_grid(smcrt_ctx* c, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax) {
    if (!c) return set_err("null ctx");
    if (nxg < 1 || nyg < 1 || nzg < 1 || !(xmax > 0) || !(ymax > 0) || !(zmax > 0)) return set_err("smcrt_set_grid: invalid grid");
    c->nxg = nxg; c->nyg = nyg; c->nzg = nzg;
    c->gmax[0] = xmax; c->gmax[1] = ymax; c->gmax[2] = zmax;
    c->scene_dirty = true;  // the culling grid spans the voxel-grid box
    const size_t bytes = (size_t)nxg * nyg * nzg * sizeof(float);
    for (DeviceState& D : c->devs) {
        free_grids(D);
        CU(cudaSetDevice(D.dev));
        CU(cudaMalloc(&D.jmean, bytes));
        CU(cudaMalloc(&D.absorb, bytes));
        CU(cudaMalloc(&D.emission, bytes));
        CU(cudaMemsetAsync(D.jmean, 0, bytes, D.stream));
        CU(cudaMemsetAsync(D.absorb, 0, bytes, D.stream));
        CU(cudaMemsetAsync(D.emission, 0, bytes, D.stream));
        CU(cudaStreamSynchronize(D.stream));
    }
    // scratch of the sparse read-back (smcrt_fetch, device 0) and of the sparse reduce (every device): pair list for up to 1/64 of
    // the voxels; pinned host mirror on device 0
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceState& D = c->devs[g];
        const size_t nv = (size_t)nxg * nyg * nzg;
        CU(cudaSetDevice(D.dev));
        cudaFree(D.nz_idx); cudaFree(D.nz_val); cudaFreeHost(D.nz_idx_h); cudaFreeHost(D.nz_val_h); cudaFree(D.rx_idx); cudaFree(D.rx_val);
        D.nz_idx = nullptr; D.nz_val = nullptr; D.nz_idx_h = nullptr; D.nz_val_h = nullptr; D.nz_cap = 0;
        D.rx_idx = nullptr; D.rx_val = nullptr; D.rx_cap = 0;
        if (nv < (1ull << 32) && nv >= (1u << 16)) {
            const size_t cap = std::max<size_t>(nv / 64, 4096);
            if (!D.nz_cursor) CU(cudaMalloc(&D.nz_cursor, 8));
            CU(cudaMalloc(&D.nz_idx, cap * 4));
            CU(cudaMalloc(&D.nz_val, cap * 4));
            if (g == 0) {
                CU(cudaMallocHost(&D.nz_idx_h, cap * 4));
                CU(cudaMallocHost(&D.nz_val_h, cap * 4));
            }
            D.nz_cap = cap;
            cudaFuncAttributes fa;
            CU(cudaFuncGetAttributes(&fa, nnz_pack_kernel));  // loads the kernel now rather than inside the first fetch
        }
    }
    return 0;
}

// ---- scene -----------------------------------------------------------------------------------------------
static void fold_transform(const double* xf /*16, Fortran order*/, double m[12], int* cls) {
    // p'_j = sum_i M(i,j) p_i + M(4,j)  (src/vector_class.f90:300-302), M(i,j) at xf[(j-1)*4 + (i-1)]
    for (int j = 0; j < 3; ++j) {
        m[4 * j + 0] = xf[j * 4 + 0];
        m[4 * j + 1] = xf[j * 4 + 1];
        m[4 * j + 2] = xf[j * 4 + 2];
        m[4 * j + 3] = xf[j * 4 + 3];
    }
    const bool rot_id = m[0] == 1 && m[1] == 0 && m[2] == 0 && m[4] == 0 && m[5] == 1 && m[6] == 0 && m[8] == 0 && m[9] == 0 && m[10] == 1;
    const bool no_t = m[3] == 0 && m[7] == 0 && m[11] == 0;
    bool rigid = true;  // rows of the 3x3 block orthonormal
    for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b) {
            const double dot = m[4 * a] * m[4 * b] + m[4 * a + 1] * m[4 * b + 1] + m[4 * a + 2] * m[4 * b + 2];
            if (std::fabs(dot - (a == b ? 1.0 : 0.0)) > 1e-9) rigid = false;
        }
    *cls = rot_id ? (no_t ? XF_IDENTITY : XF_TRANSLATE) : (rigid ? XF_AFFINE : XF_NONRIGID);
}

struct Compiler {
    int n_nodes;
    const int32_t *kind, *first_child, *n_child;
    const double *xform, *params;
    std::vector<DevPrim>* prims;
    std::vector<DevPrimD>* primsD;
    std::vector<DevInstr>* prog;
    std::vector<DevInstrD>* progD;
    int dmax = 0, pmax = 0;
    std::string err;

    int add_prim(int node) {
        DevPrim a{};
        DevPrimD b{};
        double m[12];
        int cls;
        fold_transform(xform + 16 * (size_t)node, m, &cls);
        a.kind = b.kind = kind[node];
        a.xf = b.xf = cls;
        for (int i = 0; i < 12; ++i) { a.m[i] = (float)m[i]; b.m[i] = m[i]; }
        for (int i = 0; i < 8; ++i) { a.p[i] = (float)params[8 * (size_t)node + i]; b.p[i] = params[8 * (size_t)node + i]; }
        prims->push_back(a);
        primsD->push_back(b);
        return (int)prims->size() - 1;
    }
    void emit(int op, int a, double f0 = 0, double f1 = 0, double f2 = 0, double g = 0) {
        DevInstr i{};
        i.op = op; i.a = a; i.f[0] = (float)f0; i.f[1] = (float)f1; i.f[2] = (float)f2; i.g = (float)g;
        DevInstrD d{};
        d.op = op; d.a = a; d.f[0] = f0; d.f[1] = f1; d.f[2] = f2; d.g = g;
        prog->push_back(i);
        progD->push_back(d);
    }
    // returns false on error; dd / pd: current stack depths
    bool gen(int node, int dd, int pd, int depth) {
        if (node < 0 || node >= n_nodes) { err = "node index out of range"; return false; }
        if (depth > 32) { err = "scene tree too deep (cycle?)"; return false; }
        const int k = kind[node];
        const double* p = params + 8 * (size_t)node;
        if (k >= SMCRT_SPHERE && k <= SMCRT_PLANE) {
            emit(I_PRIM, add_prim(node));
            dmax = std::max(dmax, dd + 1);
            return true;
        }
        const int nc = n_child[node], fc = first_child[node];
        if (k >= SMCRT_MODEL_UNION && k <= SMCRT_MODEL_INTERSECTION) {
            if (nc < 1) { err = "model node without children"; return false; }
            if (!gen(fc, dd, pd, depth + 1)) return false;
            for (int i = 1; i < nc; ++i) {
                if (!gen(fc + i, dd + 1, pd, depth + 1)) return false;
                emit(I_CSG, k, p[0]);
            }
            return true;
        }
        if (nc != 1) { err = "modifier node must have exactly one child"; return false; }
        switch (k) {
            case SMCRT_MOD_EXTRUDE:
                if (!gen(fc, dd, pd, depth + 1)) return false;
                emit(I_EXTRUDE, 0, p[0]);
                return true;
            case SMCRT_MOD_ONION:
                if (!gen(fc, dd, pd, depth + 1)) return false;
                emit(I_ONION, 0, p[0]);
                return true;
            case SMCRT_MOD_REVOLUTION:
                emit(I_PUSH_REV, 0, p[0], p[1], p[2], p[3]);
                pmax = std::max(pmax, pd + 1);
                if (!gen(fc, dd, pd + 1, depth + 1)) return false;
                emit(I_POP_P, 0);
                return true;
            case SMCRT_MOD_TWIST:
            case SMCRT_MOD_BEND:
                emit(k == SMCRT_MOD_TWIST ? I_PUSH_TWIST : I_PUSH_BEND, 0, p[0]);
                pmax = std::max(pmax, pd + 1);
                if (!gen(fc, dd, pd + 1, depth + 1)) return false;
                emit(I_POP_P, 0);
                return true;
            case SMCRT_MOD_ELONGATE:
                emit(I_PUSH_ELONG, 0, p[0], p[1], p[2]);
                pmax = std::max(pmax, pd + 1);
                if (!gen(fc, dd, pd + 1, depth + 1)) return false;
                emit(I_POP_P, 0);
                emit(I_ELONG_ADD, 0, p[0], p[1], p[2]);
                return true;
        }
        err = "unsupported node kind " + std::to_string(k);
        return false;
    }
};

static void set_optics(DevTop& T, double mus, double mua, double hgg, double n) {
    // mono(), src/opticalProps/opticalProperties.f90:107-125
    const double kappa = mus + mua;
    const double albedo = (mua < 1e-9) ? 1.0 : mus / kappa;
    T.kappa = (float)kappa; T.albedo = (float)albedo; T.hgg = (float)hgg; T.n = (float)n; T.mua = (float)mua;
}

extern "C" int smcrt_set_scene(smcrt_ctx* c, int n_nodes, const int32_t* kind, const int32_t* first_child, const int32_t* n_child,
                               const double* xform, const double* params, int n_top, const int32_t* top_node, const double* mus,
                               const double* mua, const double* hgg, const double* n_ref) {
    if (!c) return set_err("null ctx");
    if (n_nodes < 1 || n_top < 1 || !kind || !xform || !params || !top_node || !mus || !mua || !hgg || !n_ref)
        return set_err("smcrt_set_scene: invalid arguments");
    std::vector<int32_t> zeros(n_nodes, 0);
    if (!first_child) first_child = zeros.data();
    if (!n_child) n_child = zeros.data();
    std::vector<DevPrim> prims;
    std::vector<DevPrimD> primsD;
    std::vector<DevInstr> prog;
    std::vector<DevInstrD> progD;
    std::vector<DevTop> tops(n_top);
    Compiler comp{n_nodes, kind, first_child, n_child, xform, params, &prims, &primsD, &prog, &progD};
    for (int t = 0; t < n_top; ++t) {
        const int node = top_node[t];
        if (node < 0 || node >= n_nodes) return set_err("smcrt_set_scene: top_node[%d]=%d out of range", t, node);
        DevTop& T = tops[t];
        const int k = kind[node];
        if (k >= SMCRT_SPHERE && k <= SMCRT_PLANE) {
            T.mode = 0;
            T.first = comp.add_prim(node);
            T.count = 1;
        } else {
            T.mode = 1;
            T.first = (int)prog.size();
            comp.dmax = comp.pmax = 0;
            if (!comp.gen(node, 0, 0, 0)) return set_err("smcrt_set_scene: top-level SDF %d: %s", t + 1, comp.err.c_str());
            if (comp.dmax > 4 || comp.pmax > 3)
                return set_err("smcrt_set_scene: top-level SDF %d needs stack depth (%d distances, %d points) beyond the engine's (4,3)",
                               t + 1, comp.dmax, comp.pmax);
            T.count = (int)prog.size() - T.first;
        }
        set_optics(T, mus[t], mua[t], hgg[t], n_ref[t]);
    }
    // culling needs |d(p) - d(q)| <= |p - q|: true for every primitive and CSG/modifier of the reference except twist / bend
    // (they warp space) and for rigid transforms only
    bool lip = true;
    for (int i = 0; i < n_nodes; ++i) {
        if (kind[i] == SMCRT_MOD_TWIST || kind[i] == SMCRT_MOD_BEND) lip = false;
        const double* m = xform + 16 * (size_t)i;
        for (int a = 0; a < 3 && lip; ++a)
            for (int b = 0; b < 3; ++b) {
                double dot = 0;
                for (int k = 0; k < 3; ++k) dot += m[a * 4 + k] * m[b * 4 + k];
                if (std::fabs(dot - (a == b ? 1.0 : 0.0)) > 1e-9) lip = false;
            }
    }
    c->scene_lipschitz = lip;
    c->prims.swap(prims); c->primsD.swap(primsD); c->prog.swap(prog); c->progD.swap(progD); c->tops.swap(tops);
    {  // a different scene invalidates the register-budget choice (same bytes: e.g. a driver re-sending its scene every call)
        uint64_t h = fnv1a(1469598103934665603ull, c->primsD.data(), c->primsD.size() * sizeof(DevPrimD));
        h = fnv1a(h, c->progD.data(), c->progD.size() * sizeof(DevInstrD));
        h = fnv1a(h, c->tops.data(), c->tops.size() * sizeof(DevTop));
        if (h != c->scene_hash) { std::memset(c->tuned_mb, 0, sizeof c->tuned_mb); c->seg_per_packet = 0; c->seg_inline = -1; }
        c->scene_hash = h;
    }
    c->opt_mus.assign(mus, mus + n_top); c->opt_mua.assign(mua, mua + n_top);
    c->opt_hgg.assign(hgg, hgg + n_top); c->opt_n.assign(n_ref, n_ref + n_top);
    c->scene_dirty = true;
    return 0;
}

extern "C" int smcrt_set_optprops(smcrt_ctx* c, int top_index, double mus, double mua, double hgg, double n_ref) {
    if (!c) return set_err("null ctx");
    if (top_index < 1 || top_index > (int)c->tops.size()) return set_err("smcrt_set_optprops: top_index %d out of range", top_index);
    set_optics(c->tops[top_index - 1], mus, mua, hgg, n_ref);
    c->opt_mus[top_index - 1] = mus; c->opt_mua[top_index - 1] = mua; c->opt_hgg[top_index - 1] = hgg; c->opt_n[top_index - 1] = n_ref;
    c->scene_dirty = true;
    return 0;
}

// ---- source ----------------------------------------------------------------------------------------------
extern "C" int smcrt_set_source(smcrt_ctx* c, int kind, int subtype, const double* p) {
    if (!c || !p) return set_err("smcrt_set_source: null argument");
    if (kind < SMCRT_SRC_POINT || kind > SMCRT_SRC_APERTURE) return set_err("No such source!");  // init_source, photon.f90:155
    if ((kind == SMCRT_SRC_FOCUS || kind == SMCRT_SRC_ANNULUS) && (subtype < 1 || subtype > 3)) return set_err("No such beam type!");
    c->src_kind = kind; c->src_sub = subtype; c->src_alt = 0;
    for (int i = 0; i < 24; ++i) c->sp[i] = (float)p[i];
    M44 Tp = smcrt_math::identity(), Td = smcrt_math::identity();
    const double* o = p + SMCRT_SP_POS;
    if (kind == SMCRT_SRC_CIRCULAR) {  // photon.f90:243-264
        double a[3] = {1, 0, 0}, b[3] = {p[SMCRT_SP_DIR], p[SMCRT_SP_DIR + 1], p[SMCRT_SP_DIR + 2]};
        smcrt_math::normalise(b);
        if (std::fabs(a[0]) == std::fabs(b[0]) && std::fabs(a[1]) == std::fabs(b[1]) && std::fabs(a[2]) == std::fabs(b[2])) {
            a[0] = 0; a[2] = 1;
            c->src_alt = 1;
        }
        Tp = smcrt_math::matmul(smcrt_math::rotation_align(a, b), smcrt_math::invert_affine(smcrt_math::translate(o[0], o[1], o[2])));
    } else if (kind == SMCRT_SRC_FOCUS || kind == SMCRT_SRC_ANNULUS) {  // photon.f90:440-478, :927-958
        double a[3] = {0, 0, -1}, b[3] = {p[SMCRT_SP_ROT], p[SMCRT_SP_ROT + 1], p[SMCRT_SP_ROT + 2]};
        smcrt_math::normalise(b);
        const bool same = a[0] == b[0] && a[1] == b[1] && a[2] == b[2];
        const bool absame = std::fabs(a[0]) == std::fabs(b[0]) && std::fabs(a[1]) == std::fabs(b[1]) && std::fabs(a[2]) == std::fabs(b[2]);
        M44 t = smcrt_math::identity();
        if (same) {
        } else if (absame) t.a(3, 3) = -1.0;
        else t = smcrt_math::rotation_align(a, b);
        Td = t;  // dir = dir .dot. t
        if (absame && !same) t.a(3, 3) = 1.0;
        Tp = smcrt_math::matmul(t, smcrt_math::invert_affine(smcrt_math::translate(-o[0], -o[1], -o[2])));
    }
    for (int j = 0; j < 3; ++j) {
        for (int i = 0; i < 3; ++i) {
            c->Tpos[4 * j + i] = (float)Tp.a(i + 1, j + 1);
            c->Tdir[3 * j + i] = (float)Td.a(i + 1, j + 1);
        }
        c->Tpos[4 * j + 3] = (float)Tp.a(4, j + 1);
    }
    return 0;
}

// ---- detectors -------------------------------------------------------------------------------------------
extern "C" int smcrt_set_detectors(smcrt_ctx* c, int n, const int32_t* kind, const double* p, const int32_t* nbins) {
    if (!c) return set_err("null ctx");
    if (n < 0 || (n > 0 && (!kind || !p || !nbins))) return set_err("smcrt_set_detectors: invalid arguments");
    std::vector<DevDet> dets(n);
    std::vector<HostDet> hd(n);
    long long off = 0;
    for (int i = 0; i < n; ++i) {
        const double* q = p + (size_t)SMCRT_DET_PARAMS * i;
        DevDet& D = dets[i];
        std::memset(&D, 0, sizeof D);
        D.kind = kind[i];
        const int nb = nbins[i];
        if (nb < 0) return set_err("smcrt_set_detectors: negative nbins");
        const int stored = nb + 1;  // "extra bin for data beyond end of array", detectors.f90:133
        D.nbins = stored;
        for (int a = 0; a < 3; ++a) { D.pos[a] = (float)q[a]; D.dir[a] = (float)q[3 + a]; }
        long long count = stored;
        switch (kind[i]) {
            case SMCRT_DET_CIRCLE:
                D.q[0] = (float)q[6];
                D.q[1] = (float)(nb == 0 ? 1.0 : q[6] / nb);
                break;
            case SMCRT_DET_ANNULUS:
                D.q[0] = (float)q[6]; D.q[1] = (float)q[7];
                D.q[2] = (float)(nb == 0 ? 1.0 : (q[7] - q[6]) / nb);
                break;
            case SMCRT_DET_FIBRE:
                for (int a = 0; a < 3; ++a) D.pos[a] = (float)(q[a] + q[3 + a] * q[10]);  // pos + dir*frontOffset
                D.q[0] = (float)q[8];   // f1Aperture
                D.q[1] = (float)q[6];   // focalLength1
                D.q[2] = (float)q[7];   // focalLength2
                D.q[3] = (float)q[12];  // frontToPinSep
                D.q[4] = (float)q[14];  // pinAperture
                D.q[5] = (float)q[13];  // pinToBackSep
                D.q[6] = (float)q[9];   // f2Aperture
                D.q[7] = (float)q[11];  // backOffset
                D.q[8] = (float)q[15];  // acceptAngle
                D.q[9] = (float)(q[16] / 2);
                D.q[10] = (float)(nb == 0 ? 1.0 : q[16] / 2 / nb);
                break;
            case SMCRT_DET_CAMERA: {  // init_camera, detectors.f90:395-445
                const double e1[3] = {q[3] - q[0], q[4] - q[1], q[5] - q[2]}, e2[3] = {q[6] - q[0], q[7] - q[1], q[8] - q[2]};
                double nn[3] = {e2[1] * e1[2] - e2[2] * e1[1], -e2[0] * e1[2] + e2[2] * e1[0], e2[0] * e1[1] - e2[1] * e1[0]};
                smcrt_math::normalise(nn);
                const double w = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
                const double h = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
                for (int a = 0; a < 3; ++a) { D.dir[a] = (float)nn[a]; D.q[a] = (float)e1[a]; D.q[3 + a] = (float)e2[a]; }
                D.q[6] = (float)w; D.q[7] = (float)h;
                D.q[8] = (float)(nb == 0 ? 1.0 : q[9] / stored);
                D.q[9] = D.q[8];
                D.q[10] = (float)q[0]; D.q[11] = (float)q[1];
                count = (long long)stored * stored;
                break;
            }
            default: return set_err("Invalid detector type. Valid types are [circle, annulus, camera]");
        }
        D.q[13] = (float)((double)D.pos[0] * (double)D.dir[0] + (double)D.pos[1] * (double)D.dir[1] + (double)D.pos[2] * (double)D.dir[2]);
        D.pad_ = 0;  // trackHistory: set by smcrt_set_track_history (a new detector table clears it, like a new dects(:))
        D.offset = (int)off;
        hd[i] = HostDet{kind[i], nb, stored, count, off};
        off += count;
        if (off > (1ll << 30)) return set_err("smcrt_set_detectors: too many detector bins");
    }
    c->dets.swap(dets); c->hdets.swap(hd); c->det_total = off;
    c->track.assign((size_t)n, 0); c->any_track = false;
    {
        const uint64_t h = fnv1a(1469598103934665603ull, c->dets.data(), c->dets.size() * sizeof(DevDet));
        if (h != c->det_hash) std::memset(c->tuned_mb, 0, sizeof c->tuned_mb);
        c->det_hash = h;
    }
    c->scene_dirty = true;
    for (DeviceState& D : c->devs) {
        CU(cudaSetDevice(D.dev));
        cudaFree(D.det_bins);
        D.det_bins = nullptr;
        const size_t bytes = sizeof(unsigned long long) * (size_t)std::max<long long>(off, 1);
        CU(cudaMalloc(&D.det_bins, bytes));
        CU(cudaMemset(D.det_bins, 0, bytes));
    }
    return 0;
}
extern "C" int64_t smcrt_det_bins_total(const smcrt_ctx* c) { return c ? c->det_total : 0; }

extern "C" int smcrt_set_tolerances(smcrt_ctx* c, double eps0, double eps_rel, int64_t max_steps) {
    if (!c) return set_err("null ctx");
    if (eps0 > 0) c->eps0 = eps0;
    if (eps_rel > 0) c->eps_rel = eps_rel;
    if (max_steps > 0) c->max_steps = max_steps;
    return 0;
}

// ---- upload ----------------------------------------------------------------------------------------------
static int align16(int v) { return (v + 15) & ~15; }
static int fill_params(smcrt_ctx* c, DeviceState& D, KParams& P);
static const int CULL_MIN_TOPS = 8;  // below this the uniform sweep is cheaper than the indirection

// Culling grid.  For every coarse cell: dc_j = d_j(centre) in FP64 on the GPU, h = half diagonal
// (+ slack); 1-Lipschitz => d_j in [dc_j - h, dc_j + h] over the cell.
//   A (can attain min|d|):           |dc_j| - h <= U,  U = min_k (|dc_k| + h)
//   B (can be the innermost negative): dc_j - h < 0 and dc_j + h >= M,  M = max{ dc_k - h : dc_k + h < 0 }
// list = A u B (ascending index, so the "ties -> lowest index" rule of maxloc survives); far = min over the rest of |dc_j| - h.
static int build_cull(smcrt_ctx* c) {
    // The grid depends on the scene's geometry and on the voxel-grid box only: a driver that re-sends the identical scene every
    // call (the escape-function loops; bench.py's end-to-end leg) keeps it, like the kernel-variant choice.
    uint64_t key = fnv1a(c->scene_hash ^ 0x9E3779B97F4A7C15ull, c->gmax, sizeof c->gmax);
    const int dims[4] = {c->nxg, c->nyg, c->nzg, c->cull_allowed ? 1 : 0};
    key = fnv1a(key, dims, sizeof dims);
    if (c->cull_key_valid && c->cull_key == key) return 0;
    c->cull_key = key; c->cull_key_valid = true;
    c->cull_on = false;
    const int nt = (int)c->tops.size();
    for (DeviceState& D : c->devs) {
        cudaSetDevice(D.dev);
        cudaFree(D.cull_start); cudaFree(D.cull_items); cudaFree(D.cull_far); cudaFree(D.cull_clear);
        D.cull_start = D.cull_items = nullptr; D.cull_far = nullptr; D.cull_clear = nullptr;
    }
    if (!c->cull_allowed || !c->scene_lipschitz || nt < CULL_MIN_TOPS || c->nxg == 0) return 0;
    // G^3 cells, G = 10 cbrt(N) (measured: 6 -> 10 cbrt(N) is worth 12 % on the 241-capsule tree, 3 % on the 41 spheres; the lists
    // get shorter and more cells are clear), bounded by 64 and by 64 Mi (cell, SDF) pairs of set-up work
    int G = (int)std::lround(std::cbrt((double)nt) * 10.0);
    G = std::min(G, (int)std::floor(std::cbrt(64.0 * 1048576.0 / (double)nt)));
    G = std::min(std::max(G, 8), 64);
    double ext = 0;
    for (int a = 0; a < 3; ++a) {
        const double pad = 0.005 * c->gmax[a] + 1e-9;
        c->cull_n[a] = G;
        c->cull_lo[a] = -c->gmax[a] - pad;
        c->cull_cell[a] = 2.0 * (c->gmax[a] + pad) / G;
        ext = std::max(ext, c->gmax[a]);
    }
    const long long ncell = (long long)G * G * G, npair = ncell * nt;
    const double h = 0.5 * std::sqrt(c->cull_cell[0] * c->cull_cell[0] + c->cull_cell[1] * c->cull_cell[1] + c->cull_cell[2] * c->cull_cell[2]) *
                         (1.0 + 1e-6) + 2e-5 * ext;  // slack: FP32 evaluation of positions and distances inside the sweep
    DeviceState& D0 = c->devs[0];
    CU(cudaSetDevice(D0.dev));
    float* dmat = nullptr;
    CU(cudaMalloc(&dmat, sizeof(float) * (size_t)npair));
    KParams P;
    fill_params(c, D0, P);
    cull_eval_kernel<<<(unsigned)std::min<long long>((npair + 255) / 256, 148 * 16), 256, 0, D0.stream>>>(
        P, npair, c->cull_lo[0], c->cull_lo[1], c->cull_lo[2], c->cull_cell[0], c->cull_cell[1], c->cull_cell[2], G, G, dmat);
    cudaError_t ke = cudaGetLastError();
    if (ke != cudaSuccess) { cudaFree(dmat); return set_err("cull_eval_kernel: %s", cudaGetErrorString(ke)); }
    std::vector<float> dc((size_t)npair);
    ke = cudaMemcpyAsync(dc.data(), dmat, sizeof(float) * (size_t)npair, cudaMemcpyDeviceToHost, D0.stream);
    if (ke == cudaSuccess) ke = cudaStreamSynchronize(D0.stream);
    cudaFree(dmat);
    if (ke != cudaSuccess) return set_err("culling grid: %s", cudaGetErrorString(ke));
    c->launches += 1;
    std::vector<int> start((size_t)ncell + 1, 0), items;
    std::vector<float> far((size_t)ncell), clear((size_t)ncell);
    items.reserve((size_t)ncell * 4);
    // Evaluation order inside a cell: by code path (primitive kind / transform class), rarest class first, so that the lanes of a
    // warp -- each walking the list of its own cell -- meet the same kind of primitive at the same loop index (a scene of one
    // box and forty spheres: box first, then spheres).  The maxloc tie rule is kept by an explicit index compare in sweep_one.
    std::vector<int> cls(nt), cls_count(1024, 0);
    for (int j = 0; j < nt; ++j) {
        const DevTop& T = c->tops[j];
        cls[j] = T.mode ? 1023 : std::min(c->prims[T.first].kind * 4 + c->prims[T.first].xf, 1022);
        ++cls_count[cls[j]];
    }
    auto by_class = [&](int a, int b) {
        if (cls[a] != cls[b]) return cls_count[cls[a]] != cls_count[cls[b]] ? cls_count[cls[a]] < cls_count[cls[b]] : cls[a] < cls[b];
        return a < b;
    };
    // Transparent scenes (the space between the bodies neither scatters nor absorbs: sphere.toml) are crossed in steps of `far`, the
    // distance to the nearest UNLISTED surface -- one sweep iteration each.  Listing every SDF that comes within `reach` of the
    // cell makes far >= reach: fewer, longer sweeps.  Measured on sphere.toml (SMCRT_CULL_REACH = 0 / 0.1 / 0.2 / 0.3 / 0.45 of the
    // half extent): 15.2 / 14.8 / 13.6 / 12.4 / 11.0 sweeps per packet but 7.31 / 7.27 / 6.81 / 6.43 / 5.64e8 packets/s -- the
    // longer lists cost more than the saved sweeps (the floor is ~8 sweeps: two per surface crossing).  Off by default.
    double reach = 0.0;
    {
        double kap_max = 0.0;
        for (const DevTop& T : c->tops) kap_max = std::max(kap_max, (double)T.kappa);
        static const char* reach_env = getenv("SMCRT_CULL_REACH");  // fraction of the largest half extent (tuning / A-B switch)
        if (kap_max * ext < 1.0) reach = (reach_env ? atof(reach_env) : 0.0) * ext;
    }
    for (long long cell = 0; cell < ncell; ++cell) {
        const float* d = dc.data() + cell * nt;
        double U = 1e300, M = -1e300;
        for (int j = 0; j < nt; ++j) {
            U = std::min(U, std::fabs((double)d[j]) + h);
            if (d[j] + h < 0) M = std::max(M, (double)d[j] - h);
        }
        double f = 3.0e38;
        for (int j = 0; j < nt; ++j) {
            const bool A = std::fabs((double)d[j]) - h <= std::max(U, reach);
            const bool B = (d[j] - h < 0) && (d[j] + h >= M);
            if (A || B) items.push_back(j);
            else f = std::min(f, std::fabs((double)d[j]) - h);
        }
        std::sort(items.begin() + start[cell], items.end(), by_class);
        start[cell + 1] = (int)items.size();
        far[cell] = (float)f;
        bool inside_some = false;
        for (int j = 0; j < nt; ++j) inside_some = inside_some || d[j] < 0;
        // min_j |dc_j| - h: no surface within that of any point of the cell; cells outside every SDF keep the full sweep (it is what
        // notices that a packet has left the scene, inttau2.f90:143-145)
        clear[cell] = inside_some ? (float)std::max(0.0, (U - 2.0 * h) * (1.0 - 1e-6)) : 0.f;
    }
    c->cull_mean_list = (double)items.size() / (double)ncell;
    if (c->cull_mean_list > 0.6 * nt) return 0;  // nothing to gain on this scene
    for (DeviceState& D : c->devs) {
        CU(cudaSetDevice(D.dev));
        CU(cudaMalloc(&D.cull_start, sizeof(int) * start.size()));
        CU(cudaMalloc(&D.cull_items, sizeof(int) * std::max<size_t>(items.size(), 1)));
        CU(cudaMalloc(&D.cull_far, sizeof(float) * far.size()));
        CU(cudaMemcpy(D.cull_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(D.cull_items, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(D.cull_far, far.data(), sizeof(float) * far.size(), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&D.cull_clear, sizeof(float) * clear.size()));
        CU(cudaMemcpy(D.cull_clear, clear.data(), sizeof(float) * clear.size(), cudaMemcpyHostToDevice));
    }
    c->cull_on = true;
    return 0;
}
static int upload_scene(smcrt_ctx* c) {
    if (!c->scene_dirty) return 0;
    if (c->tops.empty()) return set_err("no scene set (smcrt_set_scene)");
    const int b_prims = align16((int)(c->prims.size() * sizeof(DevPrim)));
    const int b_tops = align16((int)(c->tops.size() * sizeof(DevTop)));
    const int b_prog = align16((int)(c->prog.size() * sizeof(DevInstr)));
    const int b_dets = align16((int)(c->dets.size() * sizeof(DevDet)));
    c->off_tops = b_prims; c->off_prog = b_prims + b_tops; c->off_dets = c->off_prog + b_prog;
    c->off_hot = c->off_dets + b_dets;
    c->off_detp = c->off_hot + align16((int)(c->tops.size() * sizeof(DevHot)));
    c->blob_bytes = c->off_detp + (int)(c->dets.size() * 16);
    std::vector<unsigned char> blob(c->blob_bytes, 0);
    for (size_t j = 0; j < c->dets.size(); ++j) {  // detector planes (n, n.p0): the crossing pre-test of the DETECT site
        const DevDet& D = c->dets[j];
        const float pl[4] = {D.dir[0], D.dir[1], D.dir[2], D.q[13]};
        std::memcpy(blob.data() + c->off_detp + 16 * j, pl, 16);
    }
    {  // the sweep's 32-byte view of every top-level SDF (device_scene.cuh: DevHot)
        DevHot* hot = reinterpret_cast<DevHot*>(blob.data() + c->off_hot);
        for (size_t j = 0; j < c->tops.size(); ++j) {
            const DevTop& T = c->tops[j];
            DevHot h;
            std::memset(&h, 0, sizeof h);
            h.code = HOT_GENERAL;
            h.idx[0] = T.first;
            if (T.mode != 0) { h.code = HOT_PROGRAM; h.idx[1] = T.count; }
            else {
                const DevPrim& Q = c->prims[T.first];
                if (Q.xf <= XF_TRANSLATE && (Q.kind == 1 || Q.kind == 2)) {
                    h.code = Q.kind == 1 ? HOT_SPHERE : HOT_BOX;
                    h.t[0] = h.t[1] = h.t[2] = 0.f;
                    if (Q.xf == XF_TRANSLATE) { h.t[0] = Q.m[3]; h.t[1] = Q.m[7]; h.t[2] = Q.m[11]; }
                    h.p[0] = Q.p[0]; h.p[1] = Q.p[1]; h.p[2] = Q.p[2];
                }
            }
            hot[j] = h;
        }
    }
    std::memcpy(blob.data(), c->prims.data(), c->prims.size() * sizeof(DevPrim));
    std::memcpy(blob.data() + c->off_tops, c->tops.data(), c->tops.size() * sizeof(DevTop));
    if (!c->prog.empty()) std::memcpy(blob.data() + c->off_prog, c->prog.data(), c->prog.size() * sizeof(DevInstr));
    if (!c->dets.empty()) std::memcpy(blob.data() + c->off_dets, c->dets.data(), c->dets.size() * sizeof(DevDet));
    for (DeviceState& D : c->devs) {
        CU(cudaSetDevice(D.dev));
        cudaFree(D.blob); cudaFree(D.primsD); cudaFree(D.progD);
        D.blob = nullptr; D.primsD = nullptr; D.progD = nullptr;
        CU(cudaMalloc(&D.blob, blob.size()));
        CU(cudaMemcpy(D.blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&D.primsD, std::max<size_t>(1, c->primsD.size()) * sizeof(DevPrimD)));
        CU(cudaMemcpy(D.primsD, c->primsD.data(), c->primsD.size() * sizeof(DevPrimD), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&D.progD, std::max<size_t>(1, c->progD.size()) * sizeof(DevInstrD)));
        if (!c->progD.empty()) CU(cudaMemcpy(D.progD, c->progD.data(), c->progD.size() * sizeof(DevInstrD), cudaMemcpyHostToDevice));
    }
    c->scene_dirty = false;
    return build_cull(c);
}

static const unsigned long long HIST_CAP = 1ull << 20;  // tracked hits kept per device between resets (beyond: counted only)
static const int SMEM_BIN_CAP = 8192;  // 64 KB of CTA-private Q40.24 bins at most

static int fill_params(smcrt_ctx* c, DeviceState& D, KParams& P) {
    std::memset(&P, 0, sizeof P);
    P.blob = D.blob; P.blob_bytes = c->blob_bytes;
    P.n_prims = (int)c->prims.size(); P.n_top = (int)c->tops.size(); P.n_instr = (int)c->prog.size(); P.n_det = (int)c->dets.size();
    for (const DevDet& d : c->dets) if (d.kind == SMCRT_DET_CAMERA) P.has_camera = 1;
    for (const DevTop& T : c->tops) if (T.mode == 0 && (c->prims[T.first].kind == 6 || c->prims[T.first].kind == 7)) P.has_capsule = 1;
    P.simple_scene = 1;  // every top-level SDF is a sphere or box the sweep evaluates inline (same test as the DevHot records)
    for (const DevTop& T : c->tops)
        if (T.mode != 0 || c->prims[T.first].xf > XF_TRANSLATE || (c->prims[T.first].kind != 1 && c->prims[T.first].kind != 2)) P.simple_scene = 0;
    P.off_tops = c->off_tops; P.off_prog = c->off_prog; P.off_dets = c->off_dets; P.off_hot = c->off_hot; P.off_detp = c->off_detp;
    P.primsD = D.primsD; P.progD = D.progD;
    P.nxg = c->nxg; P.nyg = c->nyg; P.nzg = c->nzg;
    const int nn[3] = {c->nxg, c->nyg, c->nzg};
    for (int a = 0; a < 3; ++a) {
        P.gmax[a] = (float)c->gmax[a];
        P.vox[a] = (float)(2.0 * c->gmax[a] / nn[a]);
        P.inv_vox[a] = (float)(nn[a] / (2.0 * c->gmax[a]));
        P.hvox[a] = (float)(c->gmax[a] / nn[a]);
    }
    P.src_kind = c->src_kind; P.src_sub = c->src_sub; P.src_alt = c->src_alt;
    P.src_table = D.src_table; P.src_tot = D.src_tot; P.src_id0 = D.src_id0; P.per_src = D.per_src;
    std::memcpy(P.sp, c->sp, sizeof P.sp);
    std::memcpy(P.Tpos, c->Tpos, sizeof P.Tpos);
    std::memcpy(P.Tdir, c->Tdir, sizeof P.Tdir);
    P.jmean = D.jmean; P.absorb = D.absorb; P.emission = D.emission;
    P.det_bins = D.det_bins; P.det_total = (int)c->det_total;
    P.det_in_smem = (c->det_total > 0 && c->det_total <= SMEM_BIN_CAP) ? 1 : 0;
    P.counters = D.counters; P.next = D.counters + C_COUNT;
    if (c->cull_on) {
        P.cull_start = D.cull_start; P.cull_items = D.cull_items; P.cull_far = D.cull_far; P.cull_clear = D.cull_clear;
        // clear cells pay off where packets interact between the bodies: some medium with a mean free path shorter than the grid
        double ext = std::max(c->gmax[0], std::max(c->gmax[1], c->gmax[2]));
        for (const DevTop& T : c->tops) if ((double)T.kappa * ext > 1.0) P.has_capsule = 1;
        for (int a = 0; a < 3; ++a) {
            P.cull_n[a] = c->cull_n[a];
            P.cull_lo[a] = (float)c->cull_lo[a];
            P.cull_inv[a] = (float)(1.0 / c->cull_cell[a]);
        }
    }
    P.dda_legacy = c->dda_legacy ? 1 : 0;
    for (int a = 0; a < 3; ++a) {
        P.jdiff[a] = D.jdiff[a];
        P.jfix[a] = (float)(268435456.0 /* 2^28 */ * nn[a] / (2.0 * c->gmax[a]));
    }
    P.jdiff_used = D.jdiff_used;
    P.seg_buf = D.seg_buf; P.seg_count = D.seg_count; P.seg_total = D.seg_total; P.seg_work = D.seg_count ? D.seg_count + SEG_MAX_SHARES : nullptr;
    static const float piece = getenv("SMCRT_SEG_PIECE") ? std::max((float)atof(getenv("SMCRT_SEG_PIECE")), 1.0f) : 64.0f;  // (experiments)
    P.seg_piece = piece;
    if (c->any_track) { P.hist_ids = D.hist_ids; P.hist_det = D.hist_det; P.hist_n = D.hist_n; P.hist_cap = HIST_CAP; }
    P.eps0 = (float)c->eps0; P.eps_rel = (float)c->eps_rel;
    static const char* wd_env = getenv("SMCRT_WATCHDOG_MS");  // (tests trip the watchdog with a tiny period)
    P.watchdog_ns = wd_env ? (unsigned long long)(atof(wd_env) * 1e6) : 20000000000ull;  // 20 s
    P.max_steps = (int)std::min<long long>(c->max_steps, 1900000ll);  // the compaction step packs sweep count and event index (<= sweeps + 100000 emit retries) into 21 bits each
    return 0;
}

// lane < 0: on the device's main stream, with the whole segment buffer; lane 0 / 1: on that lane's stream, with its half of the
// buffer and its own packet counter
static int launch_kernel(trace_kernel_t kern, const KParams& P0, DeviceState& D, int smem_bytes, bool dry, bool seg_inline, int lane) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SMCRT_BLOCK, smem_bytes));
    if (per_sm < 1) return set_err("scene does not fit in shared memory (%d bytes per CTA)", smem_bytes);
    // persistent grid: every SM full, no more; never more threads than packets
    long long blocks = (long long)D.sm_count * per_sm;
    const long long need = (P0.nphotons + SMCRT_BLOCK - 1) / SMCRT_BLOCK;
    if (blocks > need) blocks = std::max<long long>(need, 1);
    if (dry) return 0;  // the attribute/occupancy queries above have loaded the kernel (lazy module loading)
    KParams P = P0;
    const bool pathlen = (P.tally_mode & SMCRT_TALLY_PATHLENGTH) != 0;
    if (pathlen) {
        if (blocks > SEG_MAX_SHARES) return set_err("trace launch of %lld CTAs exceeds the segment buffer's %d shares", blocks, SEG_MAX_SHARES);
        // seg_cap = 0: every segment is walked by the lane that made it (the fallback of a full share, for all of them)
        P.seg_cap = seg_inline ? 0u : (unsigned int)std::min<size_t>(D.seg_records / (lane < 0 ? 1 : 2) / (size_t)blocks, 0x7fffffffu);
    }
    cudaStream_t st = D.stream;
    if (lane >= 0) {
        st = D.lane_stream[lane];
        P.seg_buf = D.seg_buf + (size_t)lane * (D.seg_records / 2) * 2;  // (2 x float4 per record)
        P.seg_count = D.seg_count + (size_t)lane * (SEG_MAX_SHARES + 1);
        P.seg_work = P.seg_count + SEG_MAX_SHARES;
        P.next = D.counters + C_COUNT + 17 + lane;
        CU(cudaMemsetAsync(P.next, 0, sizeof(unsigned long long), st));
    }
    kern<<<(unsigned)blocks, SMCRT_BLOCK, smem_bytes, st>>>(P);
    CU(cudaGetLastError());
    if (pathlen && !seg_inline) {  // walk what the launch recorded, then clear the shares for the next launch
        // 2 CTAs of 512 threads per SM (64 registers): a third more warps than 3 x 256 at 80 registers, +7 % on sphere.toml / skin
        // 2 CTAs of 512 threads per SM at 64 registers: a third more warps than 3 x 256 at 80 registers (+7 % on sphere.toml and skin)
        constexpr int DEP_THREADS = 512, DEP_CTAS = 2;
        const int dep_smem = (int)sizeof(HotTable) + (DEP_THREADS / 32) * 2 * WARPQ_CAP * 2 * (int)sizeof(float4);  // the table + the warps' two queues
        CU(cudaFuncSetAttribute(deposit_segments_kernel<DEP_THREADS, DEP_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dep_smem));
        deposit_segments_kernel<DEP_THREADS, DEP_CTAS><<<D.sm_count * DEP_CTAS, DEP_THREADS, dep_smem, st>>>(P, (int)blocks);
        clear_segment_counts_kernel<<<1, 256, 0, st>>>(P.seg_count, (int)blocks, P.seg_work);
        CU(cudaGetLastError());
    }
    return 0;
}
struct Variant { int sched; int mb; };
constexpr int NVAR = 6;
static const Variant VARIANTS[NVAR] = {{SCHED_PLAIN, 2}, {SCHED_PLAIN, 3}, {SCHED_PLAIN, 4}, {SCHED_COMPACT, 2}, {SCHED_QUEUED, 2}, {SCHED_QUEUED, 3}};
static int launch_variant(bool pl, bool hd, int var, const KParams& P, DeviceState& D, const int smem_bytes[3], bool dry = false, bool seg_inline = false, int lane = -1) {
    const Variant v = VARIANTS[var];
    const bool need = P.has_capsule != 0, simple = P.simple_scene != 0;
    // LEAN: nothing optional asked of this run (no per-packet records, diagnostics, batched sources, survival biasing)
    const bool lean = !P.out_fate && !P.out_nscatt && !P.out_dbg && !P.dbg_log && !P.src_table && !P.src_tot && !P.survival && !P.out_vert && !P.id_list;
    trace_kernel_t k = pl ? (hd ? pick_kernel_pl1_hd1(v.sched, v.mb, need, simple, lean) : pick_kernel_pl1_hd0(v.sched, v.mb, need, simple, lean))
                          : (hd ? pick_kernel_pl0_hd1(v.sched, v.mb, need, simple, lean) : pick_kernel_pl0_hd0(v.sched, v.mb, need, simple, lean));
    return launch_kernel(k, P, D, smem_bytes[v.sched], dry, seg_inline, lane);
}

// Path-length deposits waiting in the difference grids -> jmean (prefix sums along each touched axis; the grids come back zero).
// Called before anything reads jmean: fetch, the NCCL reduce, and before the fixed-point entries could overflow.
static int scan_pathlength(smcrt_ctx* c, DeviceState& D) {
    if (!D.jdiff_dirty || !D.jdiff[0]) return 0;
    CU(cudaSetDevice(D.dev));
    const int blocks = D.sm_count * 8;
    jdiff_scan_kernel<0><<<blocks, 256, 0, D.stream>>>(D.jdiff[0], D.jmean, D.jdiff_used, c->nxg, c->nyg, c->nzg, 2.0 * c->gmax[0] / c->nxg / 268435456.0);
    jdiff_scan_kernel<1><<<blocks, 256, 0, D.stream>>>(D.jdiff[1], D.jmean, D.jdiff_used, c->nxg, c->nyg, c->nzg, 2.0 * c->gmax[1] / c->nyg / 268435456.0);
    jdiff_scan_kernel<2><<<blocks, 256, 0, D.stream>>>(D.jdiff[2], D.jmean, D.jdiff_used, c->nxg, c->nyg, c->nzg, 2.0 * c->gmax[2] / c->nzg / 268435456.0);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(D.jdiff_used, 0, 16, D.stream));
    c->launches += 3;
    D.jdiff_dirty = false;
    D.jdiff_packets = 0;
    return 0;
}

static int run_on_device(smcrt_ctx* c, DeviceState& D, long long nphotons, uint64_t seed, long long id_offset, int tally_mode,
                         int survival, double threshold, double chance, int* out_fate, int* out_nscatt, int* out_events,
                         float* out_pos, int* out_sweeps = nullptr, float* out_dbg = nullptr) {
    KParams P;
    fill_params(c, D, P);
    P.nphotons = nphotons; P.id_offset = (unsigned long long)id_offset; P.rec_id0 = P.id_offset;
    P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
    P.tally_mode = tally_mode; P.survival = survival ? 1 : 0;
    P.threshold = (float)(threshold > 0 ? threshold : 0.01);  // THRESHOLD, src/constants.f90:28
    P.chance = (float)(chance > 0 ? chance : 0.1);            // CHANCE, src/constants.f90:30
    P.out_fate = out_fate; P.out_nscatt = out_nscatt; P.out_events = out_events; P.out_pos = out_pos; P.out_sweeps = out_sweeps; P.out_dbg = out_dbg;
    P.dbg_pid = c->dbg_pid; P.dbg_log = c->dbg_log; P.dbg_cap = c->dbg_cap;
    if (c->replay_ids) {  // smcrt_history_replay: these packets only, vertex lists out, nothing noted, nothing tallied
        P.id_list = c->replay_ids; P.id_list_n = nphotons; P.out_vert = c->replay_vert; P.out_nvert = c->replay_nvert; P.out_hit = c->replay_hit;
        P.max_vert = c->replay_max_vert; P.hist_n = nullptr; P.det_bins = c->replay_bins; P.det_in_smem = 0;
    }
    CU(cudaSetDevice(D.dev));
    if (tally_mode & SMCRT_TALLY_PATHLENGTH) {
        if (!D.jdiff[0] && !P.dda_legacy) {  // first path-length run on this grid
            size_t nv;
            n_voxels(c, &nv);
            if (!D.jdiff_used) { CU(cudaMalloc(&D.jdiff_used, 16)); CU(cudaMemsetAsync(D.jdiff_used, 0, 16, D.stream)); }
            for (int a = 0; a < 3; ++a) {
                CU(cudaMalloc(&D.jdiff[a], nv * sizeof(long long)));
                CU(cudaMemsetAsync(D.jdiff[a], 0, nv * sizeof(long long), D.stream));
                P.jdiff[a] = D.jdiff[a];
            }
            P.jdiff_used = D.jdiff_used;
        }
        if (!D.seg_buf) {
            // a quarter of the free memory, 32 MB .. 6 GB (a record is 32 bytes; a packet of sphere.toml makes 3, one of skin 40)
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            size_t bytes = std::min<size_t>(std::max<size_t>(free_b / 4, 32ull << 20), 6ull << 30);
            if (const char* e = getenv("SMCRT_SEG_MB")) bytes = std::max<size_t>((size_t)atoll(e), 1) << 20;
            CU(cudaMalloc(&D.seg_buf, bytes));
            D.seg_records = bytes / 32;
            CU(cudaMalloc(&D.seg_count, sizeof(unsigned int) * 2 * (SEG_MAX_SHARES + 1)));  // per lane: the shares' counts + the deposit kernel's work counter
            CU(cudaMemsetAsync(D.seg_count, 0, sizeof(unsigned int) * 2 * (SEG_MAX_SHARES + 1), D.stream));
            for (cudaStream_t& s : D.lane_stream) CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            for (cudaEvent_t* ev : {&D.ev_fork, &D.ev_lane[0], &D.ev_lane[1]}) CU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
            CU(cudaMalloc(&D.seg_total, 8));
            CU(cudaMemsetAsync(D.seg_total, 0, 8, D.stream));
            P.seg_buf = D.seg_buf; P.seg_count = D.seg_count; P.seg_total = D.seg_total; P.seg_work = D.seg_count + SEG_MAX_SHARES;
        }
        // an entry of a difference grid holds < 2^63 for 2^32 full-chord deposits (2^28 units each, weights <= 1/chance)
        if (D.jdiff_packets + nphotons > (1ll << 32)) { int rc = scan_pathlength(c, D); if (rc) return rc; }
        if (!P.dda_legacy) { D.jdiff_dirty = true; D.jdiff_packets += nphotons; }
    }
    static const bool no_ff = getenv("SMCRT_NO_FIRST_FLIGHT") != nullptr;  // A/B switch: every packet takes the ordinary launch sweeps
    if (P.src_kind == SMCRT_SRC_PENCIL && !no_ff) {
        if (!D.ff) CU(cudaMalloc(&D.ff, sizeof(FirstFlight)));
        first_flight_kernel<<<1, 32, 0, D.stream>>>(P, D.ff);
        CU(cudaGetLastError());
        P.ff = D.ff;
    }
    CU(cudaMemsetAsync(P.next, 0, sizeof(unsigned long long), D.stream));
    P.seg_off = (c->blob_bytes + (P.det_in_smem ? (int)c->det_total * 8 : 0) + 15) & ~15;  // the CTA's segment counter (path-length mode)
    const int smem_plain = P.seg_off + 16;
    P.xchg_off = smem_plain;
    const int smem_bytes[3] = {smem_plain, P.xchg_off + 64 + XCHG_WORDS * 4 * SMCRT_BLOCK, P.xchg_off + queued_smem_bytes(SMCRT_BLOCK)};
    CU(cudaEventRecord(D.ev0, D.stream));
    const bool pl = (tally_mode & SMCRT_TALLY_PATHLENGTH) != 0, hd = !c->dets.empty();
    // Which variant wins depends on the scene, so the first large run of a scene spends equal slices of its own packets on the
    // candidates; smcrt_wait reads their device time stamps and later runs use the fastest.  No packet is traced twice: streams
    // depend on (seed, id) only, so a run split into id ranges is the same run.
    static const char* force_mb = getenv("SMCRT_MINBLOCKS_FORCE");
    static const char* force_var = getenv("SMCRT_VARIANT_FORCE");
    const bool compact_ok = c->tops.size() <= 65535 && nphotons >= 4 * SMCRT_BLOCK;  // layer indices share a word in the packet record
    int forced = -1;
    if (force_var) forced = std::min(std::max(atoi(force_var), 0), NVAR - 1);
    else if (c->compact_allowed) forced = 3;  // SMCRT_COMPACT=1 (the compacted kernel)
    else if (force_mb) forced = std::min(std::max(atoi(force_mb), 2), 4) - 2;
    int var = forced >= 0 ? forced : (c->tuned_mb[pl][hd] ? c->tuned_mb[pl][hd] - 1 : 1);
    if (VARIANTS[var].sched != SCHED_PLAIN && !compact_ok) var = 1;
    int n_launch = 0;
    long long chunk_max = std::numeric_limits<long long>::max();
    if (pl) {
        static const char* force_inline = getenv("SMCRT_SEG_INLINE");  // A/B switch: 0 = deposit kernel, 1 = inline walks
        if (force_inline) c->seg_inline = atoi(force_inline) ? 1 : 0;
        if ((c->seg_per_packet <= 0 || (c->seg_inline < 0 && P.nphotons >= (1ll << 23))) && P.nphotons > 0) {
            // First path-length run of this scene: trace a small chunk and count the segments it makes.  A large run also times
            // a second chunk with the segments walked inline: which is faster depends on the scene (the pencil beam of the slab
            // scene makes few, cheap range updates whose atomics hide behind the transport code when issued inline; walking the
            // voxels of refracted rays inline starves the kernel's instruction cache), and later runs of the scene use the winner.
            const bool time_both = c->seg_inline < 0 && P.nphotons >= (1ll << 23);
            const long long n_cal = time_both ? std::min<long long>(std::max<long long>(P.nphotons / 16, 1ll << 20), 1ll << 22) : std::min<long long>(P.nphotons, 1ll << 18);
            { int rc = launch_variant(pl, hd, var, P, D, smem_bytes, true); if (rc) return rc; }  // load the kernel outside the brackets
            CU(cudaMemsetAsync(D.seg_total, 0, 8, D.stream));
            for (int pass = 0; pass < (time_both ? 2 : 1); ++pass) {
                KParams Q = P;
                Q.nphotons = n_cal;
                CU(cudaEventRecord(D.tune_ev[pass], D.stream));
                int rc = launch_variant(pl, hd, var, Q, D, smem_bytes, false, pass == 1 || c->seg_inline == 1);
                if (rc) return rc;
                P.nphotons -= n_cal; P.id_offset += (unsigned long long)n_cal;
                CU(cudaMemsetAsync(P.next, 0, sizeof(unsigned long long), D.stream));
                ++n_launch;
            }
            CU(cudaEventRecord(D.tune_ev[time_both ? 2 : 1], D.stream));
            unsigned long long made = 0;
            CU(cudaMemcpyAsync(&made, D.seg_total, 8, cudaMemcpyDeviceToHost, D.stream));
            CU(cudaStreamSynchronize(D.stream));
            c->seg_per_packet = std::max((double)made / (double)(n_cal * (time_both ? 2 : 1)), 0.25);
            if (time_both) {
                float t_rec = 0, t_inl = 0;
                CU(cudaEventElapsedTime(&t_rec, D.tune_ev[0], D.tune_ev[1]));
                CU(cudaEventElapsedTime(&t_inl, D.tune_ev[1], D.tune_ev[2]));
                c->seg_inline = t_inl < t_rec ? 1 : 0;
            }
        }
        if (c->seg_inline != 1) chunk_max = std::max<long long>(1ll << 16, (long long)((double)D.seg_records / (1.25 * std::max(c->seg_per_packet, 0.25))));
    }
    const long long TUNE_MIN = 8ll << 20;
    if (forced < 0 && compact_ok && !c->tuned_mb[pl][hd] && &D == &c->devs[0] && !out_fate && P.nphotons >= TUNE_MIN && D.tuning < 0) {
        const long long slice = std::min(std::min<long long>(std::max<long long>(nphotons / 64, 1ll << 20), 1ll << 23), chunk_max);
        for (int k = 0; k < NVAR; ++k) {  // load the kernels first: the load would otherwise sit inside the event brackets
            int rc = launch_variant(pl, hd, k, P, D, smem_bytes, true);
            if (rc) return rc;
        }
        CU(cudaMemsetAsync(D.counters + C_COUNT + 1, 0, sizeof(unsigned long long) * 16, D.stream));
        for (int k = 0; k < NVAR; ++k) {
            KParams Q = P;
            Q.nphotons = slice; Q.id_offset = P.id_offset + (unsigned long long)(k * slice);
            Q.tstamp = D.counters + C_COUNT + 1 + 2 * k;
            CU(cudaEventRecord(D.tune_ev[k], D.stream));
            int rc = launch_variant(pl, hd, k, Q, D, smem_bytes, false, c->seg_inline == 1);
            if (rc) return rc;
            CU(cudaMemsetAsync(P.next, 0, sizeof(unsigned long long), D.stream));
        }
        CU(cudaEventRecord(D.tune_ev[NVAR], D.stream));
        D.tuning = (pl ? 2 : 0) | (hd ? 1 : 0);
        P.nphotons -= NVAR * slice; P.id_offset += (unsigned long long)(NVAR * slice);
        n_launch += NVAR;
    }
    // Path-length mode: a launch may produce no more segments than the segment buffer holds (a CTA whose share is full walks its
    // segments inline: correct, but slow), so the run is cut into chunks of packets sized from the scene's measured segments per
    // packet, with a margin for the spread between CTAs.
    // A run of several chunks alternates between the two lanes (chunks of half the buffer, of equal size).  A launch ends with its
    // slowest history -- ~10 ms for a packet of sphere.toml that is reflected 1000 times inside a sphere, whatever the chunk -- and on
    // ONE stream the next trace kernel and the deposit kernel would wait for it with the machine empty.  On two streams the next
    // chunk's CTAs move in as this chunk's CTAs leave, and each lane's deposit kernel follows its trace kernel in stream order.
    long long left = P.nphotons;
    static const bool no_lanes = getenv("SMCRT_PL_LANES") && atoi(getenv("SMCRT_PL_LANES")) == 0;  // A/B switch
    const bool lanes = pl && c->seg_inline != 1 && left > chunk_max && !no_lanes;
    if (lanes) {
        const long long half_max = std::max<long long>(chunk_max / 2, 1ll << 15);
        const long long n_chunks = (left + half_max - 1) / half_max;
        chunk_max = (left + n_chunks - 1) / n_chunks;
        CU(cudaEventRecord(D.ev_fork, D.stream));
        for (cudaStream_t s : D.lane_stream) CU(cudaStreamWaitEvent(s, D.ev_fork, 0));
    }
    for (int k = 0; left > 0; ++k) {
        const long long chunk = std::min(left, chunk_max);
        KParams Q = P;
        Q.nphotons = chunk;
        int rc = launch_variant(pl, hd, var, Q, D, smem_bytes, false, c->seg_inline == 1, lanes ? (k & 1) : -1);
        if (rc) return rc;
        P.id_offset += (unsigned long long)chunk;
        left -= chunk;
        ++n_launch;
        if (left > 0 && !lanes) CU(cudaMemsetAsync(P.next, 0, sizeof(unsigned long long), D.stream));
    }
    if (lanes)
        for (int b = 0; b < 2; ++b) {
            CU(cudaEventRecord(D.ev_lane[b], D.lane_stream[b]));
            CU(cudaStreamWaitEvent(D.stream, D.ev_lane[b], 0));
        }
    CU(cudaEventRecord(D.ev1, D.stream));
    D.ran = true;
    c->launches += (pl ? 3 : 1) * n_launch;
    c->touched_modes |= tally_mode;
    c->dirty_modes |= tally_mode;
    return 0;
}

static int check_ready(smcrt_ctx* c) {
    if (!c) return set_err("null ctx");
    if (c->nxg == 0) return set_err("no grid set (smcrt_set_grid)");
    if (c->devs.empty() || !c->devs[0].det_bins) {
        int rc = smcrt_set_detectors(c, 0, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    return upload_scene(c);
}

extern "C" int smcrt_run_async(smcrt_ctx* c, int64_t nphotons, uint64_t seed, int64_t id_offset, int tally_mode, int survival_bias,
                               double threshold, double chance) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (nphotons < 0) return set_err("smcrt_run: negative nphotons");
    if (c->pending) return set_err("smcrt_run_async: previous run not waited for");
    const int G = (int)c->devs.size();
    for (int g = 0; g < G; ++g) {
        // contiguous id ranges per GPU: [g*N/G, (g+1)*N/G)
        const long long lo = (long long)((__int128)nphotons * g / G), hi = (long long)((__int128)nphotons * (g + 1) / G);
        c->devs[g].ran = false;
        if (hi > lo) {
            rc = run_on_device(c, c->devs[g], hi - lo, seed, id_offset + lo, tally_mode, survival_bias, threshold, chance, nullptr, nullptr,
                               nullptr, nullptr);
            if (rc) return rc;
        }
    }
    c->pending = true;
    return 0;
}
extern "C" int smcrt_wait(smcrt_ctx* c) {
    if (!c) return set_err("null ctx");
    double ms = 0;
    for (DeviceState& D : c->devs) {  // every device first: an error below must not leave another device's run in flight
        CU(cudaSetDevice(D.dev));
        CU(cudaStreamSynchronize(D.stream));
        // (the main stream has joined the chunk lanes of a path-length run; after a launch error it has not)
        for (cudaStream_t ls : D.lane_stream) if (ls) CU(cudaStreamSynchronize(ls));
    }
    int wd_dev = -1;
    for (DeviceState& D : c->devs) {
        CU(cudaSetDevice(D.dev));
        if (D.ran) {
            float t = 0;
            CU(cudaEventElapsedTime(&t, D.ev0, D.ev1));
            ms = std::max(ms, (double)t);
        }
        if (D.ran) {  // trace_queued's watchdog (kernels.cuh): a run that tripped it is incomplete
            unsigned long long wd = 0;
            CU(cudaMemcpy(&wd, D.counters + C_SPARE, sizeof wd, cudaMemcpyDeviceToHost));
            if (wd) {
                CU(cudaMemset(D.counters + C_SPARE, 0, sizeof wd));
                wd_dev = D.dev;
                D.tuning = -1;  // (a trial that tripped it is void)
            }
        }
        if (D.tuning >= 0) {  // the trial slices of run_on_device: keep the fastest kernel variant
            // Ranked by the time from kernel start until the packet pool ran EMPTY (device time stamps): the events around a
            // slice also contain its tail -- the few longest histories finishing alone -- which is the same few milliseconds for
            // every variant and every run length, and would decide a trial of short slices by luck.
            unsigned long long ts[2 * NVAR] = {0};
            CU(cudaMemcpy(ts, D.counters + C_COUNT + 1, sizeof ts, cudaMemcpyDeviceToHost));
            double best = 0;
            int arg = 1;
            for (int k = 0; k < NVAR; ++k) {
                float te = 0;
                CU(cudaEventElapsedTime(&te, D.tune_ev[k], D.tune_ev[k + 1]));
                const double t = ts[2 * k + 1] > ts[2 * k] ? 1e-6 * (double)(ts[2 * k + 1] - ts[2 * k]) : (double)te;
                if (k == 0 || t < best) { best = t; arg = k; }
            }
            c->tuned_mb[(D.tuning >> 1) & 1][D.tuning & 1] = 1 + arg;
            D.tuning = -1;
        }
    }
    if (c->pending) c->last_ms = ms;
    c->pending = false;
    if (wd_dev >= 0)
        return set_err("queue-scheduled kernel: watchdog fired on device %d (a warp found the queues empty for the whole watchdog period); "
                       "this run's tallies are incomplete", wd_dev);
    return 0;
}
extern "C" int smcrt_run(smcrt_ctx* c, int64_t nphotons, uint64_t seed, int64_t id_offset, int tally_mode, int survival_bias,
                         double threshold, double chance) {
    int rc = smcrt_run_async(c, nphotons, seed, id_offset, tally_mode, survival_bias, threshold, chance);
    if (rc) return rc;
    return smcrt_wait(c);
}
extern "C" uint64_t smcrt_last_fetch_bytes(const smcrt_ctx* c) { return c ? c->last_fetch_bytes : 0; }
extern "C" int smcrt_kernel_variant(const smcrt_ctx* c, int tally_mode) {
    if (!c) return -1;
    return c->tuned_mb[(tally_mode & SMCRT_TALLY_PATHLENGTH) ? 1 : 0][c->dets.empty() ? 0 : 1] - 1;
}
extern "C" int smcrt_segment_mode(const smcrt_ctx* c) { return c ? c->seg_inline : -1; }
extern "C" double smcrt_segments_per_packet(const smcrt_ctx* c) { return c ? c->seg_per_packet : 0.0; }
extern "C" double smcrt_last_run_ms(const smcrt_ctx* c) { return c ? c->last_ms : 0.0; }
extern "C" int64_t smcrt_launch_count(const smcrt_ctx* c) { return c ? c->launches : 0; }

// ---- reduce + fetch --------------------------------------------------------------------------------------
// Sum of one float tally grid over the ranks into the root's.  A pencil-beam slab run touches ~2000 of 1.25e8 voxels: when every
// rank's grid has fewer non-zero voxels than the pair scratch holds (1/64 of the grid) the ranks exchange (index, value) pairs --
// one scan at HBM speed, an all-gather of the counts, grouped send/recv of a few KB, a scatter-add on the root -- instead of
// reducing 500 MB over NVLink (1.0-1.6 ms per grid at 8 GPUs, the limiter of a one-step job's scaling).  Dense grids take
// ncclReduce.  `me` lists this process's devices with their ranks in the communicator.
static int reduce_grid(smcrt_ctx* c, float* DeviceState::*grid, int root, int nranks, const std::vector<int>& ranks) {
    size_t nv;
    n_voxels(c, &nv);
    static const bool no_sparse = getenv("SMCRT_NO_SPARSE_REDUCE") != nullptr;
    bool sparse = !no_sparse && nranks > 1;
    for (DeviceState& D : c->devs) sparse = sparse && D.nz_cap > 0;
    std::vector<unsigned long long> counts((size_t)nranks, 0ull);
    if (sparse) {
        for (DeviceState& D : c->devs) {
            CU(cudaSetDevice(D.dev));
            if (!D.nz_counts) CU(cudaMalloc(&D.nz_counts, sizeof(unsigned long long) * 64));
            if (nranks > 64) return set_err("sparse reduce: more than 64 ranks");
            CU(cudaMemsetAsync(D.nz_cursor, 0, 8, D.stream));
            nnz_pack_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(D.*grid, (long long)nv, D.nz_idx, D.nz_val, D.nz_cursor, (unsigned long long)D.nz_cap);
            CU(cudaGetLastError());
        }
        NC(nccl::GroupStart());
        for (DeviceState& D : c->devs) {
            CU(cudaSetDevice(D.dev));
            NC(nccl::AllGather(D.nz_cursor, D.nz_counts, 1, nccl::ncclUint64, D.comm, D.stream));
        }
        NC(nccl::GroupEnd());
        DeviceState& D0 = c->devs[0];
        CU(cudaSetDevice(D0.dev));
        CU(cudaMemcpyAsync(counts.data(), D0.nz_counts, sizeof(unsigned long long) * nranks, cudaMemcpyDeviceToHost, D0.stream));
        CU(cudaStreamSynchronize(D0.stream));
        for (int r = 0; r < nranks; ++r) sparse = sparse && counts[r] <= c->devs[0].nz_cap;  // (the scan gives up beyond the cap)
    }
    if (!sparse) {
        NC(nccl::GroupStart());
        for (DeviceState& D : c->devs) {
            CU(cudaSetDevice(D.dev));
            NC(nccl::Reduce(D.*grid, D.*grid, nv, nccl::ncclFloat32, nccl::ncclSum, root, D.comm, D.stream));
        }
        NC(nccl::GroupEnd());
        return 0;
    }
    size_t total = 0;
    for (int r = 0; r < nranks; ++r) if (r != root) total += counts[r];
    DeviceState* R = nullptr;
    for (size_t g = 0; g < c->devs.size(); ++g) if (ranks[g] == root) R = &c->devs[g];
    if (R && total > R->rx_cap) {
        CU(cudaSetDevice(R->dev));
        cudaFree(R->rx_idx); cudaFree(R->rx_val);
        R->rx_cap = std::max<size_t>(total, 4096) * 2;
        CU(cudaMalloc(&R->rx_idx, R->rx_cap * 4));
        CU(cudaMalloc(&R->rx_val, R->rx_cap * 4));
    }
    NC(nccl::GroupStart());
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceState& D = c->devs[g];
        CU(cudaSetDevice(D.dev));
        if (ranks[g] == root) {
            size_t off = 0;
            for (int r = 0; r < nranks; ++r) {
                if (r == root || !counts[r]) continue;
                NC(nccl::Recv(D.rx_idx + off, counts[r], nccl::ncclUint32, r, D.comm, D.stream));
                NC(nccl::Recv(D.rx_val + off, counts[r], nccl::ncclFloat32, r, D.comm, D.stream));
                off += counts[r];
            }
        } else if (counts[ranks[g]]) {
            NC(nccl::Send(D.nz_idx, counts[ranks[g]], nccl::ncclUint32, root, D.comm, D.stream));
            NC(nccl::Send(D.nz_val, counts[ranks[g]], nccl::ncclFloat32, root, D.comm, D.stream));
        }
    }
    NC(nccl::GroupEnd());
    if (R && total) {
        CU(cudaSetDevice(R->dev));
        scatter_add_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, R->stream>>>(R->*grid, R->rx_idx, R->rx_val, (long long)total);
        CU(cudaGetLastError());
    }
    return 0;
}
static int reduce_buffers(smcrt_ctx* c, int root_rank_or_dev) {
    // root receives in place; only the grids a run could have written (every rank runs the same modes, so the collectives match up)
    for (DeviceState& D : c->devs) { int rc = scan_pathlength(c, D); if (rc) return rc; }
    std::vector<int> ranks(c->devs.size());
    for (size_t g = 0; g < c->devs.size(); ++g) ranks[g] = c->comm_rank ? c->rank : (int)g;
    const int nranks = c->comm_rank ? c->nranks : (int)c->devs.size();
    int rc = 0;
    if ((c->touched_modes & SMCRT_TALLY_PATHLENGTH) && (rc = reduce_grid(c, &DeviceState::jmean, root_rank_or_dev, nranks, ranks))) return rc;
    if ((c->touched_modes & SMCRT_TALLY_ABSORB) && (rc = reduce_grid(c, &DeviceState::absorb, root_rank_or_dev, nranks, ranks))) return rc;
    if ((c->touched_modes & SMCRT_TALLY_EMISSION) && (rc = reduce_grid(c, &DeviceState::emission, root_rank_or_dev, nranks, ranks))) return rc;
    NC(nccl::GroupStart());
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceState& D = c->devs[g];
        CU(cudaSetDevice(D.dev));
        NC(nccl::Reduce(D.det_bins, D.det_bins, (size_t)std::max<long long>(c->det_total, 1), nccl::ncclUint64, nccl::ncclSum,
                        root_rank_or_dev, D.comm, D.stream));
        NC(nccl::Reduce(D.counters, D.counters, (size_t)C_COUNT, nccl::ncclUint64, nccl::ncclSum, root_rank_or_dev, D.comm, D.stream));
    }
    NC(nccl::GroupEnd());
    for (DeviceState& D : c->devs) {
        CU(cudaSetDevice(D.dev));
        CU(cudaStreamSynchronize(D.stream));
    }
    return 0;
}
// zero a large host array with a few threads (a 500 MB grid: ~10 ms instead of ~50 ms)
static void parallel_zero(float* p, size_t n) {
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < (1u << 22) || hw == 1) { std::memset(p, 0, n * sizeof(float)); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + hw - 1) / hw;
    for (unsigned t = 0; t < hw; ++t) {
        const size_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        if (hi > lo) th.emplace_back([=] { std::memset(p + lo, 0, (hi - lo) * sizeof(float)); });
    }
    for (std::thread& t : th) t.join();
}

static int zero_device_tallies(smcrt_ctx* c, DeviceState& D) {
    size_t nv;
    n_voxels(c, &nv);
    CU(cudaSetDevice(D.dev));
    if (D.jdiff_dirty && D.jdiff[0]) {  // deposits waiting in the difference grids are dropped with the rest
        for (int a = 0; a < 3; ++a) CU(cudaMemsetAsync(D.jdiff[a], 0, nv * sizeof(long long), D.stream));
        CU(cudaMemsetAsync(D.jdiff_used, 0, 16, D.stream));
        D.jdiff_dirty = false; D.jdiff_packets = 0;
    }
    // (only the grids a run has written since they were last cleared: a 500^3 grid is 500 MB)
    if (c->dirty_modes & SMCRT_TALLY_PATHLENGTH) CU(cudaMemsetAsync(D.jmean, 0, nv * 4, D.stream));
    if (c->dirty_modes & SMCRT_TALLY_ABSORB) CU(cudaMemsetAsync(D.absorb, 0, nv * 4, D.stream));
    if (c->dirty_modes & SMCRT_TALLY_EMISSION) CU(cudaMemsetAsync(D.emission, 0, nv * 4, D.stream));
    CU(cudaMemsetAsync(D.det_bins, 0, sizeof(unsigned long long) * (size_t)std::max<long long>(c->det_total, 1), D.stream));
    CU(cudaMemsetAsync(D.counters, 0, sizeof(unsigned long long) * C_COUNT, D.stream));
    if (D.hist_n) CU(cudaMemsetAsync(D.hist_n, 0, 8, D.stream));
    CU(cudaStreamSynchronize(D.stream));
    return 0;
}

extern "C" int smcrt_comm_unique_id(char id_out[128]) {
    if (nccl::load()) return -1;
    nccl::ncclUniqueId id;
    NC(nccl::GetUniqueId(&id));
    std::memcpy(id_out, id.internal, 128);
    return 0;
}
extern "C" int smcrt_comm_init(smcrt_ctx* c, int nranks, int rank, const char id[128]) {
    if (!c) return set_err("null ctx");
    if (c->devs.size() != 1) return set_err("smcrt_comm_init: rank mode needs a context with exactly one GPU");
    if (nccl::load()) return -1;
    nccl::ncclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    CU(cudaSetDevice(c->devs[0].dev));
    NC(nccl::CommInitRank(&c->devs[0].comm, nranks, uid, rank));
    c->comm_rank = true; c->nranks = nranks; c->rank = rank;
    return 0;
}
extern "C" int smcrt_comm_reduce(smcrt_ctx* c, int root) {
    if (!c) return set_err("null ctx");
    if (!c->comm_rank) return set_err("smcrt_comm_reduce: smcrt_comm_init was not called");
    int rc = check_ready(c);
    if (rc) return rc;
    rc = reduce_buffers(c, root);
    if (rc) return rc;
    // non-root ranks have handed their tallies over: clear them so a later accumulate does not double count
    if (c->rank != root) return zero_device_tallies(c, c->devs[0]);
    return 0;
}

extern "C" int smcrt_fetch(smcrt_ctx* c, float* jmean, float* absorb, float* emission, double* det_bins, smcrt_counters* counters,
                           int accumulate) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (c->pending && (rc = smcrt_wait(c))) return rc;
    const int G = (int)c->devs.size();
    if (G > 1) {
        if (!c->comm_all) {
            if (nccl::load()) return -1;
            std::vector<nccl::ncclComm_t> comms(G);
            std::vector<int> ids(G);
            for (int g = 0; g < G; ++g) ids[g] = c->devs[g].dev;
            NC(nccl::CommInitAll(comms.data(), G, ids.data()));
            for (int g = 0; g < G; ++g) c->devs[g].comm = comms[g];
            c->comm_all = true;
        }
        if ((rc = reduce_buffers(c, 0))) return rc;
        for (int g = 1; g < G; ++g)
            if ((rc = zero_device_tallies(c, c->devs[g]))) return rc;
    }
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    if ((rc = scan_pathlength(c, D))) return rc;
    size_t nv;
    n_voxels(c, &nv);
    std::vector<float> tmp;
    c->last_fetch_bytes = 0;
    const bool no_sparse = getenv("SMCRT_NO_SPARSE_FETCH") != nullptr;
    auto pull = [&](float* host, const float* dev) -> int {
        if (!host) return 0;
        // sparse path: scan the grid on the device (one HBM pass), copy only the non-zero voxels
        if (!no_sparse && D.nz_cap) {
            const size_t cap = D.nz_cap;  // scratch allocated by smcrt_set_grid
            CU(cudaMemsetAsync(D.nz_cursor, 0, 8, D.stream));
            nnz_pack_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(dev, (long long)nv, D.nz_idx, D.nz_val, D.nz_cursor, (unsigned long long)cap);
            CU(cudaGetLastError());
            unsigned long long nnz = 0;
            CU(cudaMemcpyAsync(&nnz, D.nz_cursor, 8, cudaMemcpyDeviceToHost, D.stream));
            CU(cudaStreamSynchronize(D.stream));
            c->launches += 1;
            c->last_fetch_bytes += 8;
            if (nnz <= cap) {
                if (nnz) {
                    CU(cudaMemcpyAsync(D.nz_idx_h, D.nz_idx, nnz * 4, cudaMemcpyDeviceToHost, D.stream));
                    CU(cudaMemcpyAsync(D.nz_val_h, D.nz_val, nnz * 4, cudaMemcpyDeviceToHost, D.stream));
                }
                if (!accumulate) parallel_zero(host, nv);  // overlaps the copies
                CU(cudaStreamSynchronize(D.stream));
                if (accumulate) for (size_t k = 0; k < nnz; ++k) host[D.nz_idx_h[k]] += D.nz_val_h[k];
                else for (size_t k = 0; k < nnz; ++k) host[D.nz_idx_h[k]] = D.nz_val_h[k];
                c->last_fetch_bytes += nnz * 8;
                return 0;
            }
        }
        c->last_fetch_bytes += nv * 4;
        if (!accumulate) {
            CU(cudaMemcpy(host, dev, nv * 4, cudaMemcpyDeviceToHost));
        } else {
            tmp.resize(nv);
            CU(cudaMemcpy(tmp.data(), dev, nv * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nv; ++i) host[i] += tmp[i];
        }
        return 0;
    };
    if ((rc = pull(jmean, D.jmean)) || (rc = pull(absorb, D.absorb)) || (rc = pull(emission, D.emission))) return rc;
    if (det_bins && c->det_total > 0) {
        std::vector<unsigned long long> raw((size_t)c->det_total);
        CU(cudaMemcpy(raw.data(), D.det_bins, raw.size() * 8, cudaMemcpyDeviceToHost));
        c->last_fetch_bytes += raw.size() * 8;
        for (size_t i = 0; i < raw.size(); ++i) {
            const double v = (double)raw[i] / 16777216.0;
            det_bins[i] = accumulate ? det_bins[i] + v : v;
        }
    }
    if (counters) {
        unsigned long long raw[C_COUNT];
        CU(cudaMemcpy(raw, D.counters, sizeof raw, cudaMemcpyDeviceToHost));
        c->last_fetch_bytes += sizeof raw;
        smcrt_counters k{};
        k.nscatt = (double)raw[C_NSCATT];
        k.sweeps = (double)raw[C_SWEEPS];
        k.sdf_evals = (double)raw[C_SWEEPS] * (double)c->tops.size();
        k.bounces = (double)raw[C_BOUNCES];
        k.launched = (double)raw[C_LAUNCHED];
        k.emit_retries = (double)raw[C_RETRIES];
        k.lost = (double)raw[C_LOST];
        k.det_hits = (double)raw[C_DETHITS];
        k.voxel_crossings = (double)raw[C_VOXELS];
        k.deposit_atomics = (double)raw[C_REDS];
        if (accumulate) {
            counters->nscatt += k.nscatt; counters->sdf_evals += k.sdf_evals; counters->bounces += k.bounces;
            counters->launched += k.launched; counters->emit_retries += k.emit_retries; counters->lost += k.lost;
            counters->sweeps += k.sweeps; counters->det_hits += k.det_hits;
            counters->voxel_crossings += k.voxel_crossings; counters->deposit_atomics += k.deposit_atomics;
        } else
            *counters = k;
    }
    return 0;
}

extern "C" int smcrt_reset_tallies(smcrt_ctx* c) {
    int rc = check_ready(c);
    if (rc) return rc;
    c->touched_modes = 0;
    for (DeviceState& D : c->devs)
        if ((rc = zero_device_tallies(c, D))) return rc;
    c->dirty_modes = 0;
    return 0;
}

extern "C" int smcrt_pin_host(void* ptr, uint64_t bytes) {
    if (!ptr || !bytes) return set_err("smcrt_pin_host: null buffer");
    CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
extern "C" int smcrt_unpin_host(void* ptr) {
    CU(cudaHostUnregister(ptr));
    return 0;
}

// ---- probes ----------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)) == cudaSuccess ? 0 : -1; }
    template <typename T> T* as() { return (T*)p; }
};
std::vector<float> to_f(const double* d, size_t n) {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = (float)d[i];
    return f;
}
int up_f(DevBuf& b, const double* d, size_t n) {
    std::vector<float> f = to_f(d, n);
    if (b.alloc(n * 4)) return -1;
    return cudaMemcpy(b.p, f.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}
int down_f(DevBuf& b, double* d, size_t n) {
    std::vector<float> f(n);
    if (cudaMemcpy(f.data(), b.p, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (size_t i = 0; i < n; ++i) d[i] = (double)f[i];
    return 0;
}
int grid_for(long long n) { return (int)std::min<long long>(std::max<long long>((n + 255) / 256, 1), 148 * 8); }
}  // namespace
#define PROBE_FAIL() set_err("probe: CUDA error: %s", cudaGetErrorString(cudaGetLastError()))

extern "C" int smcrt_probe_sdf(smcrt_ctx* c, int top_index, int64_t n, const double* pos, double* dist, double* normal) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (top_index < 0 || top_index > (int)c->tops.size()) return set_err("smcrt_probe_sdf: top_index out of range");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    KParams P;
    fill_params(c, D, P);
    const size_t nd = top_index > 0 ? (size_t)n : (size_t)n * c->tops.size();
    DevBuf bp, bd, bn;
    if (up_f(bp, pos, 3 * (size_t)n) || bd.alloc(nd * 4) || bn.alloc(3 * (size_t)n * 4)) return PROBE_FAIL();
    const bool want_n = normal && top_index > 0;
    CU(cudaFuncSetAttribute(probe_sdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->blob_bytes));
    probe_sdf_kernel<<<grid_for(n), 256, c->blob_bytes, D.stream>>>(P, top_index, n, bp.as<float>(), bd.as<float>(),
                                                                     want_n ? bn.as<float>() : nullptr);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    if (down_f(bd, dist, nd)) return PROBE_FAIL();
    if (want_n && down_f(bn, normal, 3 * (size_t)n)) return PROBE_FAIL();
    return 0;
}
extern "C" int smcrt_probe_ray(smcrt_ctx* c, int top_index, int64_t n, const double* pos, const double* dir, double* dist, double* bound,
                               int32_t* exact) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (top_index < 1 || top_index > (int)c->tops.size() || !pos || !dir || !dist || !bound) return set_err("smcrt_probe_ray: invalid arguments");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    KParams P;
    fill_params(c, D, P);
    DevBuf bp, bv, bd, bb, be;
    if (up_f(bp, pos, 3 * (size_t)n) || up_f(bv, dir, 3 * (size_t)n) || bd.alloc(n * 4) || bb.alloc(n * 4) || be.alloc(n * 4)) return PROBE_FAIL();
    CU(cudaFuncSetAttribute(probe_ray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->blob_bytes));
    probe_ray_kernel<<<grid_for(n), 256, c->blob_bytes, D.stream>>>(P, top_index, n, bp.as<float>(), bv.as<float>(), bd.as<float>(),
                                                                     bb.as<float>(), be.as<int>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    if (down_f(bd, dist, n) || down_f(bb, bound, n)) return PROBE_FAIL();
    if (exact) CU(cudaMemcpy(exact, be.p, 4 * n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int smcrt_probe_fresnel(smcrt_ctx* c, int64_t n, const double* dir, const double* nrm, const double* n1, const double* n2,
                                   const double* xi, double* dir_out, double* refl_coeff, int32_t* rflag) {
    if (!c) return set_err("null ctx");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    DevBuf a, b, e, f, g, o, r, fl;
    if (up_f(a, dir, 3 * n) || up_f(b, nrm, 3 * n) || up_f(e, n1, n) || up_f(f, n2, n) || up_f(g, xi, n) || o.alloc(12 * n) ||
        r.alloc(4 * n) || fl.alloc(4 * n))
        return PROBE_FAIL();
    probe_fresnel_kernel<<<grid_for(n), 256, 0, D.stream>>>(n, a.as<float>(), b.as<float>(), e.as<float>(), f.as<float>(), g.as<float>(),
                                                            o.as<float>(), r.as<float>(), fl.as<int>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    if (down_f(o, dir_out, 3 * n)) return PROBE_FAIL();
    if (refl_coeff && down_f(r, refl_coeff, n)) return PROBE_FAIL();
    if (rflag) CU(cudaMemcpy(rflag, fl.p, 4 * n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int smcrt_probe_scatter(smcrt_ctx* c, int64_t n, const double* dir, const double* hgg, const double* xi, double* dir_out) {
    if (!c) return set_err("null ctx");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    DevBuf a, b, e, o;
    if (up_f(a, dir, 3 * n) || up_f(b, hgg, n) || up_f(e, xi, 2 * n) || o.alloc(12 * n)) return PROBE_FAIL();
    probe_scatter_kernel<<<grid_for(n), 256, 0, D.stream>>>(n, a.as<float>(), b.as<float>(), e.as<float>(), o.as<float>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    if (down_f(o, dir_out, 3 * n)) return PROBE_FAIL();
    return 0;
}
extern "C" int smcrt_probe_emit(smcrt_ctx* c, int64_t n, const double* xi4, double* pos, double* dir, int32_t* cell) {
    if (!c) return set_err("null ctx");
    if (c->nxg == 0) return set_err("no grid set (smcrt_set_grid)");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    KParams P;
    fill_params(c, D, P);
    DevBuf a, p, d, ce;
    if (up_f(a, xi4, 4 * n) || p.alloc(12 * n) || d.alloc(12 * n) || ce.alloc(12 * n)) return PROBE_FAIL();
    probe_emit_kernel<<<grid_for(n), 256, 0, D.stream>>>(P, n, a.as<float>(), p.as<float>(), d.as<float>(), ce.as<int>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    if (down_f(p, pos, 3 * n) || down_f(d, dir, 3 * n)) return PROBE_FAIL();
    if (cell) CU(cudaMemcpy(cell, ce.p, 12 * n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int smcrt_probe_detector(smcrt_ctx* c, int det_index, int64_t n, const double* start, const double* dir, const double* seg_len,
                                    int32_t* hit, int32_t* bin) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (det_index < 1 || det_index > (int)c->dets.size()) return set_err("smcrt_probe_detector: det_index out of range");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    KParams P;
    fill_params(c, D, P);
    DevBuf a, b, e, h, bi;
    if (up_f(a, start, 3 * n) || up_f(b, dir, 3 * n) || up_f(e, seg_len, n) || h.alloc(4 * n) || bi.alloc(4 * n)) return PROBE_FAIL();
    probe_detector_kernel<<<grid_for(n), 256, 0, D.stream>>>(P, det_index, n, a.as<float>(), b.as<float>(), e.as<float>(), h.as<int>(),
                                                             bi.as<int>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    c->launches += 1;
    CU(cudaMemcpy(hit, h.p, 4 * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(bin, bi.p, 4 * n, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int smcrt_trace_packets(smcrt_ctx* c, int64_t n, uint64_t seed, int64_t id_offset, int tally_mode, int survival_bias,
                                   int32_t* fate, int32_t* nscatt, double* final_pos, int32_t* n_events, int32_t* n_sweeps) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (c->pending) return set_err("smcrt_trace_packets: a run is pending");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    DevBuf f, s, e, p, w, dbg;
    if (f.alloc(4 * n) || s.alloc(4 * n) || e.alloc(4 * n) || p.alloc(12 * n) || w.alloc(4 * n)) return PROBE_FAIL();
    const char* dbg_pid_s = getenv("SMCRT_DEBUG_PID");  // engine diagnostics: boundary-event log of one packet
    DevBuf dlog;
    const int dcap = 256;
    if (dbg_pid_s && (dlog.alloc(64 * dcap) || cudaMemset(dlog.p, 0, 64 * dcap) != cudaSuccess)) return PROBE_FAIL();
    c->dbg_pid = dbg_pid_s ? atoll(dbg_pid_s) : -1; c->dbg_log = dbg_pid_s ? dlog.as<float>() : nullptr; c->dbg_cap = dcap;
    const char* dbg_path = getenv("SMCRT_DEBUG_LOST");  // engine diagnostics: dump the state of step-capped packets
    if (dbg_path && (dbg.alloc(48 * n) || cudaMemset(dbg.p, 0, 48 * n) != cudaSuccess)) return PROBE_FAIL();
    CU(cudaMemset(f.p, 0xff, 4 * n));
    CU(cudaMemset(s.p, 0, 4 * n));  // the kernel increments the scatter count in place
    rc = run_on_device(c, D, n, seed, id_offset, tally_mode, survival_bias, -1, -1, f.as<int>(), s.as<int>(), e.as<int>(), p.as<float>(), w.as<int>(), dbg_path ? dbg.as<float>() : nullptr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(D.stream));
    float t = 0;
    CU(cudaEventElapsedTime(&t, D.ev0, D.ev1));
    c->last_ms = t;
    if (fate) CU(cudaMemcpy(fate, f.p, 4 * n, cudaMemcpyDeviceToHost));
    if (nscatt) CU(cudaMemcpy(nscatt, s.p, 4 * n, cudaMemcpyDeviceToHost));
    if (n_events) CU(cudaMemcpy(n_events, e.p, 4 * n, cudaMemcpyDeviceToHost));
    if (n_sweeps) CU(cudaMemcpy(n_sweeps, w.p, 4 * n, cudaMemcpyDeviceToHost));
    if (dbg_pid_s) {
        std::vector<float> h(16 * dcap);
        CU(cudaMemcpy(h.data(), dlog.p, 64 * dcap, cudaMemcpyDeviceToHost));
        for (int k = 0; k < dcap; ++k)
            if (h[16 * k + 1] != 0.f) {
                fprintf(stderr, "ev=%d", k + 1);
                for (int q = 0; q < 16; ++q) fprintf(stderr, " %.9g", h[16 * k + q]);
                fprintf(stderr, "\n");
            }
        c->dbg_log = nullptr;
    }
    if (dbg_path) {
        std::vector<float> h(12 * (size_t)n);
        std::vector<int> hf((size_t)n);
        CU(cudaMemcpy(h.data(), dbg.p, 48 * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hf.data(), f.p, 4 * n, cudaMemcpyDeviceToHost));
        if (FILE* fp = fopen(dbg_path, "a")) {
            for (int64_t i = 0; i < n; ++i)
                if (hf[i] == 3 && h[12 * i + 9] != 0.f) {
                    fprintf(fp, "id=%lld", (long long)(id_offset + i));
                    for (int k = 0; k < 12; ++k) fprintf(fp, " %.9g", h[12 * i + k]);
                    fprintf(fp, "\n");
                }
            fclose(fp);
        }
    }
    if (final_pos && down_f(p, final_pos, 3 * n)) return PROBE_FAIL();
    return 0;
}
// Batched point sources: the body of the escape-function drivers (cart/cyl_calc_escape_sym, src/kernelsMod.f90:533-642,
// 959-1071) for MANY grid cells in one launch.  Per source the reference does: layer = maxloc(d, mask=d<0) at the position;
// layer == 0 or kappa(layer) == 0 -> escape = 0 without running; else run_MCRT with an isotropic point source there and
// escape(det) = total_dect / nphotons.  Here every active source owns nphotons_per_source consecutive packet ids.
extern "C" int smcrt_run_sources(smcrt_ctx* c, int64_t n_src, const double* pos, int64_t nphotons_per_source, uint64_t seed,
                                 int64_t id_offset, int tally_mode, int survival_bias, double threshold, double chance,
                                 double* det_totals, int32_t* layer_out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (n_src < 1 || !pos || nphotons_per_source < 1 || !det_totals) return set_err("smcrt_run_sources: invalid arguments");
    if (c->pending) return set_err("smcrt_run_sources: a run is pending");
    const int n_top = (int)c->tops.size(), n_det = (int)c->dets.size();
    if (n_det < 1) return set_err("smcrt_run_sources: no detectors set");
    // layer of every source position (FP64 SDF evaluation on the device, same code as smcrt_probe_sdf)
    std::vector<double> dist((size_t)n_src * n_top);
    rc = smcrt_probe_sdf(c, 0, n_src, pos, dist.data(), nullptr);
    if (rc) return rc;
    std::vector<int64_t> active;
    std::vector<float> table;
    for (int64_t i = 0; i < n_src; ++i) {
        int layer = 0;
        double best = 0.0;
        for (int t = 0; t < n_top; ++t) {  // maxloc(distances, mask = distances < 0): first maximum wins
            const double d = dist[(size_t)i * n_top + t];
            if (d < 0.0 && (layer == 0 || d > best)) { layer = t + 1; best = d; }
        }
        if (layer_out) layer_out[i] = layer;
        bool inside = true;
        for (int a = 0; a < 3; ++a) inside = inside && pos[3 * i + a] >= -c->gmax[a] && pos[3 * i + a] < c->gmax[a];
        if (layer != 0 && inside && c->tops[layer - 1].kappa != 0.f) {
            active.push_back(i);
            for (int a = 0; a < 3; ++a) table.push_back((float)pos[3 * i + a]);
        }
    }
    std::fill(det_totals, det_totals + (size_t)n_src * n_det, 0.0);
    if (active.empty()) return 0;
    // isotropic point emitter; the position comes from the table
    const int keep_kind = c->src_kind, keep_sub = c->src_sub, keep_alt = c->src_alt;
    c->src_kind = SMCRT_SRC_POINT;
    const int G = (int)c->devs.size();
    const long long n_act = (long long)active.size(), total = n_act * nphotons_per_source;
    std::vector<DevBuf> tab(G), tot(G);
    int err = 0;
    for (int g = 0; g < G && !err; ++g) {
        DeviceState& D = c->devs[g];
        if (cudaSetDevice(D.dev) != cudaSuccess || tab[g].alloc(table.size() * 4) || tot[g].alloc((size_t)n_act * n_det * 8) ||
            cudaMemcpyAsync(tab[g].p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, D.stream) != cudaSuccess ||
            cudaMemsetAsync(tot[g].p, 0, (size_t)n_act * n_det * 8, D.stream) != cudaSuccess) {
            err = set_err("smcrt_run_sources: device buffers: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        D.src_table = tab[g].as<float>(); D.src_tot = tot[g].as<unsigned long long>();
        D.src_id0 = (unsigned long long)id_offset; D.per_src = nphotons_per_source;
        const long long lo = (long long)((__int128)total * g / G), hi = (long long)((__int128)total * (g + 1) / G);
        D.ran = false;
        if (hi > lo) err = run_on_device(c, D, hi - lo, seed, id_offset + lo, tally_mode, survival_bias, threshold, chance, nullptr, nullptr, nullptr, nullptr);
    }
    c->pending = true;
    const int wrc = smcrt_wait(c);
    std::vector<unsigned long long> h((size_t)n_act * n_det);
    for (int g = 0; g < G; ++g) {
        DeviceState& D = c->devs[g];
        if (!err && !wrc && D.src_tot) {
            if (cudaSetDevice(D.dev) != cudaSuccess || cudaMemcpy(h.data(), D.src_tot, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
                err = set_err("smcrt_run_sources: download: %s", cudaGetErrorString(cudaGetLastError()));
            else
                for (long long k = 0; k < n_act; ++k)
                    for (int d = 0; d < n_det; ++d) det_totals[(size_t)active[k] * n_det + d] += (double)h[(size_t)k * n_det + d] / 16777216.0;
        }
        D.src_table = nullptr; D.src_tot = nullptr; D.per_src = 0; D.src_id0 = 0;
    }
    c->src_kind = keep_kind; c->src_sub = keep_sub; c->src_alt = keep_alt;
    return err ? err : wrc;
}
// ---- trackHistory (src/historyStack.f90, src/detectors/detector_base.f90:157-162) --------------------------------------------
extern "C" int smcrt_set_track_history(smcrt_ctx* c, int n, const int32_t* track) {
    if (!c) return set_err("null ctx");
    if (n != (int)c->dets.size() || (n > 0 && !track)) return set_err("smcrt_set_track_history: %d flags for %d detectors", n, (int)c->dets.size());
    c->any_track = false;
    for (int i = 0; i < n; ++i) {
        c->track[(size_t)i] = track[i] ? 1 : 0;
        c->dets[(size_t)i].pad_ = track[i] ? 1 : 0;
        c->any_track = c->any_track || track[i];
    }
    c->scene_dirty = true;  // the detector records live in the scene blob
    if (c->any_track)
        for (DeviceState& D : c->devs) {
            CU(cudaSetDevice(D.dev));
            if (D.hist_n) continue;
            CU(cudaMalloc(&D.hist_ids, HIST_CAP * 8));
            CU(cudaMalloc(&D.hist_det, HIST_CAP * 4));
            CU(cudaMalloc(&D.hist_n, 8));
            CU(cudaMemset(D.hist_n, 0, 8));
        }
    return 0;
}
extern "C" int smcrt_history_hits(smcrt_ctx* c, int64_t max_hits, uint64_t* ids, int32_t* det, int64_t* total) {
    if (!c || !total) return set_err("smcrt_history_hits: null argument");
    int rc = check_ready(c);
    if (rc) return rc;
    if (c->pending && (rc = smcrt_wait(c))) return rc;
    std::vector<std::pair<uint64_t, int32_t>> all;
    int64_t seen = 0;
    for (DeviceState& D : c->devs) {
        if (!D.hist_n) continue;
        CU(cudaSetDevice(D.dev));
        unsigned long long n = 0;
        CU(cudaMemcpy(&n, D.hist_n, 8, cudaMemcpyDeviceToHost));
        seen += (int64_t)n;
        const size_t k = (size_t)std::min<unsigned long long>(n, HIST_CAP);
        std::vector<unsigned long long> hi(k);
        std::vector<int> hd(k);
        if (k) {
            CU(cudaMemcpy(hi.data(), D.hist_ids, k * 8, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(hd.data(), D.hist_det, k * 4, cudaMemcpyDeviceToHost));
        }
        for (size_t i = 0; i < k; ++i) all.emplace_back((uint64_t)hi[i], (int32_t)hd[i] + 1);  // 1-based detector index
    }
    std::sort(all.begin(), all.end());  // by packet id: the order does not depend on the schedule
    *total = seen;
    for (int64_t i = 0; i < std::min<int64_t>(max_hits, (int64_t)all.size()); ++i) {
        if (ids) ids[i] = all[(size_t)i].first;
        if (det) det[i] = all[(size_t)i].second;
    }
    return 0;
}
extern "C" int smcrt_history_replay(smcrt_ctx* c, int64_t n, const uint64_t* ids, uint64_t seed, int survival_bias, int max_vertices,
                                    float* vertices, int32_t* n_vertices, int32_t* hit_vertex) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (n < 1 || !ids || max_vertices < 2 || !vertices || !n_vertices) return set_err("smcrt_history_replay: invalid arguments");
    if (c->pending) return set_err("smcrt_history_replay: a run is pending");
    std::vector<unsigned long long> sorted(ids, ids + n);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    const int64_t m = (int64_t)sorted.size();
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    DevBuf bi, bv, bn, bh, bb;
    if (bi.alloc((size_t)m * 8) || bv.alloc((size_t)m * max_vertices * 16) || bn.alloc((size_t)m * 4) || bh.alloc((size_t)m * 4) ||
        bb.alloc((size_t)std::max<long long>(c->det_total, 1) * 8))
        return PROBE_FAIL();
    CU(cudaMemcpy(bi.p, sorted.data(), (size_t)m * 8, cudaMemcpyHostToDevice));
    CU(cudaMemset(bn.p, 0, (size_t)m * 4));
    CU(cudaMemset(bh.p, 0xff, (size_t)m * 4));
    c->replay_ids = bi.as<unsigned long long>(); c->replay_vert = bv.as<float4>(); c->replay_nvert = bn.as<int>(); c->replay_hit = bh.as<int>();
    c->replay_bins = bb.as<unsigned long long>(); c->replay_max_vert = max_vertices;
    const int keep_touched = c->touched_modes, keep_dirty = c->dirty_modes;
    DevBuf bc;  // the run's counters are put back afterwards: a replay is not part of the job
    if (bc.alloc(sizeof(unsigned long long) * C_COUNT)) return PROBE_FAIL();
    CU(cudaMemcpyAsync(bc.p, D.counters, sizeof(unsigned long long) * C_COUNT, cudaMemcpyDeviceToDevice, D.stream));
    rc = run_on_device(c, D, m, seed, 0, 0 /* no voxel tallies */, survival_bias, -1, -1, nullptr, nullptr, nullptr, nullptr);
    c->replay_ids = nullptr; c->replay_vert = nullptr; c->replay_nvert = nullptr; c->replay_hit = nullptr; c->replay_bins = nullptr;
    c->touched_modes = keep_touched; c->dirty_modes = keep_dirty;
    if (rc) return rc;
    CU(cudaStreamSynchronize(D.stream));
    CU(cudaMemcpy(D.counters, bc.p, sizeof(unsigned long long) * C_COUNT, cudaMemcpyDeviceToDevice));
    // back in the caller's order
    std::vector<float> hv((size_t)m * max_vertices * 4);
    std::vector<int> hn((size_t)m), hh((size_t)m);
    CU(cudaMemcpy(hv.data(), bv.p, hv.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hn.data(), bn.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hh.data(), bh.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
        const size_t k = (size_t)(std::lower_bound(sorted.begin(), sorted.end(), (unsigned long long)ids[i]) - sorted.begin());
        std::memcpy(vertices + (size_t)i * max_vertices * 4, hv.data() + k * (size_t)max_vertices * 4, (size_t)max_vertices * 16);
        n_vertices[i] = hn[k];
        if (hit_vertex) hit_vertex[i] = hh[k];
    }
    return 0;
}
extern "C" int smcrt_inverse_mcrt(smcrt_ctx* c, int top_index, int find_mask, const double* bounds, int max_steps, int64_t nphotons,
                                  uint64_t seed, int tally_mode, const double* targets, double* table, int* best_step) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (top_index < 1 || top_index > (int)c->tops.size()) return set_err("smcrt_inverse_mcrt: top_index %d out of range", top_index);
    if (!(find_mask & 15)) return set_err("Please select at least one of mus, mua, hgg, n to find with inverse MCRT");  // :1577
    if (max_steps < 1 || nphotons < 1 || !targets || !table) return set_err("smcrt_inverse_mcrt: invalid arguments");
    const int n_det = (int)c->hdets.size();
    int n_target = 0;
    for (int d = 0; d < n_det; ++d) n_target += targets[d] != -1.0;
    if (!n_target) return set_err("smcrt_inverse_mcrt: no detector has a target value");
    static const double ref_bounds[8] = {0.0, 100.0, 0.0, 100.0, -1.0, 1.0, 1.0, 20.0};  // :1604-1611
    const double* B = bounds ? bounds : ref_bounds;
    const int t = top_index - 1;
    const double keep[4] = {c->opt_mus[t], c->opt_mua[t], c->opt_hgg[t], c->opt_n[t]};
    uint64_t s = seed ^ 0x696E7665727365ull;  // trial points: SplitMix64 (the reference's ran2 stream is unpinned)
    auto uni = [&]() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    };
    std::vector<double> bins((size_t)std::max<long long>(c->det_total, 1));
    int best = 0;
    for (int k = 0; k < max_steps && !rc; ++k) {
        double v[4];
        for (int q = 0; q < 4; ++q) v[q] = (find_mask >> q) & 1 ? uni() * (B[2 * q + 1] - B[2 * q]) + B[2 * q] : keep[q];
        if ((rc = smcrt_set_optprops(c, top_index, v[0], v[1], v[2], v[3]))) break;
        if ((rc = smcrt_reset_tallies(c))) break;                                   // reset(dects), :1646
        if ((rc = smcrt_run(c, nphotons, seed + (uint64_t)k, 0, tally_mode, 0, -1.0, -1.0))) break;
        if ((rc = smcrt_fetch(c, nullptr, nullptr, nullptr, bins.data(), nullptr, 0))) break;
        double e = 0.0;                                                              // inverse_evaluate, :1753-1787
        for (int d = 0; d < n_det; ++d) {
            if (targets[d] == -1.0) continue;
            double total = 0.0;
            for (long long b = 0; b < c->hdets[d].count; ++b) total += bins[(size_t)(c->hdets[d].offset + b)];
            e += std::fabs(total / (double)nphotons - targets[d]);
        }
        e = -e / n_target;
        for (int q = 0; q < 4; ++q) table[5 * k + q] = v[q];
        table[5 * k + 4] = e;
        if (e > table[5 * best + 4]) best = k;
    }
    const int rc2 = smcrt_set_optprops(c, top_index, keep[0], keep[1], keep[2], keep[3]);
    if (!rc) smcrt_reset_tallies(c);
    if (best_step) *best_step = best;
    return rc ? rc : rc2;
}
extern "C" int smcrt_bench_red(smcrt_ctx* c, int pattern, int span, int64_t n_ops, double* ops_per_s) {
    if (!c || !ops_per_s) return set_err("smcrt_bench_red: null argument");
    if (c->nxg == 0) return set_err("smcrt_bench_red: no grid set (smcrt_set_grid)");
    if (pattern < 0 || pattern > 2 || span < 1 || n_ops < 1) return set_err("smcrt_bench_red: invalid arguments");
    if (c->pending) return set_err("smcrt_bench_red: a run is pending");
    DeviceState& D = c->devs[0];
    CU(cudaSetDevice(D.dev));
    const long long nvox = (long long)c->nxg * c->nyg * c->nzg, stride = (long long)c->nxg * c->nyg;
    if (span > c->nzg) span = c->nzg;
    if (nvox <= span) return set_err("smcrt_bench_red: grid too small");
    const int threads = 256, blocks = D.sm_count * 8;
    const int per_thread = (int)std::max<long long>(1, n_ops / ((long long)threads * blocks));
    red_bench_kernel<<<blocks, threads, 0, D.stream>>>(D.jmean, nvox, stride, span, pattern, std::min(per_thread, 64));  // warm-up
    CU(cudaEventRecord(D.ev0, D.stream));
    red_bench_kernel<<<blocks, threads, 0, D.stream>>>(D.jmean, nvox, stride, span, pattern, per_thread);
    CU(cudaEventRecord(D.ev1, D.stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, D.ev0, D.ev1));
    *ops_per_s = (double)per_thread * threads * blocks / (ms * 1e-3);
    CU(cudaMemsetAsync(D.jmean, 0, sizeof(float) * nvox, D.stream));  // the benchmark scribbles on the path-length grid
    CU(cudaStreamSynchronize(D.stream));
    return 0;
}
extern "C" int smcrt_probe_philox(uint64_t seed, uint64_t packet_id, uint32_t event, uint32_t out[4]) {
    philox4x32_10(event, (uint32_t)packet_id, (uint32_t)(packet_id >> 32), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), out);
    return 0;
}
