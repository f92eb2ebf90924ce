/*
 * smcrt.h — C ABI of libsmcrt_gpu.so, the B200 (sm_100a) photon-packet transport engine that
 * replaces the body of signedMCRT/RSMCRT's `run_MCRT` (reference: src/kernelsMod.f90:1790-1898).
 *
 * The reference has no FFI for this path: its hot path is reached through ordinary Fortran module
 * calls and communicates through module-global state.  The de-facto operator boundary is the
 * `run_MCRT` dummy-argument list + the module globals it touches; every entry point
 * below cites the reference state it carries across.  All signatures are plain pointers / sizes so a
 * Fortran 2018 host binds them with ISO_C_BINDING (see INTEGRATION.md for the `bind(C)` interface
 * block and the `select type` flattener a maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; `smcrt_last_error()` gives the message.
 *     (the reference's convention is `error stop`; the Fortran shim maps non-zero to `error stop`).
 *   - all host arrays are owned by the caller; the library copies during the call and never retains
 *     pointers.  Host precision is real64 (reference `wp`, src/constants.f90:18); the f64 -> f32 drop
 *     happens inside smcrt_set_*.
 *   - indices that name a top-level SDF ("layer") are 1-based like the reference; 0 = "outside all".
 *   - single-threaded entry is assumed (the reference calls run_MCRT from serial context).
 */
#ifndef SMCRT_H
#define SMCRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smcrt_ctx smcrt_ctx; /* opaque */

/* ---- SDF node kinds (reference: src/sdfs/sdfs.f90:494-735, sdfModifiers.f90:286-491) -------- */
enum smcrt_node_kind {
    SMCRT_SPHERE = 1,     /* params: r                                  sdfs.f90:494-508 */
    SMCRT_BOX = 2,        /* params: hx,hy,hz (HALF lengths as stored)   sdfs.f90:510-525 */
    SMCRT_TORUS = 3,      /* params: oradius, iradius                    sdfs.f90:527-542 */
    SMCRT_CYLINDER = 4,   /* params: a(3), b(3), radius                  sdfs.f90:544-581 */
    SMCRT_TRIPRISM = 5,   /* params: h1, h2                              sdfs.f90:583-597 */
    SMCRT_SEGMENT = 6,    /* params: a(3), b(3)   (radius fixed 0.1)     sdfs.f90:599-626 */
    SMCRT_CAPSULE = 7,    /* params: a(3), b(3), r                       sdfs.f90:628-648 */
    SMCRT_CONE = 8,       /* params: a(3), b(3), ra, rb                  sdfs.f90:650-686 */
    SMCRT_EGG = 9,        /* params: r1, r2, h                           sdfs.f90:688-718 */
    SMCRT_PLANE = 10,     /* params: a(3) unit normal                    sdfs.f90:720-735 */
    /* CSG `model`: left fold of op over children (sdf_base.f90:146-161); params: k */
    SMCRT_MODEL_UNION = 20,        /* sdfModifiers.f90:428-441 */
    SMCRT_MODEL_SMOOTHUNION = 21,  /* sdfModifiers.f90:443-459 */
    SMCRT_MODEL_SUBTRACTION = 22,  /* sdfModifiers.f90:461-476 */
    SMCRT_MODEL_INTERSECTION = 23, /* sdfModifiers.f90:478-491 */
    /* single-child modifiers; their own transform is identity and unused (sdfModifiers.f90:180-280) */
    SMCRT_MOD_REVOLUTION = 30, /* params: o, cx,cy,cz    :303-321 */
    SMCRT_MOD_EXTRUDE = 31,    /* params: h              :286-301 */
    SMCRT_MOD_ONION = 32,      /* params: thickness      :323-333 */
    SMCRT_MOD_TWIST = 33,      /* params: k              :353-371 */
    SMCRT_MOD_BEND = 34,       /* params: k              :373-391 */
    SMCRT_MOD_ELONGATE = 35    /* params: sx,sy,sz       :335-351 */
    /* `displacement` carries a Fortran procedure pointer and `repeat` is `error stop` in the
       reference (:393-426); neither can cross a C ABI and no shipped scene uses them. */
};
#define SMCRT_NODE_PARAMS 8

/* ---- photon sources (reference: src/photon.f90:214-1043) ---------------------------------- */
enum smcrt_source_kind {
    SMCRT_SRC_POINT = 1,    /* photon.f90:311-359 */
    SMCRT_SRC_PENCIL = 2,   /* photon.f90:652-710 */
    SMCRT_SRC_UNIFORM = 3,  /* photon.f90:566-649 */
    SMCRT_SRC_CIRCULAR = 4, /* photon.f90:214-308 */
    SMCRT_SRC_FOCUS = 5,    /* photon.f90:361-563 */
    SMCRT_SRC_ANNULUS = 6,  /* photon.f90:850-1043 */
    SMCRT_SRC_DSLIT = 7,    /* photon.f90:712-780  double slit, hard-coded geometry in units of the wavelength */
    SMCRT_SRC_APERTURE = 8  /* photon.f90:782-848  square aperture, idem */
};
/* source parameter block, 24 doubles (what `photon_origin` + the emitters' dict keys hold) */
enum smcrt_source_slot {
    SMCRT_SP_POS = 0,       /* [0..2]  photon_origin%pos            photon.f90:88-97 */
    SMCRT_SP_DIR = 3,       /* [3..5]  photon_origin%n{x,y,z}p */
    SMCRT_SP_P1 = 6,        /* [6..8]  dict pos1%x..z  (uniform)    photon.f90:596-606 */
    SMCRT_SP_P2 = 9,        /* [9..11] dict pos2 */
    SMCRT_SP_P3 = 12,       /* [12..14] dict pos3 */
    SMCRT_SP_RADIUS = 15,   /* dict radius (circular); dslit / aperture: this%wavelength (constant spectrum, parse_spectrum.f90:52-117) */
    SMCRT_SP_FOCAL = 16,    /* dict focalLength */
    SMCRT_SP_BEAM = 17,     /* dict beam_size */
    SMCRT_SP_RLO = 18,      /* dict rlo */
    SMCRT_SP_RHI = 19,      /* dict rhi */
    SMCRT_SP_SIGMA = 20,    /* dict sigma */
    SMCRT_SP_ROT = 21       /* [21..23] dict rotation%x..z (already normalised by the parser) */
};
#define SMCRT_SOURCE_PARAMS 24
/* `subtype` of smcrt_set_source: focus_type / annulus_type strings of the reference */
enum smcrt_source_subtype {
    SMCRT_FOCUS_SQUARE = 1, SMCRT_FOCUS_CIRCLE = 2, SMCRT_FOCUS_GAUSSIAN = 3,
    SMCRT_ANNULUS_TOPHAT = 1, SMCRT_ANNULUS_BESSEL = 2, SMCRT_ANNULUS_GAUSSIAN = 3
};

/* ---- detectors (reference: src/detectors/detectors.f90) ------------------------------------ */
enum smcrt_detector_kind {
    SMCRT_DET_CIRCLE = 1,  /* p: pos(3) dir(3) radius                                   :107-164 */
    SMCRT_DET_ANNULUS = 2, /* p: pos(3) dir(3) r1 r2                                    :166-244 */
    SMCRT_DET_FIBRE = 3,   /* p: pos(3) dir(3) f1 f2 f1Ap f2Ap frontOff backOff frontToPin
                                 pinToBack pinAp acceptAngle coreDiameter               :246-393 */
    SMCRT_DET_CAMERA = 4   /* p: p1(3) p2(3) p3(3) maxval                               :395-469 */
};
#define SMCRT_DET_PARAMS 20

/* ---- tally selection: the reference's compile-time flags become run-time bits -------------- */
enum smcrt_tally_mode {
    SMCRT_TALLY_ABSORB = 1,     /* recordWeight, kernelsMod.f90:2202-2220 (always on in the reference) */
    SMCRT_TALLY_PATHLENGTH = 2, /* -Dpathlength: update_grids DDA, inttau2.f90:408-445 */
    SMCRT_TALLY_EMISSION = 4    /* state%render_source: recordEmissionLocation, kernelsMod.f90:2184-2200 */
};

/* per-run counters the reference keeps on the packet but never prints (photon.f90:48) plus engine health */
typedef struct smcrt_counters {
    double nscatt;        /* Σ scatter events            (kernelsMod.f90:1965, reduction(+:nscatt)) */
    double sdf_evals;     /* Σ packet%cnts               (inttau2.f90:67,83,138,183,219,232) */
    double bounces;       /* Σ packet%bounces            (inttau2.f90:311) */
    double launched;      /* packets launched (== nphotons) */
    double emit_retries;  /* re-emissions of the start-voxel rejection loop (kernelsMod.f90:1939-1943) */
    double lost;          /* packets killed by an engine guard (step cap / bounces>1000 / layer 0) */
    double sweeps;        /* eval-all sweeps executed (engine instrumentation; == sdf_evals / n_top) */
    double det_hits;      /* detector hits recorded */
    double voxel_crossings; /* -Dpathlength: voxels the straight segments crossed = deposits of the reference's loop (inttau2.f90:417-441) */
    double deposit_atomics; /* -Dpathlength: atomics the engine issued for them (range updates cover many voxels with four) */
} smcrt_counters;

/* ---- life cycle ---------------------------------------------------------------------------- */
/* n_gpus devices driven from this one process (the reference's OpenMP thread count becomes a GPU count,
   kernelsMod.f90:1833-1836).  device_ids may be NULL (= 0..n_gpus-1); n_gpus==0 means "all visible". */
int smcrt_create(smcrt_ctx** out, int n_gpus, const int* device_ids);
void smcrt_destroy(smcrt_ctx* ctx);
const char* smcrt_last_error(void);
/* returns e.g. "smcrt-b200 0.1 sm_100a" */
const char* smcrt_version(void);

/* state%grid = init_grid_cart(n*, *max)  (src/grid.f90:119-159, parse.f90:110).  Allocates (and zeroes)
   the tally grids like alloc_array/zarray (src/setup.f90:29-30, :170-190). */
int smcrt_set_grid(smcrt_ctx* ctx, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax);

/* array(:) of type(sdf), flattened.  Node i: kind[i]; children (models: n_child>=1,
   modifiers: exactly 1) are the CONTIGUOUS nodes first_child[i] .. first_child[i]+n_child[i]-1;
   xform = n_nodes x 16, each 4x4 in Fortran column-major order exactly as `sdf_base%transform` is stored
   (row-vector convention p' = p.M, translation in row 4: vector_class.f90:292-304);
   params = n_nodes x SMCRT_NODE_PARAMS.  top_node[t] is the root node of top-level SDF t+1 ("layer" t+1,
   order matters: maxloc(ds, mask=ds<0), inttau2.f90:84,220);  optics per top-level SDF as `mono(mus,mua,hgg,n)`
   (opticalProperties.f90:107-125; kappa/albedo/g2 are derived inside with the same rule). */
int smcrt_set_scene(smcrt_ctx* ctx, int n_nodes, const int32_t* kind, const int32_t* first_child,
                    const int32_t* n_child, const double* xform, const double* params, int n_top,
                    const int32_t* top_node, const double* mus, const double* mua, const double* hgg,
                    const double* n_ref);
/* sdf%updateOptProp (sdf_base.f90:244-255), used by inverse_MCRT between runs. top_index is 1-based. */
int smcrt_set_optprops(smcrt_ctx* ctx, int top_index, double mus, double mua, double hgg, double n_ref);

/* packet = photon(name) + set_photon + dict keys (parse_source.f90:17-264).  p: SMCRT_SOURCE_PARAMS doubles. */
int smcrt_set_source(smcrt_ctx* ctx, int kind, int subtype, const double* p);

/* dects(:)  (parse_detectors.f90:17-141).  p = n x SMCRT_DET_PARAMS, nbins = the USER nbins (the stored
   count is nbins+1, detectors.f90:133; cameras (nbins+1)^2).  Tallies are zeroed. */
int smcrt_set_detectors(smcrt_ctx* ctx, int n, const int32_t* kind, const double* p, const int32_t* nbins);
/* number of doubles smcrt_fetch writes to det_bins (Σ stored bins, detector order) */
int64_t smcrt_det_bins_total(const smcrt_ctx* ctx);

/* engine knobs without a reference counterpart.  eps0: absolute floor of the boundary tolerance
   (reference eps = 1e-8, inttau2.f90:56); the engine uses max(eps0, eps_rel*|pos|_inf) because the
   arithmetic is FP32.  <=0 keeps the default.  max_steps: per-packet sweep cap. */
int smcrt_set_tolerances(smcrt_ctx* ctx, double eps0, double eps_rel, int64_t max_steps);

/* ---- run: the photon loop of run_MCRT (kernelsMod.f90:1861-1888) ---------------------------- */
/* Launches packets with global ids [id_offset, id_offset+nphotons); the Philox stream of a packet depends
   only on (seed, global id), so a job split over ranks/GPUs is the same job.  tally_mode: OR of
   smcrt_tally_mode.  survival_bias!=0 selects survivalBiasPropagation (kernelsMod.f90:1979-2067) with
   THRESHOLD/CHANCE (constants.f90:28-30; pass <=0 for the reference values 0.01 / 0.1).
   Tallies ACCUMULATE across calls, like the module arrays do across run_MCRT calls. Blocking. */
int smcrt_run(smcrt_ctx* ctx, int64_t nphotons, uint64_t seed, int64_t id_offset, int tally_mode,
              int survival_bias, double threshold, double chance);
/* non-blocking variant + wait, so a host can overlap (used by the benchmark to time on the device) */
int smcrt_run_async(smcrt_ctx* ctx, int64_t nphotons, uint64_t seed, int64_t id_offset, int tally_mode,
                    int survival_bias, double threshold, double chance);
int smcrt_wait(smcrt_ctx* ctx);
/* device time of the last completed run in ms (CUDA events on the launch stream; max over this ctx's GPUs) */
double smcrt_last_run_ms(const smcrt_ctx* ctx);
/* kernels launched by this ctx since creation (the benchmark's gpu_launches claim) */
int64_t smcrt_launch_count(const smcrt_ctx* ctx);

/* ---- results -------------------------------------------------------------------------------- */
/* Sums the tallies of all GPUs of this ctx (NCCL reduce to the first device when n_gpus>1; the intent of the
   dead mpi_reduce block, kernelsMod.f90:2351-2357) and copies them out.  Grids: nxg*nyg*nzg floats, x fastest
   (Fortran (nxg,nyg,nzg) order, setup.f90:180-183); any pointer may be NULL.  accumulate!=0: host += device
   (the reference's `jmeanGLOBAL = jmean` after in-place accumulation), else overwrite. */
int smcrt_fetch(smcrt_ctx* ctx, float* jmean, float* absorb, float* emission, double* det_bins,
                smcrt_counters* counters, int accumulate);
/* Device -> host bytes the last smcrt_fetch moved.  A grid whose non-zero voxels are few (< 1/64 of the grid; e.g. a pencil
 * beam in a wide slab) is read back as (index, value) pairs after one scan on the device instead of as the whole array;
 * the result in the caller's array is identical.  SMCRT_NO_SPARSE_FETCH=1 disables it. */
uint64_t smcrt_last_fetch_bytes(const smcrt_ctx* ctx);
/* zarray + detector reset (setup.f90:192-205, kernelsMod.f90:2418-2439) */
int smcrt_reset_tallies(smcrt_ctx* ctx);
/* Page-lock (cudaHostRegister) / release a host buffer that smcrt_fetch will be given repeatedly, e.g. the module arrays
   jmean/absorb/emission of src/iarray.f90: device->host copies into pinned memory run at full PCIe/C2C bandwidth. Optional. */
int smcrt_pin_host(void* ptr, uint64_t bytes);
int smcrt_unpin_host(void* ptr);

/* ---- multi-process (one rank per GPU) reduce, NCCL over NVLink ------------------------------ */
/* rank 0 calls smcrt_comm_unique_id (128 bytes), the host broadcasts it (MPI / torch.distributed / file),
   every rank calls smcrt_comm_init.  smcrt_comm_reduce then sums the tally grids, detector bins and
   counters of all ranks onto rank `root` with ONE grouped ncclReduce per buffer. */
int smcrt_comm_unique_id(char id_out[128]);
int smcrt_comm_init(smcrt_ctx* ctx, int nranks, int rank, const char id[128]);
int smcrt_comm_reduce(smcrt_ctx* ctx, int root);

/* ---- deterministic-component probes (parity tests) ------------------------ */
/* Evaluate top-level SDF `top_index` (1-based; 0 = all: out has n*n_top values, SDF fastest) at n points
   (pos = n x 3 doubles) with the engine's FP32 device code.  normal (n x 3, may be NULL) is only written
   for a single SDF (replaces calcNormal, sdf_base.f90:166-190). */
int smcrt_probe_sdf(smcrt_ctx* ctx, int top_index, int64_t n, const double* pos, double* dist, double* normal);
/* The directional step bound of top-level SDF `top_index` (engine-specific; replaces the `min |d|` step of inttau2.f90:155-176
 * for bodies with a closed-form ray intersection): dist = signed distance at pos, bound >= |dist| = how far the packet may move
 * along dir without crossing this body's surface (1e30 = never), exact = 1 when bound IS the along-ray distance to the surface. */
int smcrt_probe_ray(smcrt_ctx* ctx, int top_index, int64_t n, const double* pos, const double* dir, double* dist, double* bound,
                    int32_t* exact);
/* reflect_refract with a supplied uniform (surfaces.f90:14-127): in dir/nrm n x 3, n1,n2,xi n; out new dir,
   R (Fresnel coefficient), rflag. */
int smcrt_probe_fresnel(smcrt_ctx* ctx, int64_t n, const double* dir, const double* nrm, const double* n1,
                        const double* n2, const double* xi, double* dir_out, double* refl_coeff, int32_t* rflag);
/* photon%scatter with supplied uniforms (photon.f90:1045-1103): xi = n x 2 (cos-theta draw, phi draw). */
int smcrt_probe_scatter(smcrt_ctx* ctx, int64_t n, const double* dir, const double* hgg, const double* xi,
                        double* dir_out);
/* emit with supplied uniforms (4 per packet) using the source set by smcrt_set_source; out pos/dir n x 3,
   cell n x 3 (1-based voxel, -1 outside: grid.f90:51-78). */
int smcrt_probe_emit(smcrt_ctx* ctx, int64_t n, const double* xi4, double* pos, double* dir, int32_t* cell);
/* record_hit on one straight segment per entry against detector det_index (1-based), tallies untouched:
   out hit flag and the bin index (1-based; cameras: idx + (idy-1)*nbinsX). */
int smcrt_probe_detector(smcrt_ctx* ctx, int det_index, int64_t n, const double* start, const double* dir,
                         const double* seg_len, int32_t* hit, int32_t* bin);
/* Trace packets [id_offset, id_offset+n) exactly like smcrt_run (tallies ARE updated) and also return, per
   packet: fate (0 absorbed, 1 left geometry/grid, 2 roulette, 3 lost), scatter count, final position, RNG events
   consumed (lost packets: minus the engine-guard code) and sweeps executed.  Any out pointer may be NULL. */
int smcrt_trace_packets(smcrt_ctx* ctx, int64_t n, uint64_t seed, int64_t id_offset, int tally_mode,
                        int survival_bias, int32_t* fate, int32_t* nscatt, double* final_pos,
                        int32_t* n_events, int32_t* n_sweeps);
/* Which kernel variant the engine settled on for this scene and tally mode after timing the candidates on slices of the first
 * large run (>= 8 Mi packets): 0..2 = one packet per thread at 2/3/4 resident CTAs per SM (128/80/64 registers), 3 = the same
 * with CTA-level event compaction (2 CTAs), 4..5 = queue-scheduled (packets in shared-memory slots, warps pull batches of one
 * state) at 2/3 CTAs; -1 = not timed yet (variant 1 is used).  Environment override: SMCRT_VARIANT_FORCE=0..5. */
int smcrt_kernel_variant(const smcrt_ctx* ctx, int tally_mode);
/* -Dpathlength runs (update_grids, inttau2.f90:408-445): how the straight segments of the current scene are deposited.
 * 0 = the trace kernels record them and a deposit kernel walks the voxels, 1 = the trace kernels walk them inline, -1 = not timed
 * yet (the first large path-length run of a scene times both on slices of its own packets).  And the measured number of straight
 * segments per packet, which sizes the launches against the segment buffer (0 = not measured yet). */
int smcrt_segment_mode(const smcrt_ctx* ctx);
double smcrt_segments_per_packet(const smcrt_ctx* ctx);
/* Batched point sources in ONE launch: the body of the escape-function drivers cart_calc_escape_sym / cyl_calc_escape_sym
 * (src/kernelsMod.f90:533-642, 959-1071), which call run_MCRT once per symmetry-grid cell with set_photon(cell centre).
 * pos: n_src x 3 emission points (already rotated/shifted by the caller, :576-585).  Per source, as in the reference:
 * layer = maxloc(d, mask = d < 0) (:592-596, returned in layer_out if non-NULL); layer == 0 or kappa(layer) == 0 -> no packets,
 * totals 0 (:599-609); else nphotons_per_source packets from an isotropic point source (photon.f90:311-359).
 * det_totals: n_src x n_det (source-major), the sum of each detector's bins (total_dect); escapeSymmetry = total / nphotons (:617).
 * Voxel tallies accumulate over all sources, like repeated run_MCRT calls.  Packet ids: source k of the ACTIVE list owns
 * [id_offset + k*nphotons_per_source, +nphotons_per_source).  Blocking. */
int smcrt_run_sources(smcrt_ctx* ctx, int64_t n_src, const double* pos, int64_t nphotons_per_source, uint64_t seed,
                      int64_t id_offset, int tally_mode, int survival_bias, double threshold, double chance,
                      double* det_totals, int32_t* layer_out);
/* trackHistory (src/historyStack.f90:89-108,184-226; [[detectors]] trackHistory, parse_detectors.f90:175-182).  The reference
 * pushes (position, step) at the launch and at every interaction point (kernelsMod.f90:1954,1959) and writes the list when the
 * packet hits a detector that tracks histories (detector_base.f90:157-160) -- serial builds only ("incompatable with OpenMP").
 * Here the transport kernels only NOTE which packet hit which tracking detector (16 bytes per hit); because a packet's random
 * stream depends on (seed, packet id) alone, its vertex list is then produced by tracing just those packets AGAIN with recording
 * switched on: no per-packet vertex storage for the 1e8 packets that hit nothing.
 *   smcrt_set_track_history   one flag per detector of the current table (a new smcrt_set_detectors clears them)
 *   smcrt_history_hits        (packet id, 1-based detector) of up to max_hits hits since the last reset, sorted by id; *total = all
 *                             hits seen (2^20 are kept per GPU)
 *   smcrt_history_replay      vertices: n x max_vertices x 4 floats (x, y, z, event index; the launch point first, every interaction
 *                             point, then the hit point on the detector plane with the detector's 1-based index as 4th component);
 *                             n_vertices: how many each packet produced up to the END of its history (may exceed max_vertices:
 *                             the rest is not stored); hit_vertex: the vertex count at its first tracked hit (-1 = none).  seed and
 *                             survival_bias must be those of the run.  Tallies and counters of the context are left untouched. */
int smcrt_set_track_history(smcrt_ctx* ctx, int n_det, const int32_t* track);
int smcrt_history_hits(smcrt_ctx* ctx, int64_t max_hits, uint64_t* packet_ids, int32_t* det_index, int64_t* total);
int smcrt_history_replay(smcrt_ctx* ctx, int64_t n, const uint64_t* packet_ids, uint64_t seed, int survival_bias, int max_vertices,
                         float* vertices, int32_t* n_vertices, int32_t* hit_vertex);
/* inverse_MCRT (src/kernelsMod.f90:1462-1751): search for the optical properties of ONE top-level SDF that make the detectors read
 * their target values.  The reference's loop is the skeleton of a LIPO search: every step draws a trial point uniformly inside
 * fixed bounds for the properties being sought (mus, mua in [0,100], g in [-1,1], n in [1,20]; :1604-1611), runs run_MCRT, and
 * scores it with inverse_evaluate = -mean_i |total_i / nphotons - target_i| over the detectors whose target is not -1
 * (:1753-1787); detectors are reset between steps (:1646).  Same here, in one call:
 *   top_index   1-based index into array(:) (the reference looks it up from [inverse] layer, :1582-1592)
 *   find_mask   1 = mus, 2 = mua, 4 = g, 8 = n are sought; the others keep the scene's values
 *   bounds      {mus_lo, mus_hi, mua_lo, mua_hi, g_lo, g_hi, n_lo, n_hi} or NULL for the reference's
 *   targets     per detector, -1 = no target (detector%targetValue)
 *   table       max_steps x 5 (row-major): mus, mua, g, n, error of every step -- gradDescentData(:, 1:5)
 *   best_step   0-based row with the largest (least negative) error
 * Step k traces packet ids [id_offset, id_offset + nphotons) with seed + k.  One deliberate difference: the reference draws the
 * trial into its table but runs every step with the ORIGINAL properties (`mono(mus, mua, hgg, n)`, :1694 -- the trial values
 * never reach the scene); here the trial is what runs.  The scene's properties are restored on return.  Blocking. */
int smcrt_inverse_mcrt(smcrt_ctx* ctx, int top_index, int find_mask, const double* bounds, int max_steps, int64_t nphotons, uint64_t seed,
                       int tally_mode, const double* targets, double* table, int* best_step);
/* red.global.add.f32 throughput of device 0 on the context's own path-length grid (the secondary bound of -Dpathlength mode,
 * update_grids src/inttau2.f90:417-441).  pattern 0: uniform-random voxels; 1: every thread walks the same z-column of `span`
 * voxels (beam axis of a pencil source); 2: runs of `span` x-consecutive voxels from random starts (DDA-like).  Zeroes jmean. */
int smcrt_bench_red(smcrt_ctx* ctx, int pattern, int span, int64_t n_ops, double* ops_per_s);
/* The engine's Philox4x32-10 block for (seed, packet id, event index): 4 words. */
int smcrt_probe_philox(uint64_t seed, uint64_t packet_id, uint32_t event, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SMCRT_H */
