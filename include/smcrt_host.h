/*
 * smcrt_host.h — C ABI of the host-side mirror of the reference's config / scene-setup / output layer.
 *
 * The reference's host is Fortran 2018 (src/parse/*.f90, src/setup.f90, src/setupGeometry.f90,
 * src/writer.f90, src/kernelsMod.f90 setup/finalise/default_MCRT).  No Fortran toolchain exists in this
 * image, so the same interface is provided in C++ above the engine's C ABI (include/smcrt.h): same TOML
 * keys and defaults, same geom_name dispatch and scene contents (App. E), same output
 * bytes (App. D / §8f N1).  It lives in the same shared object as the engine (libsmcrt_gpu.so).
 */
#ifndef SMCRT_HOST_H
#define SMCRT_HOST_H

#include <stdint.h>
#include "smcrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smcrt_config smcrt_config; /* opaque: `state` + `dict` + dects + flattened array(:) */

/* parse_params (src/parse/parse.f90:20-72).  res_dir is where geometry side files live (the reference
   hard-codes "res/": setupGeometry.f90:586-627); may be NULL (= directory of the toml file). */
int smcrt_config_load(const char* toml_path, const char* res_dir, smcrt_config** out);
int smcrt_config_loads(const char* toml_text, const char* res_dir, smcrt_config** out);
void smcrt_config_free(smcrt_config* cfg);

/* parsed values (state%..., src/sim_state.f90:10-58) */
int smcrt_config_grid(const smcrt_config* cfg, int32_t n[3], double half_extent[3]);
int64_t smcrt_config_nphotons(const smcrt_config* cfg);
int64_t smcrt_config_iseed(const smcrt_config* cfg);
const char* smcrt_config_geom_name(const smcrt_config* cfg);
const char* smcrt_config_source_name(const smcrt_config* cfg);
int smcrt_config_render_source(const smcrt_config* cfg);
/* source block as smcrt_set_source wants it */
int smcrt_config_source(const smcrt_config* cfg, int32_t* kind, int32_t* subtype, double p[SMCRT_SOURCE_PARAMS]);
/* detectors in the reference's dects(:) order: circles, annuli, fibres, cameras (parse_detectors.f90:119-137) */
int smcrt_config_n_detectors(const smcrt_config* cfg);
int smcrt_config_detectors(const smcrt_config* cfg, int32_t* kind, double* p /* n x SMCRT_DET_PARAMS */, int32_t* nbins);
const char* smcrt_config_detector_id(const smcrt_config* cfg, int i /*0-based, dects order*/);

/* setup_simulation's geom_name dispatch (src/setup.f90:33-60) -> flattened array(:) */
int smcrt_config_scene_sizes(const smcrt_config* cfg, int32_t* n_nodes, int32_t* n_top);
int smcrt_config_scene(const smcrt_config* cfg, int32_t* kind, int32_t* first_child, int32_t* n_child, double* xform,
                       double* params, int32_t* top_node, double* mus, double* mua, double* hgg, double* n_ref);

/* push grid + scene + source + detectors into an engine context (what setup() leaves in module state) */
int smcrt_config_apply(const smcrt_config* cfg, smcrt_ctx* ctx);

/* ---- driver arithmetic ------------------------------------------------------------------------ */
/* inverse_evaluate (src/kernelsMod.f90:1753-1787): error = -mean_i |total_i / nphotons - target_i| over the detectors whose
   target is not -1 ("no target").  totals: per-detector sums of bins (total_dect).  Returns -1 when no detector has a target
   (the reference divides by zero there). */
int smcrt_inverse_evaluate(int n_det, const double* totals, const double* targets, int64_t nphotons, double* error);
/* escape-function grid cell centre (cart_calc_escape_sym, src/kernelsMod.f90:576-585): ((i - 0.5) / n) * 2 max - max for the
   1-based cell (m, n, o), then the two row-vector rotations and the shift.  rot_z / rot_off: 4x4 as stored by the reference
   (column-major, vec .dot. mat), may be NULL (identity). */
int smcrt_escape_cell_centre(int m, int n, int o, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax,
                             const double* rot_z, const double* rot_off, const double* grid_pos, double* out_xyz);

/* ---- output layer (src/writer.f90) ---------------------------------------------------------- */
/* normalise_fluence (writer.f90:25-52): array *= nx*ny*nz / nphotons, evaluated like the reference */
int smcrt_normalise_fluence(float* array, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax, int64_t nphotons);
/* write_3d_r4_nrrd (writer.f90:382-424, header :304-337).  meta (may be NULL) is dumped between header and
   data like toml_dump(dict). */
int smcrt_write_nrrd_f32(const char* path, const float* data, int nxg, int nyg, int nzg, const char* meta);
/* write_detected_photons (writer.f90:55-134): out_dir/detector_<i>.dat for every non-camera detector */
int smcrt_write_detectors(const smcrt_config* cfg, const double* det_bins, const char* out_dir);
/* history_stack_t%write / %finish (src/historyStack.f90:163-226): the vertex lists of the detected packets as ONE file, type by
   extension like init_historyStack (:47-55): .obj ("v x y z" lines, then one "l i j k ..." polyline per packet, 1-based vertex
   indices), .ply (ascii header with the final vertex / edge counts, vertices, then vertex-index pairs) or .json ({"<k>_0": [[x,y,z],
   ...], ...}), numbers as es15.8e2 like the reference's writers.  vertices: n x max_vertices x 4 floats, counts[k] <= max_vertices
   vertices of packet k are written. */
int smcrt_history_write(const char* path, int64_t n, int max_vertices, const float* vertices, const int32_t* counts);
/* [[detectors]] trackHistory of detector i (dects(:) order) and the [[detectors]] historyFileName (default "photPos.obj") */
int smcrt_config_detector_track(const smcrt_config* cfg, int i);
const char* smcrt_config_history_filename(const smcrt_config* cfg);
/* metadata text written into NRRD headers (the `dict` toml dump, kernelsMod.f90:2378-2382) */
const char* smcrt_config_metadata(const smcrt_config* cfg);

/* checkpoint (src/writer.f90:426-457): "tomlfile=<name>\nphotons_run=<n>\n" followed by the raw float32 jmean grid (x fastest).
   smcrt_checkpoint_read is the reader of default_MCRT's load_checkpoint branch (src/kernelsMod.f90:52-72): toml_out receives the
   input-deck name (NUL-terminated, truncated to toml_cap), jmean (may be NULL) the n_voxels floats. */
int smcrt_checkpoint_write(const char* path, const char* toml_filename, int64_t nphotons_run, const float* jmean, int64_t n_voxels);
int smcrt_checkpoint_read(const char* path, char* toml_out, int toml_cap, int64_t* nphotons_run, float* jmean, int64_t n_voxels);

/* default_MCRT (src/kernelsMod.f90:29-83): setup -> run_MCRT -> finalise.
   Checkpoints ([simulation] load_checkpoint / checkpoint_file / checkpoint_every_n): the run is cut at multiples of
   checkpoint_every_n and the file rewritten there, but not more often than every ~2 s (one packet takes ~0.3 ns here, not ~10 us);
   a resumed run continues the SAME job (same seed, packet ids from photons_run on -- the reference re-seeds with iseed*101).  Writes out_dir/{jmean,absorb,emission,
   detectors}/... with the reference's names.  tally_mode<0: default build semantics (absorb [+emission when
   render_source]); nphotons<=0: the toml's.  photons_per_s (may be NULL) receives what the reference prints
   (kernelsMod.f90:1897). */
int smcrt_default_mcrt(const char* toml_path, const char* res_dir, const char* out_dir, int n_gpus, int tally_mode,
                       int survival_bias, int64_t nphotons, double* photons_per_s, smcrt_counters* counters);

#ifdef __cplusplus
}
#endif
#endif
