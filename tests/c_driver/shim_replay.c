/* shim_replay.c -- the call sequence of fortran/gpu_bridge.f90 (gpu_run_mcrt), from C, with the arrays laid out exactly as the
 * Fortran shim passes them: xform(16,n), params(8,n), p(20,n) column-major (= n consecutive records), top_node / first_child
 * 0-based, module arrays accumulated with accumulate = 1.  The Fortran file cannot be compiled in this image; this driver is the
 * compiled check of its side of the ABI.  Built by tests/test_gpu_c_driver.py:
 *     gcc -O2 -I include tests/c_driver/shim_replay.c -L rsmcrt_b200/lib -lsmcrt_gpu -o shim_replay
 *
 *   shim_replay scene.bin out.bin nphotons seed tally_mode survival_bias
 *
 * scene.bin (written by the test from oracle/scenes.py, little endian):
 *   int32 n_nodes, n_top, n_det, src_kind, src_sub, nxg, nyg, nzg;  double xmax, ymax, zmax;
 *   int32 kind[n_nodes], first_child[n_nodes], n_child[n_nodes], top_node[n_top];
 *   double xform[n_nodes*16], params[n_nodes*8], mus[n_top], mua[n_top], hgg[n_top], nref[n_top], src[24];
 *   int32 dkind[n_det], dnbins[n_det];  double dpar[n_det*20]
 * out.bin: int64 n_voxels, n_bins;  float absorb[n_voxels];  float jmean[n_voxels] (tally_mode & 2);  double bins[n_bins];
 *          double nscatt, launched, lost
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smcrt.h"

static void check(int rc, const char* what) {
    if (rc != 0) {
        fprintf(stderr, "libsmcrt_gpu: %s: %s\n", what, smcrt_last_error());
        exit(1); /* the shim: error stop 1 */
    }
}
static void rd(void* p, size_t bytes, FILE* f) {
    if (bytes && fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: shim_replay scene.bin out.bin nphotons seed tally_mode survival_bias\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t h[8];
    double gm[3];
    rd(h, sizeof h, f);
    rd(gm, sizeof gm, f);
    const int n_nodes = h[0], n_top = h[1], n_det = h[2], src_kind = h[3], src_sub = h[4], nxg = h[5], nyg = h[6], nzg = h[7];
    int32_t* kind = malloc(4 * (size_t)n_nodes), *first_child = malloc(4 * (size_t)n_nodes), *n_child = malloc(4 * (size_t)n_nodes);
    int32_t* top_node = malloc(4 * (size_t)n_top);
    double* xform = malloc(8 * 16 * (size_t)n_nodes), *params = malloc(8 * 8 * (size_t)n_nodes);
    double* mus = malloc(8 * (size_t)n_top), *mua = malloc(8 * (size_t)n_top), *hgg = malloc(8 * (size_t)n_top), *nref = malloc(8 * (size_t)n_top);
    double src[24];
    rd(kind, 4 * (size_t)n_nodes, f); rd(first_child, 4 * (size_t)n_nodes, f); rd(n_child, 4 * (size_t)n_nodes, f); rd(top_node, 4 * (size_t)n_top, f);
    rd(xform, 8 * 16 * (size_t)n_nodes, f); rd(params, 8 * 8 * (size_t)n_nodes, f);
    rd(mus, 8 * (size_t)n_top, f); rd(mua, 8 * (size_t)n_top, f); rd(hgg, 8 * (size_t)n_top, f); rd(nref, 8 * (size_t)n_top, f);
    rd(src, sizeof src, f);
    int32_t* dkind = malloc(4 * (size_t)(n_det + 1)), *dnbins = malloc(4 * (size_t)(n_det + 1));
    double* dpar = calloc(20 * (size_t)(n_det + 1), 8);
    rd(dkind, 4 * (size_t)n_det, f); rd(dnbins, 4 * (size_t)n_det, f); rd(dpar, 8 * 20 * (size_t)n_det, f);
    fclose(f);
    const int64_t nphotons = atoll(argv[3]);
    const uint64_t seed = strtoull(argv[4], NULL, 10);
    const int mode = atoi(argv[5]), survival = atoi(argv[6]);

    /* ---- gpu_run_mcrt, line by line */
    smcrt_ctx* ctx = NULL;
    check(smcrt_create(&ctx, 1, NULL), "smcrt_create");  /* (the shim passes 0 = every visible GPU) */
    check(smcrt_set_grid(ctx, nxg, nyg, nzg, gm[0], gm[1], gm[2]), "smcrt_set_grid");
    check(smcrt_set_scene(ctx, n_nodes, kind, first_child, n_child, xform, params, n_top, top_node, mus, mua, hgg, nref), "smcrt_set_scene");
    check(smcrt_set_source(ctx, src_kind, src_sub, src), "smcrt_set_source");
    check(smcrt_set_detectors(ctx, n_det, dkind, dpar, dnbins), "smcrt_set_detectors");
    check(smcrt_run(ctx, nphotons, seed, 0, mode, survival, -1.0, -1.0), "smcrt_run");
    const int64_t n_bins = smcrt_det_bins_total(ctx) > 0 ? smcrt_det_bins_total(ctx) : 1;
    const size_t nv = (size_t)nxg * nyg * nzg;
    float* absorb = calloc(nv, 4), *emission = calloc(nv, 4), *jmean = (mode & SMCRT_TALLY_PATHLENGTH) ? calloc(nv, 4) : NULL;
    double* bins = calloc((size_t)n_bins, 8);
    smcrt_counters cnt;
    memset(&cnt, 0, sizeof cnt);  /* accumulate = 1 ADDS to the caller's counters too */
    /* twice half of the work would do as well: the module arrays ACCUMULATE (accumulate = 1) */
    check(smcrt_fetch(ctx, jmean, absorb, emission, bins, &cnt, 1), "smcrt_fetch");
    check(smcrt_reset_tallies(ctx), "smcrt_reset_tallies");
    smcrt_destroy(ctx);

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    const int64_t hdr[2] = {(int64_t)nv, n_bins};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(absorb, 4, nv, f);
    if (jmean) fwrite(jmean, 4, nv, f);
    fwrite(bins, 8, (size_t)n_bins, f);
    const double tail[3] = {cnt.nscatt, cnt.launched, cnt.lost};
    fwrite(tail, sizeof tail, 1, f);
    fclose(f);
    return 0;
}
