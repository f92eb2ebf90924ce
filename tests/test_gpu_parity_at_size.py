"""Parity at the BASELINE packet counts: per-voxel / per-bin z-scores against the oracle at matched
counts, the literature depth-profile fits on the GPU, N GPUs == 1 GPU, and the batched-source entry point against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import RES, ROOT
from rsmcrt_b200 import api as A

pytestmark = pytest.mark.gpu


def test_sphere_toml_per_voxel_zscores_at_1e6(engine, oracle, smcrt):
    """BASELINE configs[0]: res/sphere.toml as shipped, 1e6 packets, path-length fluence.  32 batches of 31 250 packets on each
    side with INDEPENDENT seeds give a per-voxel mean and variance; the Welch statistic of every voxel of the 200^3 grid must be
    distributed like a t-statistic (the north-star 3-sigma bar, per voxel, not per coarse block)."""
    from scipy import stats
    cfg = smcrt.Config.load(RES / "sphere.toml")
    engine.apply(cfg)
    osc = oracle.OracleScene.from_toml(RES / "sphere.toml")
    mode, B, per = A.TALLY_PATHLENGTH, 32, 31_250
    nv = engine.n_voxels
    sg, qg, so, qo = (np.zeros(nv) for _ in range(4))
    for b in range(B):
        engine.reset_tallies()
        engine.run(per, 1000 + b, tally_mode=mode)
        j = engine.fetch(jmean=True, absorb=False, detectors=False)["jmean"].reshape(-1, order="F").astype(np.float64)
        sg += j; qg += j * j
        j = osc.run(per, 5000 + b, tally_mode=mode)["jmean"].reshape(-1, order="F").astype(np.float64)
        so += j; qo += j * j
    mg, mo = sg / B, so / B
    vg, vo = (qg / B - mg * mg) * B / (B - 1), (qo / B - mo * mo) * B / (B - 1)
    live = (vg > 0) & (vo > 0)
    assert live.mean() > 0.99                      # the uniform source illuminates every column
    z = (mg - mo)[live] / np.sqrt((vg + vo)[live] / B)
    dof = 2 * (B - 1)
    # bulk: the z histogram is the t distribution's, bin by bin
    edges = np.array([-np.inf, -3, -2, -1, 0, 1, 2, 3, np.inf])
    got = np.histogram(z, edges)[0] / z.size
    want = np.diff(stats.t.cdf(edges, dof))
    assert np.abs(got - want).max() < 0.012, (got, want)
    assert abs(z.mean()) < 0.02 and 0.93 < z.std() < 1.12, (z.mean(), z.std())
    # tails: a voxel sees ~25 packets, a batch 0.8: batch sums are compound-Poisson, heavier-tailed than a normal's
    assert (np.abs(z) > 4).mean() < 10 * 2 * stats.t.sf(4, dof), (np.abs(z) > 4).mean()
    assert np.abs(z).max() < 9.0
    # and the totals: mean path per packet within 3 sigma of the batch scatter
    tg, to = sg.sum() / (B * per), so.sum() / (B * per)
    assert abs(tg - to) < 5e-4 * to


def test_sphere_toml_same_stream_per_voxel_at_1e6(engine, oracle, smcrt):
    """The same job, the same Philox streams, 1e6 packets: the two fluence grids differ only where FP32 rounding flipped a discrete
    decision of a packet (a few per thousand)."""
    cfg = smcrt.Config.load(RES / "sphere.toml")
    engine.apply(cfg)
    n, mode = 1_000_000, A.TALLY_PATHLENGTH | A.TALLY_EMISSION
    engine.run(n, cfg.iseed, tally_mode=mode)
    g = engine.fetch(jmean=True, absorb=False, emission=True)
    o = oracle.OracleScene.from_toml(RES / "sphere.toml").run(n, cfg.iseed, tally_mode=mode)
    jg, jo = g["jmean"].astype(np.float64), o["jmean"].astype(np.float64)
    assert abs(jg.sum() - jo.sum()) < 2e-4 * jo.sum()
    assert np.abs(jg - jo).sum() < 0.012 * jo.sum()             # per voxel, L1
    # unit deposits at the launch voxels: FP32 vs FP64 launch positions put a few packets per 10^4 in the neighbouring voxel
    assert abs(g["emission"].sum() - n) < 0.5 and np.abs(g["emission"] - o["emission"]).sum() < 2e-4 * n
    assert g["counters"]["lost"] <= o["counters"]["lost"] + 5


def test_validation1_detector_bins_at_1e7(engine, oracle, smcrt):
    """BASELINE configs[1] scene at 1e7 packets per side, independent seeds: every populated bin of the two circle detectors within
    4 sigma of the oracle's (Poisson counts), chi-square of the z-scores consistent with its degrees of freedom, and the
    literature values Rd = 0.09739, Tt = 0.66096 (tools/validateHGG.py:14,26) within 3 sigma."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    N = 10_000_000
    engine.run(N, 2024)
    out = engine.fetch(absorb=True)
    g = out["det_bins"]
    o = oracle.OracleScene.from_toml(RES / "validation1.toml").run(N, 777, grids=True)
    ob = o["det_bins"]
    ok = (g + ob) >= 50          # (the beam is 1e-2 wide and a bin 0.2: only the innermost bins of either detector are populated)
    z = (g - ob)[ok] / np.sqrt((g + ob)[ok])
    assert ok.sum() >= 2
    assert np.abs(z).max() < 4.0, np.abs(z).max()
    assert g[~ok].sum() <= 60 * (~ok).sum() and abs(g[~ok].sum() - ob[~ok].sum()) <= 4 * np.sqrt(g[~ok].sum() + ob[~ok].sum() + 1)
    for lo, hi, lit in ((0, 101, 0.09739), (101, 202, 0.66096)):
        p = g[lo:hi].sum() / N
        assert abs(p - lit) < 3 * np.sqrt(lit * (1 - lit) / N) + 1.5e-4, (p, lit)
    # absorbed depth profile (500 z-slabs, ~330 inside the slab), Poisson z-scores against the oracle's
    zg, zo = out["absorb"].astype(np.float64).sum(axis=(0, 1)), o["absorb"].astype(np.float64).sum(axis=(0, 1))
    live = (zg + zo) >= 50
    zz = (zg - zo)[live] / np.sqrt((zg + zo)[live])
    assert np.abs(zz).max() < 4.5 and abs((zz * zz).sum() - live.sum()) < 4 * np.sqrt(2 * live.sum())


@pytest.mark.parametrize("deck,c1,k1,c2,k2,delta,N", [("validation2.toml", 5.76, 1.00, 1.31, 10.2, 0.047, 200_000),
                                                     ("validation3.toml", 6.27, 1.00, 1.18, 14.4, 0.261, 400_000)])
def test_index_mismatch_depth_profile_fit_on_the_gpu(engine, oracle, smcrt, deck, c1, k1, c2, k2, delta, N):
    """res/validation2.toml / validation3.toml (n = 1.38 slab, uniform 10 x 10 beam): the absorbed energy per unit depth follows the
    published two-exponential fit c1 exp(-k1 d/delta) - c2 exp(-k2 d/delta), d = depth below the top face at z = 1.95
    (tools/validateRIMismatch.py:28-46: both parameter sets).  Checked as a normalisation-free shape, on the GPU."""
    cfg = smcrt.Config.load(RES / deck)
    engine.apply(cfg)
    engine.run(N, cfg.iseed)
    out = engine.fetch(absorb=True)
    assert out["counters"]["lost"] <= 1e-5 * N
    prof = out["absorb"].astype(np.float64).sum(axis=(0, 1))
    (_, _, nz), (_, _, zmax) = cfg.grid
    z = (np.arange(nz) + 0.5) * 2 * zmax / nz - zmax
    depth = 1.95 - z
    sel = (depth > 0.08 * delta) & (depth < 2.0 * delta)
    fit = c1 * np.exp(-k1 * depth / delta) - c2 * np.exp(-k2 * depth / delta)
    m = (sel.sum() // 10) * 10
    p, f = prof[sel][:m].reshape(-1, 10).sum(1), fit[sel][:m].reshape(-1, 10).sum(1)
    p, f = p / p.sum(), f / f.sum()
    assert np.abs(p - f).max() < 0.12 * f.max(), np.abs(p - f).max() / f.max()
    assert prof[np.abs(z) > 1.951].sum() == 0     # nothing is absorbed outside the slab
    # and against the oracle on coarse depth bins (independent seeds)
    No = 20_000 if deck == "validation2.toml" else 60_000
    o = oracle.OracleScene.from_toml(RES / deck).run(No, 99, grids=True)["absorb"].astype(np.float64).sum(axis=(0, 1))
    k = nz // 50
    a, b = prof[: k * 50].reshape(50, k).sum(1), o[: k * 50].reshape(50, k).sum(1)
    live = (a + b) > 100
    zz = (a / N - b / No)[live] / np.sqrt((a / N**2 + b / No**2)[live])
    assert np.abs(zz).max() < 4.5, np.abs(zz).max()


def test_run_sources_matches_the_oracle_per_source(engine, oracle, smcrt):
    """smcrt_run_sources (the body of the escape-function drivers, kernelsMod.f90:533-642, for many cells in one launch) against
    the ORACLE: one oracle run per source position, isotropic point source there, the packet ids the batched launch gave that
    source.  Same streams -> per-(source, detector) totals equal up to the few histories FP32 rounding splits."""
    cfg = smcrt.Config.load(RES / "test_dects.toml")          # tau = 10 sphere, circle + annulus + camera
    engine.apply(cfg)
    pos = np.array([[0.0, 0.0, 0.0], [0.3, -0.2, 0.4], [5.0, 0.0, 0.0], [-0.5, 0.5, -0.1], [0.0, 0.0, 0.95]])
    n_per, seed = 20000, 77
    tot, layer = engine.run_sources(pos, n_per, seed)
    kind, dp, nb, _ = cfg.detectors
    offs = np.concatenate([[0], np.cumsum([(b + 1) if k != 4 else (b + 1) ** 2 for k, b in zip(kind, nb)])])
    osc = oracle.OracleScene.from_toml(RES / "test_dects.toml")
    k = 0
    for i in range(len(pos)):
        lay = osc.locate_layer(pos[i])
        assert layer[i] == lay
        if lay == 0:
            assert (tot[i] == 0).all()          # outside every SDF: escape = 0 without running (kernelsMod.f90:566-575)
            continue
        sp = np.zeros(24); sp[0:3] = pos[i]
        osc.set_source(A.SRC_POINT, 0, sp)
        o = osc.run(n_per, seed, id_offset=k * n_per, grids=False)["det_bins"]
        ref = np.array([o[offs[d]:offs[d + 1]].sum() for d in range(len(kind))])
        # (a camera counts SEGMENTS, detector_base.f90:222-229: their number depends on eps through the boundary nudges)
        tol = np.where(np.asarray(kind) == A.DET_CAMERA, 0.05, 0.005) * ref + 3
        assert np.all(np.abs(tot[i] - ref) <= tol), (i, tot[i], ref)
        k += 1
    assert tot.sum() > 0


def _n_gpus():
    try:
        return int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
    except OSError:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_n_gpu_equals_one_gpu(tmp_path, smcrt):
    """SURVEY S7's exit check.  One job on 1 GPU, on 2 GPUs of one context (smcrt_create(2): id ranges + in-process NCCL reduce at
    fetch) and on 2 ranks (one process per GPU, smcrt_comm_init / smcrt_comm_reduce): the integer-valued tallies -- detector bins
    (Q40.24), unit absorb deposits, counters -- are identical; the float path-length grid agrees to accumulation order."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    n, seed, mode = 4_000_000, 5, A.TALLY_ABSORB | A.TALLY_PATHLENGTH
    res = {}
    for g in (1, 2):
        e = smcrt.Engine(g)
        e.apply(cfg)
        e.run(n, seed, tally_mode=mode)
        res[g] = e.fetch(jmean=True, absorb=True)
        e.close()
    a, b = res[1], res[2]
    assert np.array_equal(a["det_bins"], b["det_bins"]) and np.array_equal(a["absorb"], b["absorb"])
    for k in ("launched", "nscatt", "lost", "det_hits"):
        assert a["counters"][k] == b["counters"][k], k
    assert np.allclose(a["jmean"], b["jmean"], rtol=2e-4, atol=1e-7)
    # one rank per GPU
    worker = ROOT / "tests" / "two_rank_worker.py"
    procs = [subprocess.Popen([sys.executable, str(worker), str(r), "2", str(tmp_path), str(n), str(seed), str(mode)],
                              env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(r))) for r in range(2)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    r0 = np.load(tmp_path / "rank0.npz")
    assert np.array_equal(r0["det_bins"], a["det_bins"]) and np.array_equal(r0["absorb"], a["absorb"])
    assert r0["nscatt"] == a["counters"]["nscatt"] and r0["launched"] == n
    assert np.allclose(r0["jmean"], a["jmean"], rtol=2e-4, atol=1e-7)


def test_set_optprops_matches_the_oracle(engine, oracle, smcrt):
    """inverse_MCRT changes one layer's optical properties between runs (`array(i)%updateOptProp`, src/kernelsMod.f90:1693-1698).
    smcrt_set_optprops followed by a run must give what the oracle gives after orc_set_optprops, on the same streams."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    osc = oracle.OracleScene.from_toml(RES / "validation1.toml")
    n = 200_000
    for k, (mus, mua, g, nr) in enumerate([(60.0, 25.0, 0.5, 1.0), (120.0, 4.0, 0.9, 1.0), (90.0, 10.0, 0.75, 1.4)]):
        engine.set_optprops(1, mus, mua, g, nr)
        osc.set_optprops(1, mus, mua, g, nr)
        engine.reset_tallies()
        engine.run(n, 40 + k)
        out = engine.fetch(absorb=True)
        o = osc.run(n, 40 + k, grids=True)
        assert abs(out["absorb"].sum() - o["absorb"].sum()) <= 5e-4 * n + 4
        if nr == 1.0:
            assert abs(out["det_bins"].sum() - o["det_bins"].sum()) <= 5e-4 * n + 4
            assert np.abs(out["det_bins"] - o["det_bins"]).max() <= 6 + 0.02 * o["det_bins"].max()
        else:
            # Detectors that COINCIDE with a refracting surface (the slab's faces at n = 1.4).  The reference tests the crossing
            # piece -- whose length was measured along the pre-refraction direction (quirk Q4) -- along the refracted direction:
            # t = gap / cos(theta_t) overshoots the piece for grazing exits and 1.7 % of the escaping packets are never recorded
            # (the oracle reproduces that).  The engine's end-point test is watertight: every packet that
            # leaves through a face is counted, so its total is the number of escaped packets and never below the oracle's.
            escaped = n - out["absorb"].sum()
            assert abs(out["det_bins"].sum() - escaped) <= 2
            assert 0 <= out["det_bins"].sum() - o["det_bins"].sum() <= 0.03 * escaped
        assert abs(out["counters"]["nscatt"] - o["counters"]["nscatt"]) <= 0.004 * o["counters"]["nscatt"] + 60
        zg, zo = out["absorb"].astype(np.float64).sum(axis=(0, 1)), o["absorb"].astype(np.float64).sum(axis=(0, 1))
        assert np.abs(zg - zo).max() <= 8 + np.sqrt(zo.max())


def test_inverse_mcrt_search_loop(engine, oracle, smcrt):
    """smcrt_inverse_mcrt = the loop of inverse_MCRT (src/kernelsMod.f90:1462-1751): uniform trial points inside the bounds, one run
    per trial, inverse_evaluate as the score.  Targets are the slab's own Rd / Tt at mua = 10: every row of the table must be
    reproducible by hand (set_optprops + run + the error formula, checked against the ORACLE's detectors for the best row), the
    best row must be the trial closest to the truth, and the scene's own properties must be back afterwards."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    targets = np.array([0.09739, 0.66096])                      # tools/validateHGG.py:14,26
    n, seed, steps = 400_000, 9, 10
    table, best = engine.inverse_mcrt(1, 2, targets, steps, n, seed, bounds=[0, 100, 2.0, 30.0, -1, 1, 1, 20])
    assert (table[:, 0] == 90.0).all() and (table[:, 2] == 0.75).all() and (table[:, 3] == 1.0).all()   # only mua is sought
    assert ((table[:, 1] >= 2.0) & (table[:, 1] <= 30.0)).all() and len(set(table[:, 1])) == steps
    assert (table[:, 4] <= 0).all() and best == int(np.argmax(table[:, 4]))
    assert abs(table[best, 1] - 10.0) == np.abs(table[:, 1] - 10.0).min()      # |Rd - Rd*| + |Tt - Tt*| is monotone in |mua - 10| here
    # the best row by hand, detectors from the ORACLE on the same streams
    osc = oracle.OracleScene.from_toml(RES / "validation1.toml")
    osc.set_optprops(1, *table[best, :4])
    ob = osc.run(n, seed + best, grids=False)["det_bins"]
    err = -0.5 * (abs(ob[:101].sum() / n - targets[0]) + abs(ob[101:].sum() / n - targets[1]))
    assert abs(err - table[best, 4]) < 6e-4
    # the scene is as it was
    engine.run(n, seed)
    bins = engine.fetch(absorb=False)["det_bins"]
    assert abs(bins[:101].sum() / n - targets[0]) < 2e-3 and abs(bins[101:].sum() / n - targets[1]) < 3e-3
