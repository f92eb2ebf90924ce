"""Size-independent properties at the BASELINE problem sizes (no oracle needed: the oracle cannot run 1e8 packets in seconds)."""
import numpy as np
import pytest

from conftest import RES
from rsmcrt_b200 import api as A

pytestmark = pytest.mark.gpu


def test_validation1_1e8_conservation_and_literature(engine, smcrt):
    """BASELINE configs[1] at full size (1e8 packets on one B200): every packet is either absorbed (one unit deposit in the
    absorb grid) or leaves; detector tallies are a subset of the leavers; Rd / Tt hit the literature values at 3 sigma."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    N = 100_000_000
    engine.run(N, cfg.iseed)
    out = engine.fetch(absorb=True)
    c = out["counters"]
    assert c["launched"] == N and c["lost"] == 0
    absorbed = float(out["absorb"].astype(np.float64).sum())
    assert absorbed == np.round(absorbed)                      # unit deposits: exact integers in FP32 (< 2^24 per voxel)
    assert out["absorb"].max() < 2 ** 24
    bins = out["det_bins"]
    assert (bins == np.round(bins)).all()                      # weight-1 hits in Q40.24 fixed point are exact
    Rd, Tt = bins[:101].sum() / N, bins[101:].sum() / N
    assert abs(absorbed / N + Rd + Tt - 1.0) < 2e-4             # the rest leaves through the slab edges / beyond r = 20: negligible
    assert Rd == pytest.approx(0.09739, abs=3 * np.sqrt(0.09739 * 0.90261 / N) + 1.5e-4)
    assert Tt == pytest.approx(0.66096, abs=3 * np.sqrt(0.66096 * 0.33904 / N) + 1.5e-4)
    # nothing is absorbed outside the slab |z| <= 0.01 (grid z spans +-0.015 in 500 slabs)
    prof = out["absorb"].sum(axis=(0, 1))
    z = (np.arange(500) + 0.5) * 0.03 / 500 - 0.015
    assert prof[np.abs(z) > 0.01 + 0.03 / 500].sum() == 0
    assert c["nscatt"] / N == pytest.approx(2.175, abs=5e-3)   # BASELINE.md derived anchor


def test_reproducible_and_split_invariant(engine, smcrt):
    """Packet streams depend only on (seed, id): two runs, and one run split into id ranges, give bit-identical integer tallies."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    N = 5_000_000
    engine.run(N, 42)
    a = engine.fetch(absorb=True)
    engine.reset_tallies()
    engine.run(N, 42)
    b = engine.fetch(absorb=True)
    engine.reset_tallies()
    for off in range(0, N, N // 5):
        engine.run(N // 5, 42, id_offset=off)
    c = engine.fetch(absorb=True)
    for other in (b, c):
        assert (a["det_bins"] == other["det_bins"]).all()
        assert (a["absorb"] == other["absorb"]).all()
        assert a["counters"]["nscatt"] == other["counters"]["nscatt"]
    engine.reset_tallies()
    engine.run(N, 43)
    d = engine.fetch(absorb=True)
    assert (a["det_bins"] != d["det_bins"]).any()               # a different seed is a different sample


def test_edge_cases(engine, smcrt):
    cfg = smcrt.Config.load(RES / "scat_test.toml")
    engine.apply(cfg)
    engine.run(0, 1)                                            # empty job
    assert engine.fetch()["counters"]["launched"] == 0
    engine.run(1, 1)                                            # a single packet
    engine.run(33, 1, id_offset=2 ** 40 + 7)                    # ragged count, ids beyond 32 bits
    c = engine.fetch()["counters"]
    assert c["launched"] == 34 and c["lost"] == 0
    g = engine.trace_packets(1000, 9, id_offset=2 ** 33)
    assert (g["fate"] == A.FATE_ESCAPED).all() and g["nscatt"].mean() > 30
    with pytest.raises(smcrt.SmcrtError):
        engine.run(-5, 1)
    e2 = smcrt.Engine(1)
    with pytest.raises(smcrt.SmcrtError):                       # no grid / scene set
        e2.run(10, 1)
    e2.close()
    with pytest.raises(smcrt.SmcrtError):
        engine.set_source(99, 0, np.zeros(24))                  # "No such source!"


def test_sphere_scene_1e6_energy_bookkeeping(engine, smcrt):
    """BASELINE configs[0] (res/sphere.toml as shipped, 1e6 packets): non-absorbing scene -> nothing absorbed, every packet
    leaves, path-length fluence sums to (packets x mean chord >= 2), emission grid counts every packet once."""
    cfg = smcrt.Config.load(RES / "sphere.toml")
    engine.apply(cfg)
    N = 1_000_000
    engine.run(N, cfg.iseed, tally_mode=A.TALLY_ABSORB | A.TALLY_PATHLENGTH | A.TALLY_EMISSION)
    out = engine.fetch(jmean=True, absorb=True, emission=True)
    c = out["counters"]
    assert out["absorb"].sum() == 0 and c["nscatt"] == 0
    assert c["lost"] <= 20                                      # bounce cap (> 1000 internal reflections), ~1e-5 of packets
    assert out["emission"].astype(np.float64).sum() == N
    mean_path = out["jmean"].astype(np.float64).sum() / N
    assert 2.0 <= mean_path < 2.2


def _absorbed_fraction_oracle(oracle, deck, n, res_dir=None):
    o = oracle.OracleScene.from_toml(RES / deck, res_dir).run(n, 4711, grids=True)
    return float(o["absorb"].astype(np.float64).sum()) / n, o


def test_skin_1e9_properties(engine, oracle, smcrt):
    """BASELINE configs[2] at full size: the five-layer skin stack (refractive-index mismatch at every interface), 1e9 packets on
    one GPU.  No packet is lost to an engine guard, unit absorb deposits stay exact integers, nothing is absorbed outside the
    tissue, the absorbed fraction and the scatter count per packet agree with the oracle's (1e5 packets) within its 3 sigma."""
    cfg = smcrt.Config.load(RES / "skin_b200.toml")
    engine.apply(cfg)
    N = 1_000_000_000
    engine.run(N, cfg.iseed)
    out = engine.fetch(absorb=True)
    c = out["counters"]
    assert c["launched"] == N and c["lost"] <= 1e-7 * N
    a64 = out["absorb"].astype(np.float64)
    assert out["absorb"].max() < 2 ** 24 and (out["absorb"] == np.round(out["absorb"])).all()
    n_o = 100_000
    fo, o = _absorbed_fraction_oracle(oracle, "skin_b200.toml", n_o)
    fg = a64.sum() / N
    assert abs(fg - fo) < 3 * np.sqrt(fo * (1 - fo) / n_o) + 1e-4, (fg, fo)
    so = o["counters"]["nscatt"] / n_o
    assert abs(c["nscatt"] / N - so) < 0.02 * so
    # depth profile of the absorbed energy against the oracle's, 20 slabs
    zg, zo = a64.sum(axis=(0, 1)).reshape(20, 10).sum(1) / N, o["absorb"].astype(np.float64).sum(axis=(0, 1)).reshape(20, 10).sum(1) / n_o
    assert np.all(np.abs(zg - zo) < 4 * np.sqrt(np.maximum(zo, 1e-6) / n_o) + 1e-5)


def test_lens_1e9_properties(engine, oracle, smcrt):
    """BASELINE configs[3] scene (refractive lens: intersection of two spheres, n = 1.5) at 1e9 packets, path-length fluence:
    nothing is absorbed or scattered, every packet leaves, the mean path per packet is the oracle's, no packet is lost."""
    cfg = smcrt.Config.load(RES / "lens.toml")
    engine.apply(cfg)
    N = 1_000_000_000
    engine.run(N, cfg.iseed, tally_mode=A.TALLY_ABSORB | A.TALLY_PATHLENGTH)
    out = engine.fetch(jmean=True, absorb=True)
    c = out["counters"]
    assert c["launched"] == N and c["lost"] <= 1e-6 * N and c["nscatt"] == 0 and out["absorb"].sum() == 0
    path_g = out["jmean"].astype(np.float64).sum() / N
    o = oracle.OracleScene.from_toml(RES / "lens.toml").run(200_000, 11, tally_mode=A.TALLY_PATHLENGTH)
    path_o = o["jmean"].astype(np.float64).sum() / 200_000
    assert abs(path_g - path_o) < 2e-3 * path_o, (path_g, path_o)
    assert abs(c["bounces"] / N - o["counters"]["bounces"] / 200_000) < 0.02 * o["counters"]["bounces"] / 200_000 + 1e-3


def test_vessels_1e9_properties(engine, oracle, smcrt, tmp_path):
    """BASELINE configs[4] scene (vessel tree: 240 capsules in a dermis box; the reference ships no data, tools/make_vessels.py
    writes a seeded synthetic tree in its file formats) at 1e9 packets on one GPU (the 1e10 job is ten of these: the
    committed multi-GPU bench lines).  Conservation and oracle agreement of the absorbed fraction, in and out of the vessels."""
    import sys
    sys.path.insert(0, str(RES.parent / "tools"))
    import make_vessels
    make_vessels.make(tmp_path, 240, 7)
    cfg = smcrt.Config.load(RES / "vessels.toml", res_dir=tmp_path)
    engine.apply(cfg)
    N = 1_000_000_000
    engine.run(N, cfg.iseed)
    out = engine.fetch(absorb=True)
    c = out["counters"]
    assert c["launched"] == N and c["lost"] == 0
    a64 = out["absorb"].astype(np.float64)
    assert (out["absorb"] == np.round(out["absorb"])).all() and out["absorb"].max() < 2 ** 24
    n_o = 100_000
    fo, o = _absorbed_fraction_oracle(oracle, "vessels.toml", n_o, tmp_path)
    fg = a64.sum() / N
    assert abs(fg - fo) < 3 * np.sqrt(fo * (1 - fo) / n_o) + 1e-4, (fg, fo)
    so = o["counters"]["nscatt"] / n_o
    assert abs(c["nscatt"] / N - so) < 0.02 * so
