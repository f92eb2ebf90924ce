"""GPU-vs-oracle parity on the remaining BASELINE configs (jacques / skin / lens / vessels) and on compound SDF scenes
(omg: smooth-union model; egg: revolution modifier).  For the builder-defined geometries the oracle is the
only reference there is."""
import sys

import numpy as np
import pytest

from conftest import RES, ROOT
from rsmcrt_b200 import api as A

pytestmark = pytest.mark.gpu


def both(smcrt, oracle, engine, name, n, seed=11, mode=A.TALLY_ABSORB, **kw):
    cfg = smcrt.Config.load(RES / name, **kw)
    engine.apply(cfg)
    osc = oracle.OracleScene.from_config(cfg)
    g = engine.trace_packets(n, seed, tally_mode=mode)
    out = engine.fetch(jmean=bool(mode & A.TALLY_PATHLENGTH), absorb=True)
    o = osc.run(n, seed, per_packet=True, tally_mode=mode)
    return cfg, g, out, o


def ensemble_close(g, o, n):
    """fate fractions and mean scatter counts agree within 4 sigma (independent-sample bound; same streams do better)."""
    for f in (A.FATE_ABSORBED, A.FATE_ESCAPED):
        pg, po = (g["fate"] == f).mean(), (o["fate"] == f).mean()
        assert abs(pg - po) < 4 * np.sqrt(2 * 0.25 / n) + 1e-9
    ng, no = g["nscatt"].astype(float), o["nscatt"].astype(float)
    assert abs(ng.mean() - no.mean()) < 4 * np.sqrt((ng.var() + no.var()) / n) + 1e-9


def test_jacques_block(engine, oracle, smcrt):
    cfg, g, out, o = both(smcrt, oracle, engine, "jacques.toml", 20000)
    assert out["counters"]["lost"] == 0 and (o["fate"] != A.FATE_LOST).all()
    ensemble_close(g, o, 20000)
    # absorbed-energy depth profile (uniform illumination from the top): coarse z bins, 4 sigma
    zg = out["absorb"].sum(axis=(0, 1)).reshape(20, 20).sum(1)
    zo = o["absorb"].sum(axis=(0, 1)).reshape(20, 20).sum(1).astype(float)
    assert np.all(np.abs(zg - zo) < 4 * np.sqrt(zg + zo + 1) + 3)


def test_skin_layers(engine, oracle, smcrt):
    cfg, g, out, o = both(smcrt, oracle, engine, "skin_b200.toml", 20000)
    assert out["counters"]["lost"] <= 2 and (o["fate"] == A.FATE_LOST).sum() <= 2
    ensemble_close(g, o, 20000)
    zg = out["absorb"].sum(axis=(0, 1)).reshape(20, 10).sum(1)
    zo = o["absorb"].sum(axis=(0, 1)).reshape(20, 10).sum(1).astype(float)
    assert np.all(np.abs(zg - zo) < 4 * np.sqrt(zg + zo + 1) + 3)
    # internal reflections happen at the index steps between layers
    assert out["counters"]["bounces"] > 0


def test_lens_refraction(engine, oracle, smcrt):
    """bi-convex lens = intersection(sphere, sphere): compound SDF (program interpreter + FP64 normal through the model)."""
    mode = A.TALLY_PATHLENGTH
    cfg, g, out, o = both(smcrt, oracle, engine, "lens.toml", 20000, mode=mode)
    assert out["counters"]["lost"] <= 2
    assert (g["fate"] == A.FATE_ESCAPED).mean() > 0.999
    same = (g["events"] == o["events"])
    assert same.mean() > 0.97  # same number of Fresnel events for (nearly) every packet
    # the lens focuses the beam: compare the path-length fluence in coarse blocks
    cg = out["jmean"].astype(float).reshape(10, 20, 10, 20, 10, 20).sum(axis=(1, 3, 5))
    co = o["jmean"].astype(float).reshape(10, 20, 10, 20, 10, 20).sum(axis=(1, 3, 5))
    assert abs(cg.sum() - co.sum()) < 0.005 * co.sum()
    assert np.abs(cg - co).sum() < 0.03 * co.sum()
    # exit points of identical histories coincide
    d = np.abs(g["pos"] - o["pos"])[same].max(axis=1)
    assert np.median(d) < 1e-4


def test_vessels_capsules(engine, oracle, smcrt, tmp_path):
    sys.path.insert(0, str(ROOT / "tools"))
    import make_vessels
    make_vessels.make(tmp_path, 240, 7)
    cfg, g, out, o = both(smcrt, oracle, engine, "vessels.toml", 10000, res_dir=tmp_path)
    assert cfg.scene.n_top == 241
    assert out["counters"]["lost"] == 0
    ensemble_close(g, o, 10000)
    assert abs(out["absorb"].sum() - o["absorb"].sum()) < 4 * np.sqrt(2 * 0.25 * 10000)


def test_omg_smooth_union(engine, oracle, smcrt):
    cfg, g, out, o = both(smcrt, oracle, engine, "omg.toml", 10000)
    assert out["counters"]["lost"] <= 2
    ensemble_close(g, o, 10000)


def test_egg_revolution(engine, oracle, smcrt):
    text = (RES / "egg_test.toml").read_text().replace("nxg = 500", "nxg = 100").replace("nyg = 500", "nyg = 100").replace("nzg = 500", "nzg = 100")
    text = text.replace("[geometry]", "[geometry]\nmus = [10.0, 1.0, 5.0]\nmua = [0.5, 0.1, 1.0]\nhgg = [0.8, 0.9, 0.7]\nn = [1.5, 1.35, 1.4]")
    cfg = smcrt.Config.loads(text)
    engine.apply(cfg)
    osc = oracle.OracleScene.from_config(cfg)
    n = 10000
    g = engine.trace_packets(n, 3)
    o = osc.run(n, 3, per_packet=True, grids=False)
    assert (g["fate"] == A.FATE_LOST).sum() <= 2
    ensemble_close(g, o, n)


def test_checkpoint_and_resume_is_the_same_job(smcrt, tmp_path, monkeypatch):
    """[simulation] checkpoint_every_n / load_checkpoint (kernelsMod.f90:52-72,1863; writer.f90:426-457): the run is cut at multiples
    of checkpoint_every_n and the checkpoint rewritten; resuming from it traces the packets that were still to come -- same seed,
    ids from photons_run on -- so the path-length grid equals that of the uninterrupted run."""
    from rsmcrt_b200 import api as A
    monkeypatch.setenv("SMCRT_CKPT_MIN_PIECE", "1")
    monkeypatch.setenv("SMCRT_CKPT_MIN_SECONDS", "0")
    ck = tmp_path / "run.ckpt"
    deck = (RES / "scat_test.toml").read_text()   # ships load_checkpoint=false, checkpoint_file="check.ckpt", checkpoint_every_n=10000
    assert 'checkpoint_file="check.ckpt"' in deck and "checkpoint_every_n=10000" in deck and "load_checkpoint=false" in deck
    deck = deck.replace('checkpoint_file="check.ckpt"', f'checkpoint_file="{ck}"').replace("checkpoint_every_n=10000", "checkpoint_every_n=50000")
    base = tmp_path / "deck.toml"
    base.write_text(deck)
    n = 200_000
    mode = A.TALLY_ABSORB | A.TALLY_PATHLENGTH
    smcrt.default_MCRT(base, out_dir=tmp_path / "full", tally_mode=mode, nphotons=n)
    name, run, jm = smcrt.checkpoint_read(ck, 0)
    assert name == str(base) and run == 150_000                    # the last cut before the end
    cfg = smcrt.Config.load(base)
    nv = int(np.prod(cfg.grid[0]))
    _, _, jm = smcrt.checkpoint_read(ck, nv)
    assert jm.sum() > 0
    resume = tmp_path / "resume.toml"
    resume.write_text(deck.replace("load_checkpoint=false", "load_checkpoint=true"))
    smcrt.default_MCRT(resume, out_dir=tmp_path / "resumed", tally_mode=mode, nphotons=n)

    def grid(d):
        raw = (tmp_path / d / "jmean" / "fluence.nrrd").read_bytes()
        return np.frombuffer(raw[-4 * nv:], np.float32).astype(np.float64)
    a, b = grid("full"), grid("resumed")
    assert a.sum() > 0
    assert abs(a.sum() - b.sum()) <= 1e-5 * a.sum()                 # float atomics: same deposits, different order
    assert np.abs(a - b).max() <= 1e-4 * a.max()


def test_nonrigid_transform_takes_plain_sphere_tracing(engine, oracle, smcrt):
    """A scaled primitive: the reference accepts any 4x4 transform and steps by |d| (inttau2.f90:155-192).  The engine's
    closed-form ray bounds assume a unit local direction, which only a rigid transform keeps: a non-rigid primitive must report
    bound = |d|, exact = false, and a scene with one must trace like the oracle's."""
    from common import random_dirs
    scale = np.diag([0.5, 0.5, 0.5, 1.0])                      # local = 0.5 * world: a sphere of radius 0.25 becomes one of 0.5
    shear = np.eye(4); shear[0, 1] = 0.3                       # and a sheared box
    scene = A.Scene.from_primitives([(A.SPHERE, scale.reshape(-1, order="F"), [0.25]), (A.BOX, shear.reshape(-1, order="F"), [0.9, 0.9, 0.9]),
                                     (A.BOX, None, [1.0, 1.0, 1.0])],
                                    [(5.0, 0.5, 0.6, 1.33), (1.0, 0.1, 0.0, 1.2), (0.0, 0.0, 0.0, 1.0)])
    engine.set_grid(50, 50, 50, 1.0, 1.0, 1.0)
    engine.set_scene(scene)
    sp = np.zeros(24); sp[0:3] = [0.05, -0.02, 0.03]
    engine.set_source(A.SRC_POINT, 0, sp)
    engine.set_detectors([], np.zeros((0, 20)), [])
    rng = np.random.default_rng(3)
    pos, dirs = rng.uniform(-0.9, 0.9, (20000, 3)), random_dirs(rng, 20000)
    for top in (1, 2):
        d, b, ex = engine.probe_ray(top, pos, dirs)
        assert (ex == 0).all() and np.allclose(b, np.abs(d), rtol=1e-6, atol=1e-7)
    d3, b3, ex3 = engine.probe_ray(3, pos, dirs)                # the untransformed box keeps its exact ray bound
    assert (ex3 == 1).all() and (b3 >= np.abs(d3) - 1e-6).all()
    osc = oracle.OracleScene(scene, ((50, 50, 50), (1.0, 1.0, 1.0)), (A.SRC_POINT, 0, sp))
    n = 20000
    g = engine.trace_packets(n, 9)
    o = osc.run(n, 9, per_packet=True, grids=False)
    same = (g["fate"] == o["fate"]) & (g["nscatt"] == o["nscatt"])
    assert same.mean() > 0.97, same.mean()
    assert (g["fate"] == A.FATE_LOST).mean() < 2e-3
    ensemble_close(g, o, n)


def test_queue_watchdog_reports_an_error(tmp_path):
    """trace_queued's watchdog (kernels.cuh): a warp that finds every queue empty for the whole watchdog period leaves and the run is
    reported as failed by smcrt_wait instead of hanging the device.  Tripped here on purpose with a period of one microsecond
    (warps idle for longer than that whenever the last histories of a run finish), in a subprocess: the period is read once."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import rsmcrt_b200 as R\n"
            "e = R.Engine(1); e.apply(R.Config.load(%r))\n"
            "try:\n"
            "    e.run(2000000, 1)\n"
            "    print('NO ERROR')\n"
            "except R.SmcrtError as ex:\n"
            "    print('ERROR:', ex)\n"
            "e.reset_tallies(); e.close()\n") % (str(RES.parent), str(RES / "jacques.toml"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SMCRT_WATCHDOG_MS="0.001", SMCRT_VARIANT_FORCE="4"),
                       capture_output=True, text=True, timeout=120)
    assert "ERROR:" in r.stdout and "watchdog" in r.stdout, (r.stdout, r.stderr[-500:])


def test_track_history_replay(engine, oracle, smcrt, tmp_path):
    """trackHistory (src/historyStack.f90:89-108,184-226): the run notes which packets hit a history-tracking
    detector; tracing exactly those packets again (same seed, same ids -> same streams) yields their vertex lists.
    Checked: the hit list is the set of packets the ORACLE detects on that detector; a replayed list starts at the source, has one
    vertex per interaction (the packet's scatter count, from smcrt_trace_packets), ends on the detector plane inside its radius,
    its segments have the lengths of straight flights inside the slab; tallies and counters are untouched; the three file formats."""
    cfg = smcrt.Config.load(RES / "validation1.toml")
    engine.apply(cfg)
    engine.set_track_history([0, 1])                      # the transmission detector (z = +0.01, facing +z) tracks histories
    n, seed = 20000, 3
    g = engine.trace_packets(n, seed)
    before = engine.fetch(absorb=True)
    ids, det, total = engine.history_hits()
    assert total == len(ids) and (det == 2).all() and (np.diff(ids.astype(np.int64)) > 0).all()
    assert total == before["det_bins"][101:].sum()         # every hit on detector 2, once
    # the oracle detects the same packets (same streams): compare through the per-packet fates it reports
    osc = oracle.OracleScene.from_toml(RES / "validation1.toml")
    o = osc.run(n, seed, per_packet=True, grids=False)
    assert abs(o["det_bins"][101:].sum() - total) <= 3
    v, nv, hv = engine.history_replay(ids, seed, max_vertices=64)
    assert (hv > 0).all() and (hv <= nv).all()
    after = engine.fetch(absorb=True)
    assert np.array_equal(before["det_bins"], after["det_bins"]) and np.array_equal(before["absorb"], after["absorb"])
    assert before["counters"] == after["counters"]
    src = np.array([0.0, 0.0, -0.01])
    for k in range(0, len(ids), max(len(ids) // 400, 1)):
        m = hv[k]
        if m > 64:
            continue
        p = v[k, :m]
        assert np.allclose(p[0, :3], src, atol=1e-6)                     # launch point
        assert m == g["nscatt"][int(ids[k])] + 2                         # launch + every interaction + the hit point
        assert abs(p[m - 1, 2] - 0.01) < 2e-6 and np.hypot(p[m - 1, 0], p[m - 1, 1]) <= 20.0 and p[m - 1, 3] == 2.0
        assert (np.abs(p[:m - 1, 2]) <= 0.01 + 1e-6).all()                # interactions happen inside the slab
    # the three writers (history_stack_t%write / %finish)
    from rsmcrt_b200 import _lib
    import ctypes as C
    L = _lib.load()
    cnt = np.minimum(hv, 64).astype(np.int32)
    for ext in ("obj", "ply", "json"):
        path = tmp_path / f"photPos_000.{ext}"
        _lib.check(L.smcrt_history_write(str(path).encode(), len(ids), 64, v.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_int32))))
        text = path.read_text()
        if ext == "obj":
            lines = text.splitlines()
            nvl = sum(1 for l in lines if l.startswith("v "))
            assert nvl == cnt.sum() and sum(1 for l in lines if l.startswith("l ")) == (cnt >= 2).sum()
            assert lines[0].startswith("v ") and lines[-1].startswith("l ") and int(lines[-1].split()[-1]) == nvl
            x, y, z = (float(t) for t in lines[0].split()[1:4])
            assert abs(z + 0.01) < 1e-6
        elif ext == "ply":
            assert f"element vertex {cnt.sum()}" in text and f"element edge {np.maximum(cnt - 1, 0).sum()}" in text
        else:
            import json
            d = json.loads(text)
            assert len(d) == len(ids) and len(d["0_0"]) == cnt[0] and len(d["0_0"][0]) == 3


def test_default_mcrt_writes_the_history_file(tmp_path, smcrt):
    """[[detectors]] trackHistory = true / historyFileName: default_MCRT leaves data/<name>_000.<ext> beside the other outputs
    (init_historyStack appends the 3-digit thread id, src/historyStack.f90:44-45)."""
    text = (RES / "validation1.toml").read_text().replace("nxg = 500", "nxg = 40").replace("nyg = 500", "nyg = 50").replace("nzg = 500", "nzg = 60")
    head, tail = text.rsplit("trackHistory=false", 1)
    text = head + 'trackHistory=true\nhistoryFileName="paths.json"' + tail
    toml = tmp_path / "v1hist.toml"
    toml.write_text(text)
    out = tmp_path / "data"
    pps, cn = smcrt.default_MCRT(toml, out_dir=out, nphotons=5000)
    import json
    d = json.loads((out / "paths_000.json").read_text())
    det2 = np.fromfile(out / "detectors" / "detector_2.dat", np.float64)
    assert len(d) > 0.6 * 5000 and all(len(v) >= 2 for v in d.values())
    assert abs(d["0_0"][0][2] + 0.01) < 1e-6 and abs(d["0_0"][-1][2] - 0.01) < 1e-5      # source -> transmission detector
    assert cn["launched"] == 5000 and len(det2) > 100
