import sys, time; sys.path.insert(0, ".")
import numpy as np
import rsmcrt_b200 as R
cfg = R.Config.load("res/validation1.toml"); e = R.Engine(1); e.apply(cfg)
scene = cfg.scene; kind, dp, nb, _ = cfg.detectors; sk, ss, sp = cfg.source
n = 100_000_000
for it in range(2): e.run(n, 1)
e.reset_tallies()
h = np.zeros(e.n_voxels, np.float32); hb = np.zeros(max(e.det_bins_total, 1)); e.pin_host(h); e.pin_host(hb)
def T(label, f):
    t = time.perf_counter(); r = f(); print("%-16s %8.2f ms" % (label, (time.perf_counter() - t) * 1e3), flush=True); return r
for it in range(3):
    print("--- step", it)
    T("set_scene", lambda: e.set_scene(scene)); T("set_source", lambda: e.set_source(sk, ss, sp)); T("set_detectors", lambda: e.set_detectors(kind, dp, nb))
    T("run", lambda: e.run(n, 2)); print("   kernel ms", e.last_run_ms, "launches", e.launch_count)
    T("fetch(acc)", lambda: e.fetch_into(absorb=h, det_bins=hb, accumulate=True)); print("   bytes", e.last_fetch_bytes)
    T("reset", lambda: e.reset_tallies())
T("fetch(dense-set)", lambda: e.fetch_into(absorb=h, det_bins=hb, accumulate=False))
