"""Small runs of every kernel family in one process, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`."""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A

for deck, n, mode in (("validation1.toml", 60000, 3), ("sphere.toml", 30000, 7), ("skin_b200.toml", 20000, 3), ("test_dects.toml", 20000, 1),
                      ("omg.toml", 10000, 3)):
    for inline in ("0", "1"):
        os.environ["SMCRT_SEG_INLINE"] = inline
        e = R.Engine(1)
        e.apply(R.Config.load("res/" + deck))
        e.run(n, 3, tally_mode=mode)
        e.run(n, 3, id_offset=n, tally_mode=mode, survival_bias=True)
        out = e.fetch(jmean=True, absorb=True, emission=True)
        print(deck, inline, out["counters"]["launched"], float(out["jmean"].sum()), flush=True)
        e.reset_tallies()
        e.close()
e = R.Engine(1)
e.apply(R.Config.load("res/validation1.toml"))
e.set_track_history([0, 1])
g = e.trace_packets(5000, 3)
ids, det, total = e.history_hits()
v, nv, hv = e.history_replay(ids, 3, max_vertices=32)
tot, layer = e.run_sources(np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.005]]), 2000, 5)
table, best = e.inverse_mcrt(1, 2, [0.09739, 0.66096], 3, 20000, 9)
print("history", total, int(hv.max()), "sources", tot.sum(), "inverse", best, flush=True)
e.close()
print("done")
