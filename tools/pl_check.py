"""Path-length bookkeeping: in an empty (kappa = 0) box every packet deposits exactly its chord from the source to the grid boundary."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
scene = A.Scene.from_primitives([(A.BOX, None, [1.0, 1.0, 1.0])], [(0.0, 0.0, 0.0, 1.0)])
p = np.zeros(24); p[0:3] = [0.1, -0.2, 0.3]
e = R.Engine(1)
e.set_grid(200, 200, 200, 1.0, 1.0, 1.0); e.set_scene(scene); e.set_source(A.SRC_POINT, 0, p)
n = 2_000_000
g = e.trace_packets(n, 5, tally_mode=A.TALLY_PATHLENGTH)
j = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
expect = np.linalg.norm(g["pos"] - p[:3], axis=1).sum()
print("agg" if os.environ.get("SMCRT_DDA_AGG") else "plain", "ms", e.last_run_ms, "jmean.sum", j.sum(), "expected", expect, "rel", j.sum() / expect - 1)
