"""Path-length mode A/B on the GPU box: run walker (difference grids) vs legacy voxel walker (one red per voxel crossed).

    python tools/pl_ab.py [deck:packets ...]   -> one JSON line per (deck, walker)
"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np  # noqa: E402
import rsmcrt_b200 as R  # noqa: E402
import make_vessels  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
cases = [a.split(":") for a in sys.argv[1:]] or [["validation1.toml", "1e8"], ["sphere.toml", "5e7"], ["skin_b200.toml", "2e7"],
                                                  ["lens.toml", "5e7"], ["vessels.toml", "2e7"], ["scat_test.toml", "1e7"]]
tmp = Path(tempfile.mkdtemp())
make_vessels.make(tmp, 240, 7)
for deck, n in cases:
    n = int(float(n))
    cfg = R.Config.load(ROOT / "res" / deck, res_dir=tmp) if deck == "vessels.toml" else R.Config.load(ROOT / "res" / deck)
    sums = {}
    for legacy in (0, 1):
        if legacy:
            os.environ["SMCRT_DDA_LEGACY"] = "1"
        else:
            os.environ.pop("SMCRT_DDA_LEGACY", None)
        e = R.Engine(1)
        e.apply(cfg)
        ms = []
        for _ in range(3):
            e.reset_tallies()
            e.run(n, cfg.iseed, tally_mode=3)
            ms.append(e.last_run_ms)
        t0 = time.perf_counter()
        out = e.fetch(jmean=True, absorb=False)
        fetch_ms = (time.perf_counter() - t0) * 1e3
        sums[legacy] = float(out["jmean"].sum(dtype=np.float64))
        print(json.dumps({"deck": deck, "walker": "voxels" if legacy else "runs", "packets": n, "ms": [round(m, 2) for m in ms],
                          "packets_per_s": n / min(ms[1:]) * 1e3, "fetch_ms": round(fetch_ms, 2),
                          "kernel_variant": e.kernel_variant(3), "segment_mode": e.segment_mode, "segs_per_packet": round(e.segments_per_packet, 2), "lost": int(out["counters"]["lost"])}), flush=True)
        e.close()
    print(json.dumps({"deck": deck, "jmean_sum_rel_diff": sums[0] / sums[1] - 1.0}), flush=True)
