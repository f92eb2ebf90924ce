"""A/B of environment switches on the GPU box, every setting in a process of its own (the engine reads its switches once).

    python tools/env_ab.py deck packets mode "A=1 B=2" "A=0" ...   -> one JSON line per setting ("-" = no switches)

Three runs per setting (the first one holds the trial and calibration); jmean_sum / absorbed let one check that a switch does not
change the result.
"""
import json
import os
import subprocess
import sys

WORKER = r"""
import json, sys
sys.path.insert(0, ".")
import numpy as np
import rsmcrt_b200 as R
name, n, mode = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3])
e = R.Engine(1)
e.apply(R.Config.load("res/" + name))
ms = []
for _ in range(3):
    e.reset_tallies()
    e.run(n, 5, tally_mode=mode)
    ms.append(round(e.last_run_ms, 2))
out = e.fetch(jmean=bool(mode & 2), absorb=True)
print(json.dumps({"ms": ms, "packets_per_s": n / min(ms[1:]) * 1e3, "variant": e.kernel_variant(mode), "segment_mode": e.segment_mode,
                  "jmean_sum": float(out["jmean"].sum(dtype=np.float64)) if mode & 2 else None, "absorbed": float(out["absorb"].sum(dtype=np.float64)),
                  "lost": int(out["counters"]["lost"])}))
"""

deck, n, mode = sys.argv[1:4]
for setting in sys.argv[4:]:
    env = dict(os.environ)
    if setting != "-":
        env.update(kv.split("=", 1) for kv in setting.split())
    res = subprocess.run([sys.executable, "-c", WORKER, deck, n, mode], env=env, capture_output=True, text=True, timeout=600)
    line = res.stdout.strip().splitlines()[-1] if res.returncode == 0 and res.stdout.strip() else json.dumps({"error": res.stderr[-400:]})
    print(json.dumps({"deck": deck, "packets": n, "mode": int(mode), "env": setting, **json.loads(line)}), flush=True)
