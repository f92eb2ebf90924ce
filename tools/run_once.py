"""One run of one deck (GPU box): python tools/run_once.py deck packets [tally_mode=1]   (pin a kernel with SMCRT_VARIANT_FORCE=k).
The command ncu captures: a single trace-kernel launch, no trial."""
import os
import sys
sys.path.insert(0, ".")
import rsmcrt_b200 as R

name, n = sys.argv[1], int(float(sys.argv[2]))
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
e = R.Engine(1)
e.apply(R.Config.load("res/" + name))
e.run(n, 5, tally_mode=mode)
out = e.fetch(jmean=bool(mode & 2), absorb=True)
c = out["counters"]
print(name, "mode", mode, "variant", os.environ.get("SMCRT_VARIANT_FORCE"), "ms", round(e.last_run_ms, 3), "packets/s", f"{n / e.last_run_ms * 1e3:.4g}",
      "sweeps/packet", round(c["sweeps"] / n, 3), "lost", c["lost"], "absorbed", float(out["absorb"].sum()), "detected", float(out["det_bins"].sum()),
      "path/packet", float(out["jmean"].sum(dtype="f8")) / n if mode & 2 else None, flush=True)
