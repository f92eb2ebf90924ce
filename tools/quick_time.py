"""GPU-box timing probe: photons/s per scene and tally mode (CUDA-event time of the persistent kernel)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A

cases = [("sphere.toml", 100_000, A.TALLY_ABSORB), ("sphere.toml", 100_000, A.TALLY_ABSORB | A.TALLY_PATHLENGTH),
         ("sphere.toml", 1_000_000, A.TALLY_ABSORB | A.TALLY_PATHLENGTH),
         ("validation1.toml", 10_000_000, A.TALLY_ABSORB), ("scat_test.toml", 4_000_000, A.TALLY_ABSORB),
         ("validation2.toml", 200_000, A.TALLY_ABSORB)]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]))]
for name, n, mode in cases:
    cfg = R.Config.load("res/" + name)
    e = R.Engine(1)
    e.apply(cfg)
    e.run(min(n, 100000), 1, tally_mode=mode)  # warm-up
    e.reset_tallies()
    e.run(n, cfg.iseed, tally_mode=mode)
    ms = e.last_run_ms
    c = e.fetch(absorb=False)["counters"]
    print(f"{name:18s} mode={mode} n={n:.0e} ms={ms:9.2f} photons/s={n/ms*1e3:.3e} sweeps/pkt={c['sweeps']/n:7.1f} "
          f"nscatt/pkt={c['nscatt']/n:7.3f} bounces/pkt={c['bounces']/n:6.3f} lost={c['lost']:.0f} dethits/pkt={c['det_hits']/n:.3f}", flush=True)
    e.close()
