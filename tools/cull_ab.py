import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
import make_vessels
make_vessels.make("/tmp/vess", 240, 7)
for name, n, mode, kw in (("sphere.toml", 1_000_000, 3, {}), ("vessels.toml", 1_000_000, 1, {"res_dir": "/tmp/vess"}), ("sphere.toml", 10_000_000, 3, {})):
    cfg = R.Config.load("res/" + name, **kw)
    e = R.Engine(1); e.apply(cfg)
    e.run(100000, 1, tally_mode=mode); e.reset_tallies()
    e.run(n, cfg.iseed, tally_mode=mode)
    out = e.fetch(jmean=True, absorb=True)
    c = out["counters"]
    print(f"{name} cull={'off' if os.environ.get('SMCRT_NO_CULL') else 'on'} n={n:.0e} ms={e.last_run_ms:.1f} pkt/s={n/e.last_run_ms*1e3:.3e} sweeps/pkt={c['sweeps']/n:.1f} "
          f"nscatt/pkt={c['nscatt']/n:.3f} lost={c['lost']:.0f} absorb={out['absorb'].sum():.0f} jmean={out['jmean'].sum():.1f}", flush=True)
    e.close()
