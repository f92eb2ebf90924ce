"""GPU-box diagnostic: which packets does the engine give up on, and why (reason codes in kernels.cuh)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
from oracle import binding as O

for name, n in (("scat_test.toml", 400000), ("validation2.toml", 4000), ("validation1.toml", 400000), ("sphere.toml", 100000)):
    cfg = R.Config.load("res/" + name)
    e = R.Engine(1)
    e.apply(cfg)
    mode = A.TALLY_ABSORB | (A.TALLY_PATHLENGTH if name == "sphere.toml" else 0)
    g = e.trace_packets(n, cfg.iseed, tally_mode=mode)
    lost = np.nonzero(g["fate"] == 3)[0]
    c = e.fetch(absorb=False)["counters"]
    print(f"== {name}: n={n} ms={e.last_run_ms:.2f} lost={len(lost)} sweeps/packet={c['sweeps']/n:.1f} nscatt/packet={c['nscatt']/n:.3f} "
          f"fates={np.bincount(g['fate'], minlength=4)}")
    if len(lost):
        why = -g["events"][lost]
        print("   reasons:", np.bincount(why, minlength=6))
        osc = O.OracleScene.from_config(cfg)
        for i in lost[:8]:
            o = osc.run(1, cfg.iseed, id_offset=int(i), per_packet=True, grids=False, tally_mode=mode)
            print(f"   id={i} why={-g['events'][i]} pos={g['pos'][i]} nsc={g['nscatt'][i]} | oracle fate={o['fate'][0]} nsc={o['nscatt'][0]} "
                  f"pos={o['pos'][0]} sweeps={o['counters']['sweeps']}")
    pass
    print("   sweeps per packet: mean %.1f  p99 %d  max %d   top5 %s" % (g["sweeps"].mean(), np.percentile(g["sweeps"], 99), g["sweeps"].max(), np.sort(g["sweeps"])[-5:]))
