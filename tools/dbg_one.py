import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
cfg = R.Config.load("res/" + sys.argv[1])
pid = int(sys.argv[2])
e = R.Engine(1); e.apply(cfg)
g = e.trace_packets(1, cfg.iseed, id_offset=pid, tally_mode=1)
print({k: v[0] for k, v in g.items()})
