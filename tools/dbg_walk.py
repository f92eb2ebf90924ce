"""usage: SMCRT_LIB=rsmcrt_b200/lib/libsmcrt_dbg.so SMCRT_DEBUG_PID=<pid> python tools/dbg_walk.py <deck> <seed> <pid>  (lib built with -DSMCRT_DBG_WALK)"""
import sys
sys.path.insert(0, ".")
import rsmcrt_b200 as R
cfg = R.Config.load("res/" + sys.argv[1])
seed, pid = int(sys.argv[2]), int(sys.argv[3])
e = R.Engine(1); e.apply(cfg)
g = e.trace_packets(1, seed, id_offset=pid, tally_mode=3)
print({k: v[0] for k, v in g.items()})
