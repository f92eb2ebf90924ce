"""Find single packets whose deposited path (sum of jmean) differs between engine and oracle although the histories agree."""
import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from oracle import binding as O
cfg = R.Config.load("res/validation1.toml")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 10
e = R.Engine(1); e.apply(cfg)
osc = O.OracleScene.from_config(cfg)

def tot(lo, n):
    e.reset_tallies(); e.run(n, seed, id_offset=lo, tally_mode=3)
    jg = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
    jo = osc.run(n, seed, id_offset=lo, tally_mode=3)["jmean"].astype(np.float64)
    return jg, jo

found = []
def rec(lo, n, depth):
    jg, jo = tot(lo, n)
    d = jg.sum() - jo.sum()
    if abs(d) < 2e-4:
        return
    if n == 1:
        g = e.trace_packets(1, seed, id_offset=lo, tally_mode=3)
        o = osc.run(1, seed, id_offset=lo, tally_mode=3, per_packet=True, grids=False)
        dz = jg.sum(axis=(0, 1)) - jo.sum(axis=(0, 1))
        ks = np.nonzero(np.abs(dz) > 1e-6)[0]
        print("pid", lo, "d=%+.5f" % d, "gpu", g["fate"][0], g["nscatt"][0], g["events"][0], g["pos"][0], "| orc", o["fate"][0], o["nscatt"][0],
              o["events"][0], o["pos"][0], "| z-slabs", [(int(k), float("%.3g" % dz[k])) for k in ks[:8]], flush=True)
        found.append(lo)
        return
    if len(found) >= 6:
        return
    step = max(1, n // 8)
    for a in range(lo, lo + n, step):
        rec(a, min(step, lo + n - a), depth + 1)

rec(0, 4096, 0)
