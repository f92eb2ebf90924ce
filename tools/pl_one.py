import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from oracle import binding as O
cfg = R.Config.load("res/validation1.toml")
seed, pid, k0 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
e = R.Engine(1); e.apply(cfg)
osc = O.OracleScene.from_config(cfg)
e.run(1, seed, id_offset=pid, tally_mode=3)
jg = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
jo = osc.run(1, seed, id_offset=pid, tally_mode=3)["jmean"].astype(np.float64)
print("totals", jg.sum(), jo.sum())
for k in range(k0 - 4, k0 + 5):
    cg = np.argwhere(jg[:, :, k] > 0); co = np.argwhere(jo[:, :, k] > 0)
    print(k, "gpu %.6e (%d cells) oracle %.6e (%d cells)" % (jg[:, :, k].sum(), len(cg), jo[:, :, k].sum(), len(co)),
          [(int(i), int(j), float("%.4g" % jg[i, j, k])) for i, j in cg[:6]], [(int(i), int(j), float("%.4g" % jo[i, j, k])) for i, j in co[:6]])
