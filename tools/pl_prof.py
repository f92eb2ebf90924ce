import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from oracle import binding as O
cfg = R.Config.load("res/validation1.toml")
n, seed = 20000, 3
e = R.Engine(1); e.apply(cfg)
osc = O.OracleScene.from_config(cfg)
e.run(n, seed, tally_mode=3)
jg3 = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
jo3 = osc.run(n, seed, tally_mode=3)["jmean"].astype(np.float64)
dz = (jg3 - jo3).sum(axis=(0, 1))
order = np.argsort(-np.abs(dz))[:12]
print("total diff", dz.sum())
for k in order:
    d = jg3[:, :, k] - jo3[:, :, k]
    c = np.unravel_index(np.argmax(np.abs(d)), d.shape)
    print(int(k), "dz=%+.5f" % dz[k], "gpu %.5f" % jg3[:, :, k].sum(), "max cell", c, "%+.5f" % d[c], "central4 %+.5f" % d[249:251, 249:251].sum())
print("cum diff by region: z<83 %+.5f, 83..416 %+.5f, >416 %+.5f" % (dz[:83].sum(), dz[83:417].sum(), dz[417:].sum()))
