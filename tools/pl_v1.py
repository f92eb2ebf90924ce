import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
from oracle import binding as O
cfg = R.Config.load("res/validation1.toml")
e = R.Engine(1); e.apply(cfg)
osc = O.OracleScene.from_config(cfg)
n, seed, mode = 50000, 21, 3
e.run(n, seed, tally_mode=mode)
jg = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
jo = osc.run(n, seed, tally_mode=mode)["jmean"].astype(np.float64)
zg, zo = jg.sum(axis=(0, 1)), jo.sum(axis=(0, 1))
i = np.argsort(-np.abs(zg - zo))[:8]
print("tot", jg.sum(), jo.sum())
for k in sorted(i): print(k, "z=%.6f" % ((k + 0.5) * 0.03 / 500 - 0.015), zg[k], zo[k], zg[k] - zo[k])
print("slabs near source:", [(k, round(zg[k], 4), round(zo[k], 4)) for k in range(80, 88)])
print("slabs near top:", [(k, round(zg[k], 4), round(zo[k], 4)) for k in range(414, 420)])
