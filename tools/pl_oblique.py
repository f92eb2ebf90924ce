import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
from oracle import binding as O
text = open("res/validation1.toml").read().replace("mus = [90]", "mus = [0.0]").replace("mua = [10]", "mua = [0.0]")
text = text.replace('direction = "z"', 'direction = [0.6, 0.0, 0.8]').replace("position = [0.0,0.0,-0.01]", "position = [0.0,0.0,-0.012]")
cfg = R.Config.loads(text)
n = 1000
e = R.Engine(1); e.apply(cfg)
e.run(n, 1, tally_mode=3)
jg = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64).sum(axis=(0, 1)) / n
jo = O.OracleScene.from_config(cfg).run(n, 1, tally_mode=3)["jmean"].astype(np.float64).sum(axis=(0, 1)) / n
exp = 0.03 / 500 / 0.8
print("expected per voxel", exp)
for k in (49, 50, 51, 82, 83, 84, 85, 200, 415, 416, 417, 418, 498, 499):
    print(k, "gpu %.6e  oracle %.6e   gpu/exp %.4f oracle/exp %.4f" % (jg[k], jo[k], jg[k] / exp, jo[k] / exp))
print("totals", jg.sum(), jo.sum(), "expected", (0.015 + 0.012) / 0.8)
