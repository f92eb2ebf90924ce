import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from oracle import binding as O
cfg = R.Config.load("res/validation1.toml")
n = 20000
e = R.Engine(1); e.apply(cfg)
osc = O.OracleScene.from_config(cfg)
for seed in range(1, 13):
    e.reset_tallies()
    e.run(n, seed, tally_mode=3)
    jg3 = e.fetch(jmean=True, absorb=False)["jmean"].astype(np.float64)
    jo3 = osc.run(n, seed, tally_mode=3)["jmean"].astype(np.float64)
    jg = jg3.sum(axis=(0, 1)); jo = jo3.sum(axis=(0, 1))
    d = jg3[:, :, 416] - jo3[:, :, 416]
    k = np.unravel_index(np.argmax(np.abs(d)), d.shape)
    print(seed, "416: %.4f %.4f d=%+.4f | 84: %.4f %.4f d=%+.4f | tot d=%+.4f | max cell diff in 416 at %s = %+.4f (n cells |d|>1e-4: %d)" % (
        jg[416], jo[416], jg[416] - jo[416], jg[84], jo[84], jg[84] - jo[84], jg.sum() - jo.sum(), k, d[k], (np.abs(d) > 1e-4).sum()))
