"""GPU-box check (>= 2 GPUs): one job on 1 GPU == the same job sharded over all GPUs of one process (NCCL reduce at fetch)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
import torch
n_gpu = torch.cuda.device_count()
cfg = R.Config.load("res/validation1.toml")
N = 4_000_000
res = []
for g in (1, n_gpu):
    e = R.Engine(g)
    e.apply(cfg)
    e.run(N, cfg.iseed)
    out = e.fetch(absorb=True)
    print(f"gpus={g} ms={e.last_run_ms:.2f} absorbed={out['absorb'].sum():.0f} Rd={out['det_bins'][:101].sum()/N:.5f} Tt={out['det_bins'][101:].sum()/N:.5f} "
          f"nscatt={out['counters']['nscatt']:.0f} launched={out['counters']['launched']:.0f}")
    res.append(out)
    e.close()
a, b = res
assert (a["det_bins"] == b["det_bins"]).all(), "detector bins must be bit-identical (fixed-point integer tallies)"
assert a["counters"]["nscatt"] == b["counters"]["nscatt"]
assert np.abs(a["absorb"] - b["absorb"]).max() == 0.0, "unit deposits are exact in FP32 below 2^24"
print("OK: 1-GPU and %d-GPU tallies identical" % n_gpu)
