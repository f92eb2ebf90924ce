"""Rank process of test_n_gpu_equals_one_gpu: python two_rank_worker.py RANK WORLD DIR N SEED MODE (one GPU per process)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rsmcrt_b200 as R  # noqa: E402
from rsmcrt_b200.sharding import split_range  # noqa: E402

rank, world, d, n, seed, mode = int(sys.argv[1]), int(sys.argv[2]), Path(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
e = R.Engine(1)
e.apply(R.Config.load(ROOT / "res" / "validation1.toml"))
uid_file = d / "uid.bin"
if rank == 0:
    tmp = d / "uid.tmp"
    tmp.write_bytes(R.Engine.comm_unique_id())
    tmp.rename(uid_file)
else:
    for _ in range(600):
        if uid_file.exists():
            break
        time.sleep(0.05)
e.comm_init(world, rank, uid_file.read_bytes())
lo, hi = split_range(n, world, rank)
e.run(hi - lo, seed, id_offset=lo, tally_mode=mode)
e.comm_reduce(0)
if rank == 0:
    out = e.fetch(jmean=True, absorb=True)
    np.savez(d / "rank0.npz", det_bins=out["det_bins"], absorb=out["absorb"], jmean=out["jmean"], nscatt=out["counters"]["nscatt"],
             launched=out["counters"]["launched"])
e.close()
