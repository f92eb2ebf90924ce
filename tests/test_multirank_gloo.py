"""world_size-2 test of the N>1 host logic on CPU (gloo): packet-id sharding + one sum-reduce of the tallies at the end
reproduces the single-rank job.  The per-rank transport is done by the CPU oracle here (no GPU); on the GPU box the same
sharding helper drives the engine and the reduce is ncclReduce inside libsmcrt_gpu.so (bench.py, smcrt_comm_reduce)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n_per_rank, steps, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rsmcrt_b200 as R
    from rsmcrt_b200.sharding import step_offset
    from oracle import binding as O
    cfg = R.Config.load(ROOT / "res" / "validation1.toml")
    osc = O.OracleScene.from_config(cfg)
    bins = np.zeros(202)
    nsc = 0.0
    for s in range(steps):
        r = osc.run(n_per_rank, cfg.iseed, id_offset=step_offset(s, world, rank, n_per_rank), grids=False, nthreads=2)
        bins += r["det_bins"]
        nsc += r["counters"]["nscatt"]
    t = torch.tensor(np.concatenate([bins, [nsc]]))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)   # the single reduce at the end of the job
    if rank == 0:
        np.save(Path(out_dir) / "reduced.npy", t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_job(tmp_path):
    import torch.multiprocessing as mp
    n, steps, world = 20000, 2, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, steps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "reduced.npy")
    sys.path.insert(0, str(ROOT))
    import rsmcrt_b200 as R
    from oracle import binding as O
    cfg = R.Config.load(ROOT / "res" / "validation1.toml")
    ref = O.OracleScene.from_config(cfg).run(n * steps * world, cfg.iseed, grids=False)
    # packet streams depend only on (seed, id): identical integer-valued tallies, whatever the split
    assert (got[:202] == ref["det_bins"]).all()
    assert got[202] == ref["counters"]["nscatt"]


def test_sharding_helpers():
    from rsmcrt_b200.sharding import split_range, step_offset
    n, w = 1_000_003, 8
    cover = []
    for r in range(w):
        lo, hi = split_range(n, w, r)
        cover.append((lo, hi))
    assert cover[0][0] == 0 and cover[-1][1] == n
    assert all(cover[i][1] == cover[i + 1][0] for i in range(w - 1))
    ids = sorted(step_offset(s, 4, r, 10) for s in range(3) for r in range(4))
    assert ids == list(range(0, 120, 10))
