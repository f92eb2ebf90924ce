"""Packet-id sharding of one job over ranks / GPUs.

A packet's random stream depends only on (seed, global packet id), so a job split over any number of ranks is
the same job: rank r of W runs the contiguous id block [offset, offset + n) of every step and the tallies are
summed once at the end (NCCL reduce; the intent of the dead mpi_reduce block, src/kernelsMod.f90:2351-2357).
"""
from __future__ import annotations


def step_offset(step: int, world: int, rank: int, n_per_rank: int) -> int:
    """First global packet id of `rank` in `step` when every rank runs n_per_rank packets per step (weak scaling)."""
    return (step * world + rank) * n_per_rank


def split_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of a fixed-size job of n_total packets for `rank` (strong scaling; same rule as smcrt_run over the
    GPUs of one context: lo = n*r/W)."""
    return n_total * rank // world, n_total * (rank + 1) // world
