#!/usr/bin/env python
"""bench.py — photon packets/s of the run_MCRT hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (restated oracle) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL reduce of the tallies

A "step" is one pass of the hot path over one batch of packets of the workload scene (default: the slab validation
res/validation1.toml, BASELINE configs[1], 1e8 packets per step per GPU).  `value` = packets of all ranks / device time
(CUDA events on the launch stream, max over ranks) with the scene resident in HBM; `e2e` = the same through the
public API with host buffers: scene/source/detector upload + run + download of the absorb grid, detector bins and
counters inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "photon_packets_per_s"
UNIT = "packets/s"

# ---- frozen per-operation table for the algorithmic work per packet: minimal FP32 op counts of the
# reference formulas (add=mul=cmp=abs=min/max=1, fma=2, sqrt=div=1, sin/cos/log/atan=1), identical for CPU and GPU.
F_AFFINE = 18
F_PRIM = {1: 25, 2: 37, 3: 28, 4: 70, 5: 23, 6: 47, 7: 47, 8: 80, 9: 45, 10: 23}
F_FRESNEL, F_HG, F_DET_CIRCLE, F_VOXEL, F_EMIT = 45, 60, 30, 9, 20
WORKLOADS = {
    "validation1.toml": "BASELINE configs[1]: slab validation, pencil beam, 500^3 grid, 2 circle detectors",
    "sphere.toml": "BASELINE configs[0]: 40 spheres n=1.37 in a box, uniform source, 200^3 grid, default build (absorb tallies)",
    "skin_b200.toml": "BASELINE configs[2]: five refractive tissue layers, uniform source, 200^3 grid",
    "lens.toml": "BASELINE configs[3]: refractive lens (model of two spheres), uniform source, 200^3 grid",
    "vessels.toml": "BASELINE configs[4]: vessel tree (240 capsules, seeded synthetic data in the reference's file formats) in a dermis box, 200^3 grid",
}


def deck_res_dir(scene_name: str):
    """vessels.toml reads edges / nodes / radii.dat, which the reference does not ship: a seeded synthetic tree is
    written to a scratch directory in the reference's file formats (tools/make_vessels.py), the same on every rank."""
    if scene_name != "vessels.toml":
        return None
    import tempfile
    sys.path.insert(0, str(ROOT / "tools"))
    import make_vessels
    d = Path(tempfile.mkdtemp(prefix="smcrt_vessels_"))
    make_vessels.make(d, 240, 7)
    return d
# What an ncu --set full capture of this command says about the dominant kernel (DRAM bytes of one launch, issue-slot utilisation,
# lanes per instruction), keyed on (scene, packets per step, kernel variant): profiles/bench_ncu.json, written from the capture by
# tools/ncu_bench_json.py.  A run whose kernel variant has no capture reports null, not a stale constant.
# Likewise the flops/packet of the REFERENCE algorithm (oracle counters x the op table above): measured by the N=1 run's CPU leg;
# N>1 runs (no CPU leg) take the value the last N=1 run of the scene printed, from the same file.
def ncu_facts(scene_name: str, n_step: int, variant: int) -> dict:
    try:
        table = json.loads((ROOT / "profiles" / "bench_ncu.json").read_text())
    except (OSError, ValueError):
        return {}
    return table.get(f"{scene_name}:{n_step}:{variant}", {})


def committed_reference_flops(scene_name: str):
    try:
        return json.loads((ROOT / "profiles" / "bench_ncu.json").read_text()).get("reference_flops_per_packet", {}).get(scene_name)
    except (OSError, ValueError):
        return None


def flops_per_sweep(scene) -> float:
    """One evaluation of ALL top-level SDFs (what `cnts += N` counts, src/inttau2.f90:67)."""
    total = 0.0
    kinds = scene.kind
    for k in kinds:
        if int(k) in F_PRIM:
            total += F_PRIM[int(k)] + F_AFFINE
        else:
            total += 8  # csg op / modifier arithmetic
    return total


def algorithmic_flops_per_packet(scene, n_det: int, c: dict, n: float) -> float:
    """W_flop of SURVEY §8(d) from measured expectations (counters of a run over n packets)."""
    sweeps = c["sweeps"] / n
    nscatt = c["nscatt"] / n
    fres = c.get("fresnel_events", c["bounces"]) / n
    w = sweeps * flops_per_sweep(scene)
    w += fres * (F_FRESNEL + 4 * max(F_PRIM.get(int(scene.kind[0]), 30) + F_AFFINE, 1))
    w += nscatt * F_HG
    # straight segments tested against detectors: one per tauint2 call + one per crossing ~ (nscatt + 2)
    w += (nscatt + 2.0) * n_det * F_DET_CIRCLE
    w += F_EMIT + F_VOXEL
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (recipe in B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """Host cores this process may use: the affinity mask.  NOT OMP_NUM_THREADS: torch.distributed.run exports OMP_NUM_THREADS=1
    to every rank, which collapsed the reference arm to one core at N > 1."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        n = max(1, os.cpu_count() or 1)
    return n   # handed to the oracle as num_threads(n), which overrides the environment's OMP_NUM_THREADS


def cpu_leg(deck_path, seconds_target: float, threads: int = 0):
    """Time the oracle (CPU restatement of the reference path) on a bounded sample of the same workload."""
    threads = threads or host_threads()
    from oracle import binding as O
    O.build()
    osc = O.OracleScene.from_toml(deck_path, deck_res_dir(Path(deck_path).name))
    cfg = osc.deck
    nv = int(np.prod(cfg.grid[0]))
    # calibrate on a small sample, then size the timed sample for ~seconds_target of CPU work
    r = osc.run(20000, cfg.iseed, rng_mode=1, nthreads=threads, grids=False, tally_mode=1)
    rate = 20000 / max(r["seconds"], 1e-6)
    n = int(min(max(rate * seconds_target, 50_000), 50_000_000))
    r = osc.run(n, cfg.iseed + 1, rng_mode=1, nthreads=threads, grids=(nv <= 64_000_000), tally_mode=1)
    return n, r["seconds"], r["counters"], threads


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores.  The Fortran
    binary cannot be built in this image (no compiler, un-vendored deps: SURVEY F2/F3), so this is the oracle port."""
    if rank != 0:
        return
    # the oracle's own TOML -> scene path (oracle/scenes.py): this arm never loads the product library
    threads = host_threads()
    from oracle import binding as O
    O.build()
    osc = O.OracleScene.from_toml(ROOT / "res" / args.scene, deck_res_dir(args.scene))
    cfg = osc.deck
    r = osc.run(20000, cfg.iseed, rng_mode=1, nthreads=threads, grids=False)
    rate = 20000 / max(r["seconds"], 1e-6)
    per_step = int(min(max(rate * args.cpu_seconds, 20_000), 20_000_000))
    for w in range(args.warmup):
        osc.run(max(per_step // 10, 1000), cfg.iseed + 10 + w, rng_mode=1, nthreads=threads, grids=False)
    secs = 0.0
    for s in range(args.steps):
        r = osc.run(per_step, cfg.iseed + 100 + s, rng_mode=1, nthreads=threads, grids=False)
        secs += r["seconds"]
    value = per_step * args.steps / secs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"res/{args.scene} ({WORKLOADS.get(args.scene, 'shipped input deck')})", "packets_per_step": per_step,
                   "note": "CPU restatement of the reference path (oracle/oracle.cpp, OpenMP); the gfortran/fpm binary cannot be built here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} packets/step x {args.steps} steps of res/{args.scene}, xoshiro256** per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # this arm builds its scene with oracle/scenes.py and must not have loaded the product's shared object
        "product_library_loaded": any("libsmcrt_gpu" in l for l in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }
    emit(line)


_STDOUT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded along the way write there too (NCCL prints its version banner to
    stdout at communicator set-up): everything but the final line goes to stderr."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scene", default="validation1.toml")
    ap.add_argument("--photons", type=float, default=1e8, help="packets per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per cpu_baseline sample / reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` array (path-length mode, sphere.toml)")
    ap.add_argument("--pathlength", action="store_true", help="also accumulate path-length fluence (jmean)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import rsmcrt_b200 as R
    from rsmcrt_b200 import api as A
    from rsmcrt_b200.sharding import step_offset
    R.load()  # fails loudly if the CUDA library is missing: no CPU fallback

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = R.Config.load(ROOT / "res" / args.scene, res_dir=deck_res_dir(args.scene))
    scene = cfg.scene
    kind, dp, nb, _ids = cfg.detectors
    eng = R.Engine(1, device_ids=[local_rank])
    eng.apply(cfg)
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(R.Engine.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().tolist()))

    n_step = int(args.photons)
    mode = A.TALLY_ABSORB | (A.TALLY_PATHLENGTH if args.pathlength else 0)
    seed = cfg.iseed
    nv = int(np.prod(cfg.grid[0]))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up (kernel, and the NCCL communicator: its first collective builds the NVLink channels)
    # full-size steps: the first large run of a scene is where the engine times its kernel variants (smcrt_kernel_variant)
    for w in range(args.warmup):
        eng.run(n_step, seed, id_offset=0, tally_mode=mode)
    if world > 1:
        eng.comm_reduce(0)
    eng.reset_tallies()

    # ---------------- timed region 1: device-resident (value)
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for s in range(args.steps):
        off = step_offset(s, world, rank, n_step)
        eng.run_async(n_step, seed, id_offset=off, tally_mode=mode)
        eng.wait()
        dev_ms += eng.last_run_ms
    red_ms = 0.0
    if world > 1:  # the single NCCL reduce of the tallies at the end of the job (part of the timed region)
        t0 = time.perf_counter()
        eng.comm_reduce(0)
        red_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count - launches0
    total_ms = max_over_ranks(dev_ms + red_ms)
    kernel_ms = max_over_ranks(dev_ms)
    res = eng.fetch(absorb=False) if rank == 0 else None
    packets = n_step * args.steps * world
    value = packets / (total_ms * 1e-3)

    # ---------------- timed region 2: end to end through the public API with host buffers (e2e)
    eng.reset_tallies()
    h2d = scene.kind.nbytes + scene.first_child.nbytes + scene.n_child.nbytes + scene.xform.nbytes + scene.params.nbytes + \
        scene.top_node.nbytes + 4 * scene.mus.nbytes + 24 * 8 + (kind.nbytes + dp.nbytes + nb.nbytes)
    d2h_actual = 0
    src_k, src_s, src_p = cfg.source
    # host result buffers of the caller (the reference's module arrays), page-locked once outside the timed region
    h_absorb = np.zeros(nv, np.float32)
    h_bins = np.zeros(max(eng.det_bins_total, 1))
    eng.pin_host(h_absorb)
    eng.pin_host(h_bins)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        off = step_offset(s, world, rank, n_step)
        eng.set_scene(scene)                       # host -> device: flattened scene
        eng.set_source(src_k, src_s, src_p)
        eng.set_detectors(kind, dp, nb)            # (also zeroes the detector tallies, like the escape driver's reset)
        eng.run(n_step, seed + 1, id_offset=off, tally_mode=mode)
        if world > 1:
            eng.comm_reduce(0)                     # every step produces the COMBINED result: tallies of all ranks summed on rank 0
        # device -> host, the sequence of the Fortran shim: module arrays += device tallies, then reset
        if rank == 0:
            cn = eng.fetch_into(absorb=h_absorb, det_bins=h_bins, accumulate=True)
            d2h_actual += eng.last_fetch_bytes
            _ = float(h_bins.sum()) + cn["nscatt"]
        eng.reset_tallies()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = packets / e2e_s
    # host array accumulated over the e2e steps: at N > 1 rank 0 holds the tallies of all ranks
    absorbed_per_packet = float(h_absorb.sum(dtype=np.float64)) / (n_step * args.steps * (world if rank == 0 else 1))
    eng.unpin_host(h_absorb)
    eng.unpin_host(h_bins)

    # ---------------- the other BASELINE modes / scenes: path-length mode of the workload scene, sphere.toml both modes
    extra = []
    red_peaks = {}
    try:
        red_peaks = json.loads((ROOT / "profiles" / "r01_red_peaks.json").read_text())["grids"]
    except (OSError, KeyError, ValueError):
        pass
    if not args.no_configs:
        for deck, n_x, x_mode in ((args.scene, n_step, A.TALLY_ABSORB | A.TALLY_PATHLENGTH), ("sphere.toml", 50_000_000, A.TALLY_ABSORB),
                                  ("sphere.toml", 50_000_000, A.TALLY_ABSORB | A.TALLY_PATHLENGTH)):
            xcfg = cfg if deck == args.scene else R.Config.load(ROOT / "res" / deck)
            xe = eng if deck == args.scene else R.Engine(1, device_ids=[local_rank])
            if xe is not eng:
                xe.apply(xcfg)
            ms = []
            for k in range(4):   # run 0 = calibration / variant trial of this (scene, mode); best of the rest
                xe.reset_tallies()
                barrier()
                xe.run(n_x, xcfg.iseed, id_offset=step_offset(k, world, rank, n_x), tally_mode=x_mode)
                ms.append(max_over_ranks(xe.last_run_ms))
            xc = xe.fetch(absorb=False, detectors=False)["counters"]
            rate = n_x * world / (min(ms[1:]) * 1e-3)
            xnv = int(np.prod(xcfg.grid[0]))
            entry = {"workload": f"res/{deck}", "tally_mode": x_mode, "mode": "pathlength" if x_mode & A.TALLY_PATHLENGTH else "absorb",
                     "packets_per_step_per_gpu": n_x, "value": rate, "unit": UNIT, "ms_per_step": min(ms[1:]),
                     "kernel_variant": xe.kernel_variant(x_mode), "lost_fraction": xc["lost"] / max(xc["launched"], 1.0)}
            if x_mode & A.TALLY_PATHLENGTH:
                gkey = "200^3 (32 MB, L2 resident)" if xnv * 4 < 126e6 else "500^3 (500 MB)"
                pk = red_peaks.get(gkey, {})
                vox, reds = xc["voxel_crossings"] / xc["launched"], xc["deposit_atomics"] / xc["launched"]
                # algorithmic deposits = one per voxel crossed (the reference's loop, SURVEY 8d W_atom); issued = what the engine
                # sent to L2 (a range update covers a run of voxels with 4).  peak: measured red.global.add.f32 ceiling for this
                # grid size (uniform-random pattern; the hot column of a pencil beam has a lower one, quoted beside it)
                peak = pk.get("uniform_random")
                entry["roofline"] = {"red": {"voxel_crossings_per_packet": vox, "atomics_issued_per_packet": reds,
                                             "algorithmic_red_per_s_per_gpu": rate / world * vox, "issued_red_per_s_per_gpu": rate / world * reds,
                                             "peak": peak, "peak_hot_column": pk.get("hot_column_333"), "unit": "red.f32/s per GPU",
                                             "frac": None if not peak else rate / world * vox / peak,
                                             "frac_issued": None if not peak else rate / world * reds / peak,
                                             "peak_source": f"profiles/r01_red_peaks.json [{gkey}]"},
                                     "segment_mode": xe.segment_mode, "segments_per_packet": xe.segments_per_packet}
            extra.append(entry)
            if xe is not eng:
                xe.close()
        eng.reset_tallies()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    c = res["counters"]
    n_run = c["launched"]
    variant = eng.kernel_variant(mode)
    # ---------------- CPU baseline (rank 0, N=1 only) + algorithmic work from the oracle's counters
    cpu = None
    # algorithmic work of the REFERENCE algorithm per packet: measured from the oracle's counters by the N=1 run below;
    # N>1 runs take what the last committed N=1 run of this scene printed (profiles/bench_ncu.json)
    w_flop_exec = algorithmic_flops_per_packet(scene, len(kind), c, n_run)   # the same op table on the ENGINE's counters
    w_flop = committed_reference_flops(args.scene)
    w_basis = "profiles/bench_ncu.json (what the last N=1 run printed)"
    if world == 1 and not args.no_cpu_baseline:
        n_cpu, secs, cc, threads = cpu_leg(ROOT / "res" / args.scene, args.cpu_seconds)
        cpu = {"value": n_cpu / secs, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} packets of res/{args.scene} on {threads} OpenMP threads (oracle/oracle.cpp, FP64, xoshiro256**)"}
        w_flop = algorithmic_flops_per_packet(scene, len(kind), cc, float(n_cpu))
        w_basis = "oracle counters of this run (reference algorithm, eps=1e-8)"
    if w_flop is None:
        w_flop, w_basis = w_flop_exec, "engine counters (no oracle figure available)"

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak_gpu = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s FP32 FMA at max clock, ONE GPU
    fp32_peak = fp32_peak_gpu * world                    # the job's peak: all its GPUs
    kern_rate = packets / (kernel_ms * 1e-3)             # all ranks
    achieved = kern_rate * w_flop / 1e12
    achieved_exec = kern_rate * w_flop_exec / 1e12
    facts = ncu_facts(args.scene, n_step, variant)
    # secondary ceilings: HBM traffic of the kernel and red.global.add.f32 rate, to show they are NOT the bound
    traffic = facts.get("dram_bytes_per_launch")
    hbm_peak = float(peaks.get("hbm_gbs", 6548.2))
    per_launch_s = kernel_ms / args.steps * 1e-3
    hbm = None if traffic is None else {"achieved": traffic / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s per GPU",
                                        "frac": traffic / per_launch_s / 1e9 / hbm_peak}
    red = None
    rp = red_peaks.get("500^3 (500 MB)" if nv * 4 > 126e6 else "200^3 (32 MB, L2 resident)", {}).get("uniform_random")
    if rp and absorbed_per_packet > 0:
        # absorb mode: one red.f32 per absorbed packet (detector bins are CTA-private shared-memory counters, flushed once)
        red = {"reds_per_packet": absorbed_per_packet, "achieved": kern_rate / world * absorbed_per_packet, "peak": rp, "unit": "red.f32/s per GPU",
               "frac": kern_rate / world * absorbed_per_packet / rp, "peak_source": "profiles/r01_red_peaks.json (tools/red_peak.py, uniform-random voxels)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"res/{args.scene} ({WORKLOADS.get(args.scene, 'shipped input deck')})",
                   "packets_per_step_per_gpu": n_step, "tally_mode": mode, "parallelism": f"packets sharded over {world} GPU(s), NCCL reduce at end",
                   "l2_note": f"no HBM-resident input stream: the scene lives in shared memory; tally grid {nv * 4 / 1e6:.0f} MB "
                              + ("> L2" if nv * 4 > 126e6 else "(L2 resident)"),
                   "wall_ms_timed_region": wall_ms, "nccl_reduce_ms": red_ms, "kernel_variant": variant},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_actual // args.steps),
                "d2h_note": "absorb grid read back as (index, value) pairs of its non-zero voxels after a device-side scan (smcrt_fetch); "
                            f"the dense array would be {nv * 4} bytes",
                "note": "every step: scene/source/detector upload, run" + (", NCCL reduce of all ranks' tallies to rank 0 (sparse pair exchange "
                        "when the grid is < 1/64 full)" if world > 1 else "") + ", read-back into pinned host arrays on rank 0, reset"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                     "achieved_executed": achieved_exec, "frac_executed": achieved_exec / fp32_peak,
                     "issue_slot_utilisation": facts.get("issue_slots_busy"), "lanes_per_instruction": facts.get("lanes_per_instruction"),
                     "peak_per_gpu": fp32_peak_gpu, "traffic": traffic, "hbm": hbm, "red": red,
                     "note": "path is FP32/SFU-issue bound, not HBM or tensor (SURVEY 8d).  frac = packets/s x flops/packet of the REFERENCE "
                             f"algorithm ({w_flop:.0f}: {w_basis}; ~63 plain sphere-tracing sweeps per packet) / peak: the contract figure.  "
                             f"frac_executed = the same op table on the ENGINE's own counters ({w_flop_exec:.0f} flops/packet: its directional "
                             "step bounds need ~6 sweeps): what the FMA pipes actually execute.  issue_slot_utilisation / lanes_per_instruction: "
                             "ncu on this command and kernel variant (profiles/bench_ncu.json; null when that variant has no capture).  "
                             f"peak = {world} GPU(s) x 148 SM x 128 lanes x 2 x sm_max_mhz (nominal FP32 FMA; MEASURED_PEAKS.json has no FP32 entry)"
                             + ("" if scene.n_top < 8 else ".  NB this scene has a culling grid: a sweep evaluates only its cell's candidate list, while "
                                "the flop figures count every top-level SDF per sweep (what the reference evaluates) -- an upper bound on the executed work"),
                     "flops_per_packet": w_flop, "flops_per_packet_executed": w_flop_exec,
                     "engine_sweeps_per_packet": c["sweeps"] / n_run, "engine_nscatt_per_packet": c["nscatt"] / n_run,
                     "engine_lost_fraction": c["lost"] / n_run},
        "cpu_baseline": cpu,
        "configs": extra,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
