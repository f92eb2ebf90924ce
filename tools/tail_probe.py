"""GPU-box diagnostic: per-packet sweep-count distribution of a scene (the kernel tail = the longest single history) and the
run time as a function of the step cap."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R

deck = sys.argv[1] if len(sys.argv) > 1 else "sphere.toml"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = R.Config.load("res/" + deck)
e = R.Engine(1)
e.apply(cfg)
g = e.trace_packets(n, cfg.iseed, tally_mode=mode)
sw, ev, fate = g["sweeps"].astype(np.int64), g["events"], g["fate"]
print(f"{deck} n={n:.0e}: sweeps mean {sw.mean():.2f} median {np.median(sw):.0f} p99 {np.percentile(sw, 99):.0f} p99.99 {np.percentile(sw, 99.99):.0f} max {sw.max()}")
for cut in (100, 1000, 10000, 100000):
    m = sw > cut
    print(f"  > {cut:6d} sweeps: {m.sum():8d} packets ({m.mean():.2e}), {sw[m].sum() / sw.sum():.3%} of all sweeps")
top = np.argsort(sw)[-8:][::-1]
for k in top:
    print(f"  packet {k}: sweeps {sw[k]} events {ev[k]} fate {fate[k]}")
print("lost:", (fate == 3).sum(), "with events", ev[fate == 3][:10])
for cap in (200000, 50000, 10000, 3000):
    e.set_tolerances(max_steps=cap)
    ms = []
    for _ in range(3):
        e.reset_tallies(); e.run(n, cfg.iseed, tally_mode=mode); ms.append(e.last_run_ms)
    c = e.fetch(absorb=False)["counters"]
    print(f"  max_steps {cap:6d}: ms {' '.join('%.1f' % m for m in ms)}  lost {c['lost']}")
