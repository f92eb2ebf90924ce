"""Measure the red.global.add.f32 ceilings on this GPU and write profiles/<round>_red_peaks.json."""
import json, sys
sys.path.insert(0, ".")
import rsmcrt_b200 as R
out = {"unit": "red.global.add.f32 per second, one B200, 148x8 CTAs x 256 threads", "grids": {}}
for name, (n, mx) in {"200^3 (32 MB, L2 resident)": ((200, 200, 200), (1.0, 1.0, 1.0)), "500^3 (500 MB)": ((500, 500, 500), (50.0, 50.0, 0.015))}.items():
    e = R.Engine(1)
    e.set_grid(*n, *mx)
    g = {}
    for label, pat, span in (("uniform_random", 0, 1), ("hot_column_333", 1, 333), ("hot_column_8", 1, 8), ("x_runs_64", 2, 64), ("x_runs_8", 2, 8)):
        g[label] = max(e.bench_red(pat, span, 1 << 31) for _ in range(3))
        print(name, label, "%.3e" % g[label], flush=True)
    out["grids"][name] = g
    e.close()
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/red_peaks.json", "w"), indent=1)
