"""bench.py's reference arm runs on the CPU (the oracle on the host cores): its output contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_exactly_one_json_line():
    # as under torch.distributed.run, which exports OMP_NUM_THREADS=1 to every rank: the arm must still use the host's cores
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.5"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]          # everything else (library banners, progress) goes to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photon_packets_per_s" and d["unit"] == "packets/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "validation1.toml" in d["config"]["workload"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["product_library_loaded"] is False      # scene built by oracle/scenes.py, not by libsmcrt_gpu.so
