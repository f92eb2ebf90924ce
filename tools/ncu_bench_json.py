"""profiles/bench_ncu.json from an ncu --set full capture of bench.py's dominant kernel (no GPU needed):

    python tools/ncu_bench_json.py X.ncu-rep scene packets_per_step variant [reference_flops_per_packet]

Adds / replaces the entry "scene:packets:variant" with the DRAM bytes of one launch, the issue-slot utilisation and the lanes per
instruction that bench.py quotes beside its roofline fractions, and optionally the reference algorithm's flops/packet that the
matching N=1 bench run printed (roofline.flops_per_packet), which N>1 runs reuse."""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, scene, packets, variant = sys.argv[1], sys.argv[2], int(float(sys.argv[3])), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, r = rows[0], rows[2]
val = lambda k: float(r[hdr.index(k)].replace(",", ""))
unit = lambda k: rows[1][hdr.index(k)]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = val("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + val("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
out = Path(__file__).resolve().parent.parent / "profiles" / "bench_ncu.json"
table = json.loads(out.read_text()) if out.exists() else {}
table[f"{scene}:{packets}:{variant}"] = {
    "kernel": r[hdr.index("Kernel Name")], "report": Path(rep).name, "duration_ms": val("gpu__time_duration.sum") * (1e-3 if unit("gpu__time_duration.sum") == "us" else 1.0),
    "dram_bytes_per_launch": dram, "issue_slots_busy": val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"), "warp_instructions": val("smsp__inst_executed.sum"),
    "local_loads": val("sass__inst_executed_local_loads"), "local_stores": val("sass__inst_executed_local_stores"),
    "registers": val("launch__registers_per_thread"),
}
if len(sys.argv) > 5:
    table.setdefault("reference_flops_per_packet", {})[scene] = float(sys.argv[5])
out.write_text(json.dumps(table, indent=1) + "\n")
print(json.dumps(table[f"{scene}:{packets}:{variant}"], indent=1))
