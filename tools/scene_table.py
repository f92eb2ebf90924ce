"""Steady-state packets/s of every BASELINE scene (C1..C5 + the validation decks), default tally mode (what
`fpm @runmp` builds: absorb only) and -Dpathlength mode.  One JSON line per (scene, mode) on stdout.

    python tools/scene_table.py [out.jsonl] [scale=1.0]

The first run of a scene is the engine's own variant trial; the number is the best of the two runs after it.
"""
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import rsmcrt_b200 as R  # noqa: E402
import make_vessels  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
CASES = (  # deck, packets, tally modes
    ("sphere.toml", 50_000_000, (1, 3)),
    ("validation1.toml", 100_000_000, (1, 3)),
    ("jacques.toml", 50_000_000, (1,)),
    ("skin_b200.toml", 20_000_000, (1, 3)),
    ("validation2.toml", 10_000_000, (1,)),
    ("lens.toml", 50_000_000, (1, 3)),
    ("test_dects.toml", 20_000_000, (1,)),
    ("scat_test.toml", 20_000_000, (1,)),
    ("vessels.toml", 20_000_000, (1, 3)),
)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    tmp = Path(tempfile.mkdtemp())
    make_vessels.make(tmp, 240, 7)
    lines = []
    for deck, n, modes in CASES:
        if only and deck not in only:
            continue
        n = max(int(n * scale), 100_000)
        try:
            cfg = R.Config.load(ROOT / "res" / deck, res_dir=tmp) if deck == "vessels.toml" else R.Config.load(ROOT / "res" / deck)
        except Exception as ex:  # a deck the dispatcher refuses
            print(json.dumps({"deck": deck, "skipped": str(ex)}), flush=True)
            continue
        for mode in modes:
            e = R.Engine(1)
            e.apply(cfg)
            ms = []
            for _ in range(3):
                e.reset_tallies()
                e.run(n, cfg.iseed, tally_mode=mode)
                ms.append(e.last_run_ms)
            c = e.fetch(absorb=False)["counters"]
            rec = {"deck": deck, "tally_mode": mode, "packets": n, "ms": [round(m, 2) for m in ms],
                   "packets_per_s": n / min(ms[1:]) * 1e3, "sweeps_per_packet": c["sweeps"] / n, "nscatt_per_packet": c["nscatt"] / n,
                   "bounces_per_packet": c["bounces"] / n, "lost": int(c["lost"]), "kernel_variant": e.kernel_variant(mode)}
            print(json.dumps(rec), flush=True)
            lines.append(rec)
            e.close()
    if out_path:
        Path(out_path).write_text("\n".join(json.dumps(r) for r in lines) + "\n")


if __name__ == "__main__":
    main()
