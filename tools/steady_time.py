"""Steady-state packets/s per scene: the first run picks the register budget, the later ones are the number."""
import sys; sys.path.insert(0, ".")
import rsmcrt_b200 as R
cases = (("validation1.toml", 100000000, 1), ("scat_test.toml", 20000000, 1), ("sphere.toml", 10000000, 3), ("validation2.toml", 10000000, 1),
         ("validation1.toml", 20000000, 3), ("vessels.toml", 10000000, 1))
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]))]
for deck, n, mode in cases:
    try:
        cfg = R.Config.load("res/" + deck)
    except Exception as ex:
        print(deck, "skipped:", ex); continue
    e = R.Engine(1); e.apply(cfg)
    out = []
    for it in range(3):
        e.reset_tallies(); e.run(n, cfg.iseed, tally_mode=mode); out.append(e.last_run_ms)
    c = e.fetch(absorb=False)["counters"]
    print("%-18s mode=%d n=%.0e  ms %s  packets/s %.3e  sweeps/pkt %.1f lost %d" % (deck, mode, n, " ".join("%.1f" % m for m in out), n / min(out[1:]) * 1e3, c["sweeps"] / n, c["lost"]), flush=True)
    e.close()
