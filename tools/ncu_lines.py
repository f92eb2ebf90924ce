"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": fn = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0] not in ("", "Line No") and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
        except (KeyError, ValueError):
            pass
ti = sum(a[3] for a in agg); tt = sum(a[4] for a in agg); ts = sum(a[5] for a in agg)
print("total warp inst %.3e thread inst %.3e samples %d  avg lanes %.2f" % (ti, tt, ts, tt / max(ti, 1)))
agg.sort(key=lambda a: -a[3])
for a in agg[:top]:
    print("%5.2f%% inst %5.2f%% smp lanes %5.1f  %s:%d  %s" % (100 * a[3] / ti, 100 * a[5] / max(ts, 1), a[4] / max(a[3], 1), a[0], a[1], a[2]))
if len(sys.argv) > 3:
    w = int(sys.argv[3]); b = collections.defaultdict(lambda: [0, 0, 0])
    for a in agg:
        k = (a[0], a[1] // w * w); b[k][0] += a[3]; b[k][1] += a[4]; b[k][2] += a[5]
    print("--- buckets of %d lines" % w)
    for k in sorted(b):
        v = b[k]
        if v[0] / ti > 0.004: print("%-28s %5d  %5.2f%% inst %5.2f%% smp lanes %5.1f" % (k[0], k[1], 100 * v[0] / ti, 100 * v[2] / max(ts, 1), v[1] / max(v[0], 1)))
