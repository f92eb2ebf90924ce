import sys, numpy as np
sys.path.insert(0, ".")
import rsmcrt_b200 as R
from rsmcrt_b200 import api as A
cfg = R.Config.load("res/sphere.toml")
sc = cfg.scene
e = R.Engine(1); e.apply(cfg)
g = e.trace_packets(100000, cfg.iseed, tally_mode=1)
lost = np.nonzero(g["fate"] == 3)[0]
print("lost", len(lost), "ms", e.last_run_ms, "why", np.bincount(-g["events"][lost], minlength=6) if len(lost) else None)
print("sweeps mean %.1f max %d" % (g["sweeps"].mean(), g["sweeps"].max()), np.sort(g["sweeps"])[-6:])
centers = -sc.xform.reshape(-1, 4, 4)[:, :3, 3]  # Fortran (4,j) -> flat col-major index j*4+3
centers = -sc.xform[:, [3, 7, 11]]
radii = sc.params[:, 0]
for line in open("gpurun_out/dbg.txt"):
    f = line.split()
    pid = int(f[0][3:]); v = [float(x) for x in f[1:]]
    if v[0] != -3: continue
    pos = g["pos"][pid - 0]
    L, NL, surf = int(v[1]), int(v[2]), int(v[10])
    d = np.linalg.norm(pos - centers[:40], axis=1) - radii[:40]
    inside = np.nonzero(d < 0)[0] + 1
    near = np.nonzero(np.abs(d) < 1e-5)[0] + 1
    u = np.array(v[5:8])
    c = centers[surf - 1] if surf <= 40 else None
    b = np.linalg.norm(np.cross(pos - c, u)) / radii[surf - 1] if c is not None else -1
    print(f"id={pid} layer={L} new={NL} surf={surf} r={radii[surf-1] if surf<=40 else 0:.4f} idn={v[3]:.4f} R={v[8]:.4f} steps={int(v[11])} inside={inside} near={near} b/r={b:.6f} (crit {1/1.37:.6f})")
