"""Write a seeded synthetic vessel tree in the reference's edges.dat / nodes.dat / radii.dat formats
(src/setupGeometry.f90:586-627 reads them from res/; the reference does not ship them: *.dat is git-ignored).

    python tools/make_vessels.py OUT_DIR [n_edges=240] [seed=7]

nodes.dat: "x y z" per line in voxel units (the loader rescales by res=0.001 and recentres: extents 320 x 180 x 260 -> the
0.32 x 0.18 x 0.26 cm box of res/vessels.toml); edges.dat: "i j" 1-based node indices; radii.dat: one radius per node.
"""
import sys
from pathlib import Path

import numpy as np


def make(out_dir, n_edges=240, seed=7):
    rng = np.random.default_rng(seed)
    ext = np.array([320.0, 180.0, 260.0])
    nodes = [ext * [0.5, 0.5, 0.05]]
    radii = [9.0]
    edges = []
    frontier = [(0, np.array([0.0, 0.0, 1.0]))]
    while len(edges) < n_edges:
        i = int(rng.integers(0, len(frontier)))
        parent, d = frontier.pop(i) if len(frontier) > 3 else frontier[i]
        for _ in range(int(rng.integers(1, 3))):
            nd = d + rng.normal(scale=0.55, size=3)
            nd /= np.linalg.norm(nd)
            length = rng.uniform(12.0, 40.0)
            p = np.clip(nodes[parent] + nd * length, 0.04 * ext, 0.96 * ext)
            nodes.append(p)
            radii.append(max(1.5, radii[parent] * rng.uniform(0.72, 0.95)))
            edges.append((parent + 1, len(nodes)))
            frontier.append((len(nodes) - 1, nd))
            if len(edges) >= n_edges:
                break
    # the loader normalises by max|coordinate|; pin the extents with two far-corner nodes of the last edge's family
    nodes[0] = np.maximum(nodes[0], 0)
    nodes.append(ext.copy()); radii.append(1.5)
    nodes.append(np.zeros(3)); radii.append(1.5)
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    np.savetxt(out / "nodes.dat", np.array(nodes), fmt="%.6f")
    np.savetxt(out / "edges.dat", np.array(edges, int), fmt="%d")
    np.savetxt(out / "radii.dat", np.array(radii), fmt="%.6f")
    return len(nodes), len(edges)


if __name__ == "__main__":
    a = sys.argv
    print(make(a[1], int(a[2]) if len(a) > 2 else 240, int(a[3]) if len(a) > 3 else 7))
