"""Per-packet fates, engine vs oracle on the same Philox streams: where do the histories part?  (GPU box)
    python tools/fate_vs_oracle.py deck packets"""
import sys; sys.path.insert(0, '.')
import numpy as np
import rsmcrt_b200 as R
from oracle import binding as O, scenes
O.build()
name, n = sys.argv[1], int(float(sys.argv[2]))
d = scenes.load('res/' + name)
osc = O.OracleScene(d.scene, d.grid, d.source, d.detectors)
o = osc.run(n, 123, per_packet=True, grids=False)
e = R.Engine(1); e.apply(R.Config.load('res/' + name))
g = e.trace_packets(n, 123)
same = (g['fate'] == o['fate']) & (g['nscatt'] == o['nscatt'])
print('identical', same.mean(), 'fates gpu', np.bincount(g['fate'], minlength=4), 'oracle', np.bincount(o['fate'], minlength=4))
diff = np.nonzero(~same)[0]
early = diff[(g['nscatt'][diff] < o['nscatt'][diff])]
print('gpu ended earlier:', len(early), 'later:', len(diff) - len(early))
gm = np.array(d.grid[1])
for i in early[:25]:
    print(i, 'gpu fate', g['fate'][i], 'nsc', g['nscatt'][i], 'pos/gmax', np.round(g['pos'][i] / gm, 7), '| oracle fate', o['fate'][i], 'nsc', o['nscatt'][i], 'pos/gmax', np.round(o['pos'][i] / gm, 5))
pe = g['pos'][early] / gm
print('where the early GPU endings sit (|coord|/gmax > 0.999999): +x', (pe[:, 0] > 0.999999).sum(), '-x', (pe[:, 0] < -0.999999).sum(), '+y', (pe[:, 1] > 0.999999).sum(),
      '-y', (pe[:, 1] < -0.999999).sum(), '+z', (pe[:, 2] > 0.999999).sum(), '-z', (pe[:, 2] < -0.999999).sum())
