import sys; sys.path.insert(0,'.')
import numpy as np
import rsmcrt_b200 as R
from oracle import binding as O, scenes
O.build()
name = sys.argv[1] if len(sys.argv) > 1 else 'skin_b200.toml'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000
d = scenes.load('res/'+name)
osc = O.OracleScene(d.scene, d.grid, d.source, d.detectors)
r = osc.run(n, 123, tally_mode=3)
cfg = R.Config.load('res/'+name)
e = R.Engine(1); e.apply(cfg)
e.run(n, 123, tally_mode=3)
g = e.fetch(jmean=True, absorb=True)
jo, jg = r['jmean'].astype('f8'), g['jmean'].astype('f8')
print('path/packet oracle', jo.sum()/n, 'gpu', jg.sum()/n, 'ratio', jg.sum()/jo.sum())
print('absorb oracle', r['absorb'].sum(), 'gpu', g['absorb'].sum(), 'nscatt', r['counters']['nscatt']/n, g['counters']['nscatt']/n, 'bounces', r['counters']['bounces']/n, g['counters']['bounces']/n)
zo, zg = jo.sum(axis=(0,1)), jg.sum(axis=(0,1))
k = len(zo)//20
print('z profile ratio (20 slabs):', np.round(zg.reshape(20,-1).sum(1)/np.maximum(zo.reshape(20,-1).sum(1),1e-30),4))
print('oracle z slabs:', np.round(zo.reshape(20,-1).sum(1)/n,5))
xo, xg = jo.sum(axis=(1,2)), jg.sum(axis=(1,2))
print('x profile ratio:', np.round(xg.reshape(20,-1).sum(1)/np.maximum(xo.reshape(20,-1).sum(1),1e-30),4))
