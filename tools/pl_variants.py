import sys, os; sys.path.insert(0,'.')
import numpy as np
import rsmcrt_b200 as R
name, n = sys.argv[1], int(float(sys.argv[2]))
cfg = R.Config.load('res/'+name)
e = R.Engine(1); e.apply(cfg)
e.run(n, 5, tally_mode=3)
j = e.fetch(jmean=True, absorb=False)
print(name, 'variant', os.environ.get('SMCRT_VARIANT_FORCE'), 'ms', round(e.last_run_ms,2), 'path/packet', j['jmean'].sum(dtype='f8')/n, 'lost', j['counters']['lost'], flush=True)
