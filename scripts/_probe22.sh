python - <<'PY'
import sys, os, time, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import numpy as np
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits
ts, ni = networks.CONFIGS['C4']['make'](); lb = pack_leaf_bits(ts, ni)
seeds = np.arange(4096, dtype=np.uint64)+1
for every in (10, 0, 100, 1):
    e = Engine(); e.set_network(lb, ni).set_mode(max_width=32, update_slices_every=every)
    e.generate_chains(seeds); e.set_betas(np.linspace(0,100,2000,endpoint=False)); e.costs(); e.timing(); e.run(200); e.timing(); c0=e.counters(); e.run(2000); ms,_=e.timing(); c1=e.counters()
    pr = e.progress()
    print('every', every, 'ms %.1f'%ms, 'rate %.3e'%((c1['proposals']-c0['proposals'])/(ms*1e-3)), 'best %.2f mean %.2f'%(np.log2(e.costs()[1]).min(), np.log2(e.costs()[1]).mean()), 'wrej frac %.3f'%(pr['width_rejects'].sum()/pr['proposals'].sum()), 'nslices mean %.1f'%np.mean([bin(int(x)).count('1') for row in e.slices(True) for x in row])*1 if False else '', flush=True)
    e.close()
PY
