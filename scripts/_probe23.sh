python - <<'PY'
import sys, os, time, json; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import numpy as np
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits
ts, ni = networks.CONFIGS['C4']['make'](); lb = pack_leaf_bits(ts, ni)
seeds = np.arange(1024, dtype=np.uint64)+1
e = Engine(); e.set_network(lb, ni).set_mode(max_width=32, update_slices_every=10)
e.generate_chains(seeds); e.set_betas(np.linspace(0,100,2000,endpoint=False)); e.costs()
for upto in (0, 100, 500, 1000, 2000):
    e.run(upto)
    nws, nsl = [], []
    S = e.slices()
    for c in range(0, 1024, 128):
        b = e.bits(c)
        k = np.array([sum(bin(int(x)).count('1') for x in row) for row in b])
        nws.append(int((k > 32).sum())); nsl.append(sum(bin(int(x)).count('1') for x in S[c]))
    print('sweeps', upto, 'wide nodes', nws, 'slices', nsl, flush=True)
PY
