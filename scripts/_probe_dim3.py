"""d = 3 network with max_width under the production generator: production re-slicer vs the reference's slicer (dev tool)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits, random_trees

ts, ni = networks.regular_graph(200, 0)
lb = pack_leaf_bits(ts, ni)
seeds = np.arange(4096, dtype=np.uint64) + 1
p, a, b = random_trees(lb, ni, seeds[:256])
p, a, b = (np.tile(x, (16, 1)) for x in (p, a, b))
from tnco_b200.engine import pack_index_set
sp = pack_index_set(np.random.default_rng(3).choice(ni, size=40, replace=False).tolist(), ni)
for dim, mw, sparse in ((3, 20 * np.log2(3), False), (2, 20.0, False), (2, 18.0, True), (3, 18 * np.log2(3), True)):
    e = Engine()
    e.set_network(lb, ni, dim=dim, **(dict(sparse_bits=sp, n_projs=16) if sparse else {})).set_mode(max_width=float(mw), update_slices_every=10)
    e.set_chains(p, a, b, seeds)
    e.set_betas(np.linspace(0, 100, 500, endpoint=False))
    e.run(50); e.timing(); c0 = e.counters()
    e.run(500)
    ms, _ = e.timing(); c1 = e.counters()
    t, m = e.costs()
    print(json.dumps(dict(dim=dim, sparse=sparse, verbatim=bool(os.environ.get('TNB_VERBATIM_RESLICER')), ms=round(ms, 2),
                          proposals_per_s=(c1['proposals'] - c0['proposals']) / (ms * 1e-3),
                          mean_best_log2=float(np.log2(m).mean()))), flush=True)
    e.close()
