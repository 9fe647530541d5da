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
for dim, mw in ((3, 20 * np.log2(3)), (2, 20.0)):
    e = Engine()
    e.set_network(lb, ni, dim=dim).set_mode(max_width=float(mw), update_slices_every=10)
    e.set_chains(p, a, b, seeds)
    e.set_betas(np.linspace(0, 100, 500, endpoint=False))
    e.run(50); e.timing(); c0 = e.counters()
    e.run(500)
    ms, _ = e.timing(); c1 = e.counters()
    t, m = e.costs()
    print(json.dumps(dict(dim=dim, verbatim=bool(os.environ.get('TNB_VERBATIM_RESLICER')), ms=round(ms, 2),
                          proposals_per_s=(c1['proposals'] - c0['proposals']) / (ms * 1e-3),
                          mean_best_log2=float(np.log2(m).mean()))), flush=True)
    e.close()
