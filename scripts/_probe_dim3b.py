"""Equal-sweep quality: production re-slicer vs the reference's slicer under the production generator (dev tool)."""
import json, os, sys
import numpy as np
from scipy import stats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnco_b200 import networks
from tnco_b200.engine import Engine, pack_leaf_bits, random_trees, pack_index_set

ts, ni = networks.regular_graph(200, 0)
lb = pack_leaf_bits(ts, ni)
seeds = np.arange(2048, dtype=np.uint64) + 1
p, a, b = random_trees(lb, ni, seeds)
sp = pack_index_set(np.random.default_rng(3).choice(ni, size=40, replace=False).tolist(), ni)
res = {}
for name, dim, mw, sparse in (('d3', 3, 20 * np.log2(3), False), ('d2sparse', 2, 18.0, True)):
    for verb in (0, 1):
        if verb:
            os.environ['TNB_VERBATIM_RESLICER'] = '1'
        else:
            os.environ.pop('TNB_VERBATIM_RESLICER', None)
        e = Engine()
        e.set_network(lb, ni, dim=dim, **(dict(sparse_bits=sp, n_projs=16) if sparse else {})).set_mode(max_width=float(mw), update_slices_every=10)
        e.set_chains(p, a, b, seeds)
        t0, _ = e.costs()
        e.set_betas(np.linspace(0, 100, 3000, endpoint=False))
        e.run(3000)
        t, m = e.costs()
        res[(name, verb)] = np.log2(m)
        print(name, 'verbatim' if verb else 'fast', 'init mean %.3f' % np.log2(t0).mean(), 'best mean %.4f' % np.log2(m).mean(), 'min %.3f' % np.log2(m).min(), flush=True)
        e.close()
    x, y = res[(name, 0)], res[(name, 1)]
    print(name, 'p(fast worse) = %.3g' % stats.mannwhitneyu(x, y, alternative='greater').pvalue, 'diff of means %.4f' % (x.mean() - y.mean()), 'se %.4f' % np.sqrt(x.var() / len(x) + y.var() / len(y)), flush=True)
