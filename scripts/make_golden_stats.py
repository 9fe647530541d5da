#!/usr/bin/env python3
"""Reference best-cost distributions at equal sweep counts, from the UNMODIFIED reference core (oracle/_ref), for
the statistical parity test of the production kernels (tests/test_gpu_statistics.py).

For every case: the benchmark network (tnco_b200.networks), R independent reference runs of `n_sweeps` sweeps with
betas 0 -> 100 driven like `core_` (tnco/app/infinite_memory/sa.py:199-209, finite_width/sa.py:221-231), initial
trees from the same generator the GPU arm uses (tnb_random_trees, seeds 1..R).  Output: tests/golden/stat_<case>.json
with the R values of log2_min_total_cost.  Run in the build container only (needs oracle/_ref)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CASES = {
    # name: (network constructor, max_width, n_sweeps, runs, update_slices)
    'c2_1e4': ('grid_rqc(6, 6, 12)', None, 10000, 256, 10),
    'c3_1e4': ('sycamore(14)', None, 10000, 256, 10),
    'c4_1e4': ('sycamore(20)', 32.0, 10000, 256, 10),
    'c4_3e4': ('sycamore(20)', 32.0, 30000, 256, 10),
    'c1_1e4': ('regular_graph(64, 0)', None, 10000, 256, 10),
}


def one(args):
    net, mw, n_sweeps, every, seed, P, A, B, nb, ni = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import RefChain
    rc = RefChain(P, A, B, nb, ni, seed=seed, max_width=mw)
    for s in range(n_sweeps):
        rc.update(100.0 * s / n_sweeps, update_slices=(s % every == 0))
    return rc.log2_min_total_cost


def main():
    from joblib import Parallel, delayed
    from helpers import GOLDEN, ref_core
    from tnco_b200 import networks
    from tnco_b200.engine import pack_leaf_bits, random_trees
    assert ref_core() is not None, 'build oracle/_ref first (make -C oracle ref)'
    only = set(sys.argv[1:])
    for name, (net, mw, n_sweeps, runs, every) in CASES.items():
        if only and name not in only:
            continue
        ts, ni = eval('networks.' + net)
        lb = pack_leaf_bits(ts, ni)
        n = lb.shape[0]
        seeds = np.arange(runs, dtype=np.uint64) + 1
        P, A, B = random_trees(lb, ni, seeds)
        jobs = []
        for k in range(runs):
            nb = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
            nb[:n] = lb
            for z in range(n, 2 * n - 1):
                nb[z] = nb[A[k][z]] ^ nb[B[k][z]]
            jobs.append((net, mw, n_sweeps, every, int(seeds[k]), P[k], A[k], B[k], nb, ni))
        vals = Parallel(n_jobs=-1)(delayed(one)(j) for j in jobs)
        out = dict(network=net, max_width=mw, n_sweeps=n_sweeps, update_slices=every, betas=[0, 100],
                   seeds=[1, runs], trees='tnb_random_trees(seeds), TNB_TREES_GREEDY',
                   source='oracle/_ref (unmodified reference core), scripts/make_golden_stats.py',
                   log2_min_total_cost=[float(v) for v in vals])
        with open(os.path.join(GOLDEN, f'stat_{name}.json'), 'w') as f:
            json.dump(out, f)
        v = np.array(vals)
        print(name, 'mean', v.mean(), 'min', v.min(), 'std', v.std())


if __name__ == '__main__':
    main()
