#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference C++ core (oracle/_ref/tnco_core*.so, built by
`make -C oracle ref` from /root/reference).  Run in the build container only; the fixtures are committed
so that machines without /root/reference (the GPU box) can still check the oracle and the CUDA engine
against outputs of the real reference.

Each fixture: a network, an initial tree, a seed and a beta ramp, then the reference's state after
selected sweeps (tree, costs, PRNG state digest, slices) and its final best tree.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import GOLDEN, RefChain, hyper_network, random_tree, ref_core, regular_network  # noqa: E402

CASES = [
    # name, n, seed, n_sweeps, max_width_frac, hyper, dim, every
    ('reg16_inf', 16, 1, 500, None, False, 2, 10),
    ('reg64_inf', 64, 2, 2000, None, False, 2, 10),
    ('reg64_inf_b', 64, 12, 10000, None, False, 2, 10),
    ('hyper64_inf', 64, 3, 1500, None, True, 2, 10),
    ('reg40_d3_inf', 40, 4, 1500, None, False, 3, 10),
    ('reg64_fw50', 64, 5, 1500, 0.5, False, 2, 10),
    ('reg100_fw30', 100, 6, 1500, 0.3, False, 2, 10),
    ('hyper64_fw40', 64, 7, 1000, 0.4, True, 2, 10),
    ('reg48_d3_fw50', 48, 8, 1000, 0.5, False, 3, 10),
    ('reg300_inf', 300, 9, 1000, None, False, 2, 10),
    ('reg200_fw35', 200, 10, 800, 0.35, False, 2, 10),
    ('reg1000_inf', 1000, 11, 300, None, False, 2, 10),
    # dim = 0: per-index dimensions drawn from {2, 4, 8} (the reference's dims-vector code paths)
    ('dims64_inf', 64, 13, 1500, None, False, 0, 10),
    ('dims64_fw45', 64, 14, 1200, 0.45, False, 0, 10),
    ('dimshyper48_fw50', 48, 15, 1000, 0.5, True, 0, 10),
    # sparse-index cost model (SimpleCostModelSparseInds): two more fields, number of sparse indices and n_projs
    ('sparse64_inf', 64, 16, 1500, None, False, 2, 10, 12, 8),
    ('sparse64_d3_inf', 64, 17, 1200, None, False, 3, 10, 10, 5),
    ('sparse100_fw30', 100, 18, 1200, 0.3, False, 2, 10, 12, 16),
    ('sparse64_d3_fw30', 64, 19, 1000, 0.3, False, 3, 10, 10, 2),
    ('sparsedims64_fw40', 64, 20, 1000, 0.4, False, 0, 10, 10, 6),
    ('sparsehyper64_fw40', 64, 21, 1000, 0.4, True, 2, 10, 8, 4),
    # dim = -1: general per-index dimensions drawn from {2, 3, 5, 6, 7} (sequential products / float32 width sums)
    ('gdims64_inf', 64, 22, 1200, None, False, -1, 10),
    ('gdims64_fw40', 64, 23, 1000, 0.4, False, -1, 10),
    ('gdimshyper48_fw50', 48, 24, 800, 0.5, True, -1, 10),
    ('gdimssparse64_fw40', 64, 25, 800, 0.4, False, -1, 10, 10, 12),
    # max_number_new_slices > 0 (finite_width/greedy/optimizer.hpp:226-321): three extra fields (0 sparse, 0, max_new)
    ('reg64_fw50_ns1', 64, 30, 800, 0.5, False, 2, 10, 0, 0, 1),
    ('reg100_fw30_ns4', 100, 31, 800, 0.3, False, 2, 10, 0, 0, 4),
    ('hyper64_fw40_ns2', 64, 32, 700, 0.4, True, 2, 10, 0, 0, 2),
    ('dims64_fw45_ns3', 64, 33, 700, 0.45, False, 0, 10, 0, 0, 3),
    ('gdims64_fw40_ns2', 64, 34, 600, 0.4, False, -1, 10, 0, 0, 2),
    # BASELINE.json configs C2 / C3 / C4 themselves (tnco_b200.networks builds the same networks bench.py runs);
    # n is taken from the network, and for C4 max_width_frac < 0 means the absolute max_width = -frac (32, as benchmarked)
    ('c2_grid6x6d12_inf', 'grid_rqc(6, 6, 12)', 26, 1200, None, False, 2, 10),
    ('c3_sycamore14_inf', 'sycamore(14)', 27, 800, None, False, 2, 10),
    ('c4_sycamore20_fw32', 'sycamore(20)', 28, 800, -32.0, False, 2, 10),
    ('c4_sycamore20_fw32_b', 'sycamore(20)', 29, 400, -32.0, False, 2, 3),
]


def digest(state_str):
    return np.uint32(zlib.crc32(state_str.encode()))


def main():
    assert ref_core() is not None, 'build oracle/_ref first (make -C oracle ref)'
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    for name, n, seed, n_sweeps, frac, hyper, dim, every, *sparse in CASES:
        if only and name not in only:
            continue
        if isinstance(n, str):   # a benchmark network by its tnco_b200.networks constructor
            from tnco_b200 import networks
            ts, ni = eval('networks.' + n)
            n, out = len(ts), []
        elif hyper:
            ts, ni, out = hyper_network(n, seed)
        else:
            ts, ni = regular_network(n, seed)
            out = []
        p, a, b, bits = random_tree(ts, ni, seed + 1, out)
        dims = None
        if dim == 0:
            dims = np.random.default_rng(seed).choice([2, 4, 8], size=ni).astype(np.uint64)
        if dim == -1:
            dims = np.random.default_rng(seed).choice([2, 3, 5, 6, 7], size=ni).astype(np.uint64)
            dim = 0
        mw = None
        if frac is not None and frac < 0:
            mw = -float(frac)
        elif frac is not None:
            if dims is None:
                w0 = max(sum(bin(int(v)).count('1') for v in row) for row in bits)
                mw = float(int(w0 * frac)) * float(np.log2(dim))
            else:
                l2 = np.log2(dims.astype(float))
                w0 = max(sum(l2[i] for i in range(ni) if (int(row[i >> 5]) >> (i & 31)) & 1) for row in bits)
                mw = float(int(w0 * frac))
        sp_inds, sp_bits, n_projs = np.zeros(0, np.int32), None, 0
        max_new = sparse[2] if len(sparse) > 2 else 0
        if sparse and not sparse[0]:
            sparse = []
        if sparse:
            # open indices first (the usual case: sparse output states), the rest drawn at random
            rest = [i for i in np.random.default_rng(seed + 7).permutation(ni).tolist() if i not in out]
            sp_inds = np.array(sorted((list(out) + rest)[:sparse[0]]), np.int32)
            sp_bits = np.zeros((ni + 31) // 32, np.uint32)
            for i in sp_inds.tolist():
                sp_bits[i >> 5] |= np.uint32(1 << (i & 31))
            n_projs = sparse[1]
        rc = RefChain(p, a, b, bits, ni, dim=dim if dims is None else 2, dims=dims, max_width=mw, seed=seed,
                      sparse_bits=sp_bits, n_projs=n_projs, max_number_new_slices=max_new)
        cps = sorted(set([0, 1, 2, 10, 11, n_sweeps // 3, n_sweeps // 2, n_sweeps - 1]))
        rec = dict(parent=[], child0=[], child1=[], log2_total=[], log2_min=[], prng_crc=[], slices=[],
                   min_slices=[])
        init_log2 = rc.log2_total_cost
        init_slices = rc.slices() if mw is not None else np.zeros((ni + 31) // 32, np.uint32)
        for s in range(n_sweeps):
            rc.update(100.0 * s / n_sweeps, update_slices=(s % every == 0))
            if s in cps:
                t = rc.tree()
                rec['parent'].append(t[0]); rec['child0'].append(t[1]); rec['child1'].append(t[2])
                rec['log2_total'].append(rc.log2_total_cost)
                rec['log2_min'].append(rc.log2_min_total_cost)
                rec['prng_crc'].append(digest(rc.prng_state_str()))
                z = np.zeros((ni + 31) // 32, np.uint32)
                rec['slices'].append(rc.slices() if mw is not None else z)
                rec['min_slices'].append(rc.slices(True) if mw is not None else z)
        bt = rc.tree(True)
        max_len = max(len(x) for x in ts)
        ts_arr = np.full((n, max_len), -1, np.int32)
        for i, x in enumerate(ts):
            ts_arr[i, :len(x)] = x
        np.savez_compressed(
            os.path.join(GOLDEN, name + '.npz'), ts_inds=ts_arr, n_inds=ni, output_inds=np.array(out, np.int32),
            parent=p, child0=a, child1=b, bits=bits, dim=dim, dims=np.zeros(0, np.uint64) if dims is None else dims, max_width=np.float64(-1 if mw is None else mw),
            sparse_inds=sp_inds, n_projs=n_projs, max_new=max_new,
            seed=seed, n_sweeps=n_sweeps, beta0=0.0, beta1=100.0, every=every, checkpoints=np.array(cps),
            init_log2_total=init_log2, init_slices=init_slices,
            cp_parent=np.array(rec['parent']), cp_child0=np.array(rec['child0']),
            cp_child1=np.array(rec['child1']), cp_log2_total=np.array(rec['log2_total']),
            cp_log2_min=np.array(rec['log2_min']), cp_prng_crc=np.array(rec['prng_crc']),
            cp_slices=np.array(rec['slices']), cp_min_slices=np.array(rec['min_slices']),
            best_parent=bt[0], best_child0=bt[1], best_child1=bt[2], best_bits=rc.bits(True),
            final_bits=rc.bits())
        print(name, 'log2', init_log2, '->', rc.log2_min_total_cost, 'max_width', mw)


if __name__ == '__main__':
    main()
