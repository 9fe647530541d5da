"""Shared test helpers: tiny pure-Python network / tree builders (independent of the product code)
and the driver for the compiled reference core (oracle/_ref), when it is present."""
from __future__ import annotations

import glob
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


# ----------------------------------------------------------------------------- reference core
def ref_core():
    """The unmodified reference C++ core (tnco_core pybind module) or None."""
    d = os.path.join(ROOT, 'oracle', '_ref')
    if not glob.glob(os.path.join(d, 'tnco_core*.so')):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import tnco_core  # noqa
        return tnco_core
    except Exception:  # wrong python ABI on this box, ...
        return None


# ----------------------------------------------------------------------------- networks
def regular_network(n, seed, degree=3):
    """Random `degree`-regular graph: one tensor per vertex, one index per edge."""
    import networkx as nx
    g = nx.random_regular_graph(degree, n, seed=seed)
    while not nx.is_connected(g):
        seed += 1000003
        g = nx.random_regular_graph(degree, n, seed=seed)
    ts = [[] for _ in range(n)]
    for k, (a, b) in enumerate(sorted(tuple(sorted(e)) for e in g.edges())):
        ts[a].append(k)
        ts[b].append(k)
    return ts, g.number_of_edges()


def hyper_network(n, seed, n_hyper=6, n_open=3):
    """Connected random network with some hyper-indices (index on 3+ tensors) and open indices."""
    rng = random.Random(seed)
    ts, k = regular_network(n, seed)
    for _ in range(n_hyper):
        for t in rng.sample(range(n), rng.choice([3, 4])):
            ts[t].append(k)
        k += 1
    out = []
    for _ in range(n_open):
        ts[rng.randrange(n)].append(k)
        out.append(k)
        k += 1
    return ts, k, out


def leaf_bits(ts_inds, n_inds):
    W = (n_inds + 31) // 32
    b = np.zeros((len(ts_inds), W), np.uint32)
    for t, xs in enumerate(ts_inds):
        for x in xs:
            b[t, x >> 5] |= np.uint32(1 << (x & 31))
    return b


def random_tree(ts_inds, n_inds, seed, output_inds=()):
    """Random contraction tree where every contracted pair shares an index (random edge order +
    union-find).  Returns parent, child0, child1 (reference numbering) and all node bitsets,
    using the hyper-count rule of tnco/ctree.py:169-189."""
    rng = random.Random(seed)
    n = len(ts_inds)
    N = 2 * n - 1
    W = (n_inds + 31) // 32
    holders = {}
    for t, xs in enumerate(ts_inds):
        for x in xs:
            holders.setdefault(x, []).append(t)
    count = {x: len(h) - 1 for x, h in holders.items()}
    for x in output_inds:
        count[x] += 1
    sets = [set(xs) for xs in ts_inds] + [None] * (n - 1)
    rep = list(range(N))  # union-find over leaves -> current cluster node

    def find(a):
        while rep[a] != a:
            rep[a] = rep[rep[a]]
            a = rep[a]
        return a

    parent = np.full(N, -1, np.int32)
    c0 = np.full(N, -1, np.int32)
    c1 = np.full(N, -1, np.int32)
    nxt = n
    edges = [x for x, h in holders.items() if len(h) >= 2]
    while nxt < N:
        rng.shuffle(edges)
        progressed = False
        for x in edges:
            hs = sorted({find(t) for t in holders[x]})
            if len(hs) < 2:
                continue
            a, b = rng.sample(hs, 2)
            z = nxt
            nxt += 1
            shared = sets[a] & sets[b]
            iz = sets[a] ^ sets[b]
            for s in shared:
                count[s] -= 1
                if count[s] > 0:
                    iz.add(s)
            sets[z] = iz
            c0[z], c1[z] = a, b
            parent[a] = parent[b] = z
            rep[a] = rep[b] = z
            progressed = True
            if nxt == N:
                break
        if not progressed:
            raise ValueError('network is not connected')
    bits = np.zeros((N, W), np.uint32)
    for t in range(N):
        for x in sets[t]:
            bits[t, x >> 5] |= np.uint32(1 << (x & 31))
    return parent, c0, c1, bits


def positions(row):
    return [w * 32 + b for w, v in enumerate(row.tolist()) for b in range(32) if (v >> b) & 1]


# ----------------------------------------------------------------------------- reference driver
class RefChain:
    """Drives the compiled reference exactly as examples/BaseOptimization.ipynb does (raw tnco_core)."""

    def __init__(self, parent, c0, c1, bits, n_inds, *, dim=2, dims=None, max_width=None, seed=0,
                 disable_shared_inds=False, sparse_bits=None, n_projs=None, skip_bits=None, max_number_new_slices=0):
        tc = ref_core()
        self.tc = tc
        nodes = [
            tc.Node((int(c0[i]), int(c1[i])), int(parent[i])) for i in range(len(parent))
        ]
        inds = [tc.Bitset(positions(bits[i]), n_inds) for i in range(len(parent))]
        d = int(dim) if dims is None else [int(x) for x in dims]
        ctree = tc.ContractionTree(nodes, inds, d, check_shared_inds=not disable_shared_inds)
        self.n_inds = n_inds
        self.W = (n_inds + 31) // 32
        self.finite = max_width is not None
        sp = None if sparse_bits is None else tc.Bitset(positions(sparse_bits), n_inds)
        if self.finite and sp is not None:
            cm = tc.optimize.finite_width.cost_model.SimpleCostModelSparseInds_float64_float32(
                float(max_width), sp, int(n_projs))
            self.opt = tc.optimize.finite_width.greedy.Optimizer_float64_float32(
                ctree, cm, seed=int(seed), disable_shared_inds=disable_shared_inds,
                max_number_new_slices=int(max_number_new_slices),
                skip_slices=None if skip_bits is None else tc.Bitset(positions(skip_bits), n_inds))
        elif sp is not None:
            cm = tc.optimize.infinite_memory.cost_model.SimpleCostModelSparseInds_float64(sp, int(n_projs))
            self.opt = tc.optimize.infinite_memory.Optimizer_float64(
                ctree, cm, seed=int(seed), disable_shared_inds=disable_shared_inds)
        elif self.finite:
            cm = tc.optimize.finite_width.cost_model.SimpleCostModel_float64_float32(float(max_width))
            self.opt = tc.optimize.finite_width.greedy.Optimizer_float64_float32(
                ctree, cm, seed=int(seed), disable_shared_inds=disable_shared_inds,
                max_number_new_slices=int(max_number_new_slices),
                skip_slices=None if skip_bits is None else tc.Bitset(positions(skip_bits), n_inds))
        else:
            cm = tc.optimize.infinite_memory.cost_model.SimpleCostModel_float64()
            self.opt = tc.optimize.infinite_memory.Optimizer_float64(
                ctree, cm, seed=int(seed), disable_shared_inds=disable_shared_inds)
        self.mh = tc.optimize.prob.MetropolisHastings_float64()
        self.greedy = tc.optimize.prob.Greedy_float64()

    def update(self, beta, update_slices=True, greedy=False):
        p = self.greedy if greedy else self.mh
        if not greedy:
            p.beta = float(beta)
        if self.finite:
            self.opt.update(p, update_slices=bool(update_slices))
        else:
            self.opt.update(p)

    def _tree(self, ct):
        N = len(ct.nodes)
        p, a, b = (np.full(N, -1, np.int32) for _ in range(3))
        for i, nd in enumerate(ct.nodes):
            p[i] = -1 if nd.parent is None else nd.parent
            ch = nd.children
            a[i] = -1 if ch[0] is None else ch[0]
            b[i] = -1 if ch[1] is None else ch[1]
        return p, a, b

    def _bits(self, ct):
        N = len(ct.nodes)
        o = np.zeros((N, self.W), np.uint32)
        for i, bs in enumerate(ct.inds):
            for x in bs.positions():
                o[i, x >> 5] |= np.uint32(1 << (x & 31))
        return o

    def tree(self, best=False):
        return self._tree(self.opt.min_ctree if best else self.opt.ctree)

    def bits(self, best=False):
        return self._bits(self.opt.min_ctree if best else self.opt.ctree)

    def slices(self, best=False):
        o = np.zeros(self.W, np.uint32)
        for x in (self.opt.min_slices if best else self.opt.slices).positions():
            o[x >> 5] |= np.uint32(1 << (x & 31))
        return o

    log2_total_cost = property(lambda s: s.opt.log2_total_cost)
    log2_min_total_cost = property(lambda s: s.opt.log2_min_total_cost)

    def prng_state_str(self):
        return self.opt.prng_state.strip()


def golden_sparse(g):
    """(sparse_bits [W32] | None, n_projs | None) of a golden fixture (sparse-index cost model cases)."""
    if 'sparse_inds' not in g.files or int(g['n_projs']) == 0:
        return None, None
    b = np.zeros((int(g['n_inds']) + 31) // 32, np.uint32)
    for i in g['sparse_inds'].tolist():
        b[i >> 5] |= np.uint32(1 << (i & 31))
    return b, int(g['n_projs'])


def py_merge_paths(n_tensors, paths):
    """Checker for merge_contraction_paths: replay every component path on its own position list while keeping one
    shared list of what is left in the merged network; a contraction's merged positions are where its two operands
    sit in the shared list.  Trailing (0, 1) steps join the components."""
    shared = [('t', k) for k in range(n_tensors)]
    out = []
    for c, path in enumerate(paths):
        own = [('t', k) for k in range(n_tensors)]
        for step, (x, y) in enumerate(path):
            lo, hi = min(x, y), max(x, y)
            b = own.pop(hi)
            a = own.pop(lo)
            new = ('c', c, step)
            own.append(new)
            ia, ib = shared.index(a), shared.index(b)
            out.append((min(ia, ib), max(ia, ib)))
            for k in sorted((ia, ib), reverse=True):
                shared.pop(k)
            shared.append(new)
    return out + [(0, 1)] * (len(shared) - 1)
