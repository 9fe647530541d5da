"""Core-object surface, mirroring the reference's integration tests (tests/test_utils.py:578-748,775-900):
``Optimizer(ctree, cmodel, seed).update(prob)`` in lock step with the CPU oracle (itself pinned to the compiled
reference): equal trees, costs and ``prng_state`` after every update; greedy never increases the cost;
min <= total; is_valid().  Also ContractionTree path <-> tree round trips (tests/test_utils.py:565-572)."""
import ctypes
import math
import os
import random
import subprocess

import numpy as np
import pytest

from helpers import random_tree, regular_network
from oracle import sa_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu_lib():
    subprocess.check_call(['make', '-C', os.path.join(HERE, 'emu')], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    from tnco_b200 import _lib
    return _lib.bind(ctypes.CDLL(os.path.join(HERE, 'emu', 'libtnb_emu.so')))


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    from tnco_b200 import _lib
    monkeypatch.setattr(_lib, '_LIB', request.getfixturevalue('emu_lib') if request.param == 'emu' else None)
    return request.param


def tree_to_linear_path(c0, c1):
    """Pure-Python restatement of ContractionTree.path() (tnco/ctree.py:350-388) for checking."""
    tr = so.get_contraction(c0, c1)
    n = (len(c0) + 1) // 2
    all_pos, path = list(range(n)), []
    for x, y, z in tr.tolist():
        px, py = all_pos.index(x), all_pos.index(y)
        path.append((px, py))
        if px > py:
            px, py = py, px
        all_pos.pop(py)
        all_pos.pop(px)
        all_pos.append(z)
    return path


def test_contraction_tree_from_path_and_back(backend):
    from tnco_b200.ctree import ContractionTree
    for seed in range(5):
        ts, ni = regular_network(14 + 2 * seed, seed)
        names = [[f'i{x}' for x in xs] for xs in ts]
        p, a, b, bits = random_tree(ts, ni, seed)
        path = tree_to_linear_path(a, b)
        ct = ContractionTree(path, names, 2, check_shared_inds=True)
        assert len(ct) == 2 * len(ts) - 1 and ct.n_leaves == len(ts)
        assert ct.path() == [tuple(sorted(x)) if False else x for x in ct.path()]
        # path -> tree -> path is a fixed point, and the contraction it encodes has the same cost
        ct2 = ContractionTree(ct.path(), names, 2)
        assert ct2.path() == ct.path() and ct2 == ct
        P, A, B = ct.arrays()
        nb = np.zeros((len(P), (ni + 31) // 32), np.uint32)
        order = {x: k for k, x in enumerate(ct._inds_order)}
        for z, xs in enumerate(ct.inds):
            for x in xs:
                nb[z, order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
        assert so.tree_cost(A, B, nb, ni)[0] == so.tree_cost(a, b, bits, ni)[0]
        assert ct.max_width() == max(len(xs) for xs in ct.inds)
    with pytest.raises(ValueError):
        ContractionTree([(0, 1)], [['a'], ['b']], 2, check_shared_inds=True)
    assert ContractionTree([(0, 1)], [['i', 'j'], ['j', 'k']], {'i': 2, 'j': 2, 'k': 2}).max_width() == 2.0


def test_partial_path_over_a_component(backend):
    """A path touching a subset of the tensors keeps positions over the whole network (tnco/ctree.py:350-388)."""
    from tnco_b200.ctree import ContractionTree
    ts = [['a', 'b'], ['x'], ['b', 'c'], ['x', 'y'], ['c', 'a']]
    ct = ContractionTree([(0, 2), (2, 3)], ts, 2)
    assert ct._tensors_pos == (0, 2, 4) and ct.n_leaves == 3
    assert ct.path() == [(0, 2), (2, 3)]


@pytest.mark.parametrize('max_width_frac,n_projs', [(None, None), (0.5, None), (None, 6), (0.5, 4)])
def test_optimizer_lockstep_with_oracle(backend, max_width_frac, n_projs):
    from tnco_b200.ctree import ContractionTree
    from tnco_b200.optimize import finite_width, infinite_memory
    from tnco_b200.optimize.finite_width.cost_model import SimpleCostModel as FWModel
    from tnco_b200.optimize.infinite_memory.cost_model import SimpleCostModel
    from tnco_b200.optimize.prob import Greedy, MetropolisHastings
    ts, ni = regular_network(26, 4)
    p, a, b, bits = random_tree(ts, ni, 5)
    ct = ContractionTree(tree_to_linear_path(a, b), ts, 2, check_shared_inds=True)
    P, A, B = ct.arrays()
    order = {x: k for k, x in enumerate(ct._inds_order)}
    nb = np.zeros((len(P), (ni + 31) // 32), np.uint32)
    for z, xs in enumerate(ct.inds):
        for x in xs:
            nb[z, order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    mw = None
    # sparse-index cost model (SimpleCostModel(sparse_inds=, n_projs=)): every fifth index is sparse
    sparse = [x for x in ct._inds_order if order[x] % 5 == 0] if n_projs else None
    sp = None
    if sparse:
        sp = np.zeros((ni + 31) // 32, np.uint32)
        for x in sparse:
            sp[order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    if max_width_frac is not None:
        mw = float(int(max(len(xs) for xs in ct.inds) * max_width_frac))
        opt = finite_width.Optimizer(ct, FWModel(mw, sparse_inds=sparse, n_projs=n_projs), seed=77)
    else:
        opt = infinite_memory.Optimizer(ct, SimpleCostModel(sparse_inds=sparse, n_projs=n_projs), seed=77)
    oc = so.Chain(P, A, B, nb, ni, max_width=mw, seed=77, sparse_bits=sp, n_projs=n_projs)
    assert opt.log2_total_cost == oc.log2_total_cost
    assert opt.prng_state == oc.prng_state_str()
    rnd = random.Random(1)
    for s in range(120):
        greedy = rnd.random() < 0.2
        us = rnd.random() < 0.5
        beta = 100 * rnd.random()
        if mw is None:
            opt.update(Greedy() if greedy else MetropolisHastings(beta))
        else:
            opt.update(Greedy() if greedy else MetropolisHastings(beta), update_slices=us)
        last = oc.total_cost
        oc.update(beta, prob=so.PROB_GREEDY if greedy else so.PROB_MH, update_slices=us)
        if s % 10 == 0 or s > 110:
            for x, y in zip(opt.ctree.arrays(), oc.tree()):
                assert (x == y).all()
            for x, y in zip(opt.min_ctree.arrays(), oc.tree(True)):
                assert (x == y).all()
            assert opt.log2_total_cost == oc.log2_total_cost
            assert opt.log2_min_total_cost == oc.log2_min_total_cost
            assert opt.prng_state == oc.prng_state_str()
            assert opt.min_total_cost <= opt.total_cost
            assert opt.is_valid()
            if mw is not None:
                names = ct._inds_order
                assert opt.slices == frozenset(names[i] for i in range(ni) if (oc.slices()[i >> 5] >> (i & 31)) & 1)
        if greedy and mw is None:
            assert oc.total_cost <= last
    # leaves never change (tests/test_utils.py:701-702)
    cur = opt.ctree
    assert [cur.inds[t] for t in range(26)] == [ct.inds[t] for t in range(26)]


@pytest.mark.parametrize('max_width_frac', [None, 0.5])
def test_pickle_and_resume_from_prng_state(backend, max_width_frac):
    """Core objects pickle through their constructor like the reference's (__reduce__: ctree, cmodel, prng_state
    string, min_ctree [, slices, min_slices]; tnco/optimize/infinite_memory/optimizer.py:234-247,
    finite_width/optimizer.py:330-346) -- and an object rebuilt from the ORACLE's state mid-run (tree, std::mt19937
    state text, best tree, slices) continues in lock step with it."""
    import pickle

    from tnco_b200.ctree import ContractionTree
    from tnco_b200.optimize import finite_width, infinite_memory
    from tnco_b200.optimize.finite_width.cost_model import SimpleCostModel as FWModel
    from tnco_b200.optimize.infinite_memory.cost_model import SimpleCostModel
    from tnco_b200.optimize.prob import MetropolisHastings
    ts, ni = regular_network(26, 4)
    p, a, b, bits = random_tree(ts, ni, 5)
    ct = ContractionTree(tree_to_linear_path(a, b), ts, 2, check_shared_inds=True)
    P, A, B = ct.arrays()
    order = {x: k for k, x in enumerate(ct._inds_order)}
    names = ct._inds_order
    nb = np.zeros((len(P), (ni + 31) // 32), np.uint32)
    for z, xs in enumerate(ct.inds):
        for x in xs:
            nb[z, order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    mw = None if max_width_frac is None else float(int(max(len(xs) for xs in ct.inds) * max_width_frac))
    oc = so.Chain(P, A, B, nb, ni, max_width=mw, seed=31)
    for s in range(60):
        oc.update(0.5 * s)

    def unpack(row):
        return frozenset(names[i] for i in range(ni) if (row[i >> 5] >> (i & 31)) & 1)

    leaves = [ct.inds[t] for t in range(26)]
    cur = ContractionTree.from_arrays(*oc.tree(), leaves, 2)
    best = ContractionTree.from_arrays(*oc.tree(True), leaves, 2)
    if mw is None:
        opt = infinite_memory.Optimizer(cur, SimpleCostModel(), seed=oc.prng_state_str(), _min_ctree=best)
    else:
        opt = finite_width.Optimizer(cur, FWModel(mw), seed=oc.prng_state_str(), _min_ctree=best,
                                     _slices=unpack(oc.slices()), _min_slices=unpack(oc.slices(True)))
    assert opt.prng_state == oc.prng_state_str()
    assert opt.log2_total_cost == oc.log2_total_cost and opt.log2_min_total_cost == oc.log2_min_total_cost
    for s in range(60, 100):
        oc.update(0.5 * s)
        opt.update(MetropolisHastings(0.5 * s))
        if s == 80:  # a pickle round trip in the middle changes nothing
            twin = pickle.loads(pickle.dumps(opt))
            assert twin == opt and twin.prng_state == opt.prng_state
            opt = twin
    for x, y in zip(opt.ctree.arrays(), oc.tree()):
        assert (x == y).all()
    for x, y in zip(opt.min_ctree.arrays(), oc.tree(True)):
        assert (x == y).all()
    assert opt.log2_total_cost == oc.log2_total_cost and opt.log2_min_total_cost == oc.log2_min_total_cost
    assert opt.prng_state == oc.prng_state_str()
    if mw is not None:
        assert opt.slices == unpack(oc.slices()) and opt.min_slices == unpack(oc.slices(True))
    with pytest.raises(ValueError, match='mt19937'):
        infinite_memory.Optimizer(cur, SimpleCostModel(), seed='1 2 3')


def test_skip_slices_lockstep_with_oracle(backend):
    """skip_slices (tnco/optimize/finite_width/optimizer.py:60,96-107): the slicer never takes those indices."""
    from tnco_b200.ctree import ContractionTree
    from tnco_b200.optimize import finite_width
    from tnco_b200.optimize.finite_width.cost_model import SimpleCostModel as FWModel
    from tnco_b200.optimize.prob import MetropolisHastings
    ts, ni = regular_network(30, 14)
    p, a, b, bits = random_tree(ts, ni, 15)
    ct = ContractionTree(tree_to_linear_path(a, b), ts, 2, check_shared_inds=True)
    P, A, B = ct.arrays()
    order = {x: k for k, x in enumerate(ct._inds_order)}
    names = ct._inds_order
    nb = np.zeros((len(P), (ni + 31) // 32), np.uint32)
    for z, xs in enumerate(ct.inds):
        for x in xs:
            nb[z, order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    mw = float(int(max(len(xs) for xs in ct.inds) * 0.5))
    skip = [x for x in names if order[x] % 7 == 0]
    sk = np.zeros((ni + 31) // 32, np.uint32)
    for x in skip:
        sk[order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    opt = finite_width.Optimizer(ct, FWModel(mw), seed=5, skip_slices=skip)
    oc = so.Chain(P, A, B, nb, ni, max_width=mw, seed=5, skip_bits=sk)
    assert opt.skip_slices == frozenset(skip)
    seen = set()
    for s in range(80):
        opt.update(MetropolisHastings(float(s)))
        oc.update(float(s))
        cur = frozenset(names[i] for i in range(ni) if (oc.slices()[i >> 5] >> (i & 31)) & 1)
        assert opt.slices == cur and not (cur & frozenset(skip))
        seen |= cur
    assert seen and opt.log2_total_cost == oc.log2_total_cost and opt.prng_state == oc.prng_state_str()
    assert opt.is_valid()
    with pytest.raises(ValueError, match='subset'):
        finite_width.Optimizer(ct, FWModel(mw), seed=5, skip_slices=['nope'])
    with pytest.raises(ValueError, match='Too many'):
        finite_width.Optimizer(ct, FWModel(1.0), seed=5, skip_slices=list(names))


@pytest.mark.parametrize('max_new', [1, 4])
def test_new_slice_move_lockstep_with_oracle(backend, max_new):
    """max_number_new_slices > 0 (tnco/optimize/finite_width/optimizer.py:59,122; finite_width/greedy/optimizer.hpp:
    226-321): too-wide moves may add random slices.  Same draws, same trees, same slices as the reference's restatement
    (which is in lock step with the compiled reference, tests/test_oracle.py)."""
    from tnco_b200.ctree import ContractionTree
    from tnco_b200.optimize import finite_width, infinite_memory
    from tnco_b200.optimize.finite_width.cost_model import SimpleCostModel as FWModel
    from tnco_b200.optimize.infinite_memory.cost_model import SimpleCostModel
    from tnco_b200.optimize.prob import MetropolisHastings
    ts, ni = regular_network(40, 24)
    p, a, b, bits = random_tree(ts, ni, 25)
    ct = ContractionTree(tree_to_linear_path(a, b), ts, 2, check_shared_inds=True)
    P, A, B = ct.arrays()
    order = {x: k for k, x in enumerate(ct._inds_order)}
    names = ct._inds_order
    nb = np.zeros((len(P), (ni + 31) // 32), np.uint32)
    for z, xs in enumerate(ct.inds):
        for x in xs:
            nb[z, order[x] >> 5] |= np.uint32(1 << (order[x] & 31))
    mw = float(int(max(len(xs) for xs in ct.inds) * 0.45))
    opt = finite_width.Optimizer(ct, FWModel(mw), seed=8, max_number_new_slices=max_new)
    oc = so.Chain(P, A, B, nb, ni, max_width=mw, seed=8, max_number_new_slices=max_new)
    for s in range(150):
        beta = 0.1 * s
        opt.update(MetropolisHastings(beta), update_slices=(s % 5 == 0))
        oc.update(beta, update_slices=(s % 5 == 0))
        if s % 10 == 0 or s > 140:
            cur = frozenset(names[i] for i in range(ni) if (oc.slices()[i >> 5] >> (i & 31)) & 1)
            assert opt.slices == cur, s
            assert opt.log2_total_cost == oc.log2_total_cost and opt.prng_state == oc.prng_state_str(), s
    assert oc.new_slice_counters()[1] > 0
    assert opt.is_valid() and opt.log2_min_total_cost == oc.log2_min_total_cost
    oa, ob = oc.tree()[1:], None
    tp, ta, tb = opt.ctree.arrays()
    assert (ta == oc.tree()[1]).all() and (tb == oc.tree()[2]).all()
    with pytest.raises(TypeError):
        infinite_memory.Optimizer(ct, SimpleCostModel(), seed=1, max_number_new_slices=2)


def test_precision_too_low_and_bad_input(backend):
    from tnco_b200.engine import Engine
    ts, ni = regular_network(60, 1)
    p, a, b, bits = random_tree(ts, ni, 2)
    e = Engine()
    e.set_network(bits[:60], ni, dim=2**60)   # 2^60 per index: costs overflow float64
    e.set_mode()
    e.set_chains(p[None], a[None], b[None], [1])
    with pytest.raises(ValueError, match='Precision is too low'):
        e.costs()
    ts, ni = regular_network(10, 1)
    p, a, b, bits = random_tree(ts, ni, 2)
    e2 = Engine()
    e2.set_network(bits[:10], ni)
    e2.set_mode()
    bad = a.copy()
    bad[-1] = bad[-2]
    with pytest.raises(ValueError):
        e2.set_chains(p[None], bad[None], b[None], [1])
    # hyper-index (an index on three tensors): accepted, runs the HYPER kernels; trees from the host or the device
    lb = bits[:10].copy()
    lb[0, 0] |= 1
    lb[1, 0] |= 1
    lb[2, 0] |= 1
    e3 = Engine()
    e3.set_network(lb, ni).set_mode()
    assert e3.hyper and not e2.hyper
    e3.generate_chains([1, 2])
    assert (e3.costs()[0] > 0).all()
