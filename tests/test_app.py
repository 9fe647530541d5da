"""App-level surface (mirrors tests/test_app.py:118-297 and tests/test_contraction.py:147-181,315-352 of the
reference): Optimizer factory, optimize() result format, sorting, JSON round trip, disconnected components,
slices, and a symbolic replay of the returned path.  CPU variants drive the kernel-logic emulation
(tests/emu); the `gpu` variants run the real CUDA engine."""
import ctypes
import json
import math
import os
import pickle
import random
import subprocess
from decimal import Decimal

import pytest

from helpers import regular_network

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
    if request.param == 'emu':
        monkeypatch.setattr(_lib, '_LIB', request.getfixturevalue('emu_lib'))
    else:
        monkeypatch.setattr(_lib, '_LIB', None)
    return request.param


def index_rows(ts_inds, n_inds, names=None, extra=()):
    """index-major rows ``(dim, tensor names...)`` as tnco.app.load_tn accepts (tnco/app/app.py:219-236)."""
    rows = [[2] for _ in range(n_inds)]
    for t, xs in enumerate(ts_inds):
        for x in xs:
            rows[x].append(names[t] if names else f't{t}')
    return rows + list(extra)


def replay_cost(path, ts_inds, slices=frozenset(), max_width=None):
    """Contract symbolically along a linear path: sum over steps of 2^|x u y u slices| (test_contraction.py:147-181)."""
    ts = [frozenset(x) for x in ts_inds]
    total = 0
    for x, y in path:
        x, y = sorted((x, y))
        ty = ts.pop(y)
        tx = ts.pop(x)
        total += 2**len(tx | ty | slices)
        new = tx ^ ty
        if max_width is not None:
            assert len(new - slices) <= max_width
        ts.append(new)
    assert len(ts) == 1
    return total


def test_optimizer_factory_and_pickle(backend):
    from tnco_b200.app import Optimizer
    for mw in (None, 12, float('inf')):
        opt = Optimizer(method='sa', seed=3, max_width=mw)
        assert type(opt).__module__.endswith(('finite_width.sa' if mw == 12 else 'infinite_memory.sa'))
        assert pickle.loads(pickle.dumps(opt)) == opt
    with pytest.raises(ModuleNotFoundError):
        Optimizer(method='nope')


def test_argument_errors(backend):
    from tnco_b200.app import Optimizer
    ts, ni = regular_network(10, 1)
    rows = index_rows(ts, ni)
    opt = Optimizer(seed=1)
    with pytest.raises(ValueError):
        opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False)
    with pytest.raises(ValueError):
        opt.optimize(rows, betas=(1, 1), fuse=False, decompose_hyper_inds=False, n_steps=10)
    with pytest.raises(ValueError):
        opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=-1)
    with pytest.raises(TypeError):
        opt.optimize(object(), betas=(0, 100), n_steps=10)


@pytest.mark.parametrize('max_width', [None, 9])
def test_optimize_tn(backend, max_width):
    from tnco_b200.app import Optimizer
    ts, ni = regular_network(30, 5)
    rows = index_rows(ts, ni)
    opt = Optimizer(seed=7, max_width=max_width)
    tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=100, n_runs=4)
    assert len(tn) == 30 and len(res) == 4
    # sorted by cost (test_app.py:229,278-279)
    assert [r.cost for r in res] == sorted(r.cost for r in res)
    name_to_pos = {t.tags['name']: k for k, t in enumerate(tn)}
    assert list(name_to_pos) == [f't{t}' for t in sorted(range(30), key=lambda t: min(ts[t]))] or len(name_to_pos) == 30
    for r in res:
        assert isinstance(r.cost, Decimal)
        assert len(r.path) == 29 and len(r.disconnected_paths) == 1
        assert hasattr(r, 'slices') == (max_width is not None)
        sl = r.slices if max_width is not None else frozenset()
        c = replay_cost(r.path, tn.ts_inds, sl, max_width)
        assert abs(math.log2(c) - math.log2(float(r.cost))) < 1e-4
        # JSON round trip (test_app.py:249-269)
        js = json.loads(r.to_json())
        assert [tuple(x) for x in js['path']] == [tuple(x) for x in r.path]
        assert Decimal(js['cost']) == r.cost
        if max_width is not None:
            assert frozenset(js['slices']) == r.slices
    # same seed => same results (tests/test_determinism.sh)
    tn2, res2 = Optimizer(seed=7, max_width=max_width).optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=100, n_runs=4)
    assert [r.path for r in res] == [r.path for r in res2] and [r.cost for r in res] == [r.cost for r in res2]
    out = Optimizer(seed=7, max_width=max_width, output_format='json').optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=20,
                                                                                 n_runs=2)
    js = json.loads(out)
    assert len(js['res']) == 2 and len(js['tn']['tensors']) == 30


@pytest.mark.parametrize('max_width', [None, 8])
def test_optimize_with_default_load_tn_options(backend, max_width):
    """Default arguments as a user of the reference calls it: load_tn pre-merges tensors (fuse=4, app.py:156) and the
    results refer to the returned, fused network; 'fuse_path' maps it back (app.py:410-414)."""
    import warnings

    from tnco_b200.app import Optimizer, load_tn
    ts, ni = regular_network(40, 6)
    rows = index_rows(ts, ni)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        tn, res = Optimizer(seed=11, max_width=max_width).optimize(rows, betas=(0, 100), n_steps=100, n_runs=3)
        want = load_tn(rows, seed=11)
    assert tn.ts_inds == want.ts_inds and tn.tags['fuse_path'] == want.tags['fuse_path']
    assert len(tn) == 40 - len(tn.tags['fuse_path']) < 40
    assert all(len(t.inds) <= 4 for t in tn)
    for r in res:
        assert len(r.path) == len(tn) - 1
        sl = r.slices if max_width is not None else frozenset()
        c = replay_cost(r.path, tn.ts_inds, sl, max_width)
        assert abs(math.log2(c) - math.log2(float(r.cost))) < 1e-4


def replay_cost_sparse(path, ts_inds, sparse, n_projs, slices=frozenset()):
    """replay_cost under the sparse-index model: 2^|dense| * min(2^|sparse|, n_projs) per step
    (tnco/optimize/infinite_memory/cost_model.py:47-57)."""
    ts = [frozenset(x) for x in ts_inds]
    total = 0
    for x, y in path:
        x, y = sorted((x, y))
        ty = ts.pop(y)
        tx = ts.pop(x)
        u = tx | ty | slices
        total += 2**len(u - sparse) * min(2**len(u & sparse), n_projs)
        ts.append(tx ^ ty)
    return total


@pytest.mark.parametrize('max_width', [None, 9])
def test_optimize_with_sparse_inds(backend, max_width):
    """optimize(..., n_projs=) on a network with sparse indices ('/' rows of load_tn, tnco/app/app.py:219-236):
    costs follow SimpleCostModelSparseInds; argument rules of tnco/optimize/infinite_memory/cost_model.py:79-89."""
    import warnings

    from tnco_b200.app import Optimizer
    ts, ni = regular_network(30, 8)
    rows = index_rows(ts, ni)
    sparse = frozenset(range(0, ni, 6))
    for i in sparse:
        rows[i].append('/')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')  # "fusion ... not yet supported if there are sparse indices"
        tn, res = Optimizer(seed=4, max_width=max_width).optimize(rows, betas=(0, 100), n_steps=120, n_runs=4,
                                                                    n_projs=4)
        assert tn.sparse_inds == sparse and len(tn) == 30
        for r in res:
            sl = r.slices if max_width is not None else frozenset()
            c = replay_cost_sparse(r.path, tn.ts_inds, sparse, 4, sl)
            assert abs(math.log2(c) - math.log2(float(r.cost))) < 1e-4
        # n_projs large enough never caps: the plain cost model's numbers
        tn2, res2 = Optimizer(seed=4).optimize(rows, betas=(0, 100), n_steps=50, n_runs=2, n_projs=2**40)
        for r in res2:
            assert abs(math.log2(replay_cost(r.path, tn2.ts_inds)) - math.log2(float(r.cost))) < 1e-4
        with pytest.raises(ValueError, match='n_projs'):
            Optimizer(seed=4).optimize(rows, betas=(0, 100), n_steps=10)
        with pytest.raises(ValueError, match='n_projs'):
            Optimizer(seed=4).optimize(rows, betas=(0, 100), n_steps=10, n_projs=0)


def test_disconnected_components(backend):
    """n_cc disconnected paths, single-tensor components skipped with cost 0 (sa.py:179-183, test_app.py:272-275)."""
    from tnco_b200.app import Optimizer
    ts1, n1 = regular_network(12, 2)
    ts2, n2 = regular_network(8, 3)
    ts = ts1 + [[x + n1 for x in xs] for xs in ts2] + [[n1 + n2]]
    order = list(range(len(ts)))
    random.Random(0).shuffle(order)
    ts = [ts[i] for i in order]
    rows = index_rows(ts, n1 + n2 + 1)
    tn, res = Optimizer(seed=2).optimize(rows, betas=(0, 50), fuse=False, decompose_hyper_inds=False, n_steps=50, n_runs=3)
    for r in res:
        assert len(r.disconnected_paths) == 3 and len(r.path) == len(ts) - 1
        assert sorted(len(p) for p in r.disconnected_paths) == [0, 7, 11]
        assert r.cost == sum(r.disconnected_costs)
        total = 0
        for p in r.disconnected_paths:  # each path runs independently over all tensors
            work = [frozenset(x) for x in tn.ts_inds]
            for x, y in p:
                x, y = sorted((x, y))
                ty, tx = work.pop(y), work.pop(x)
                assert tx & ty
                total += 2**len(tx | ty)
                work.append(tx ^ ty)
        assert abs(math.log2(total) - math.log2(float(r.cost))) < 1e-4
        work = [frozenset(x) for x in tn.ts_inds]
        for x, y in r.path:
            x, y = sorted((x, y))
            ty, tx = work.pop(y), work.pop(x)
            work.append(tx ^ ty)
        assert len(work) == 1 and len(work[0]) == 1


def test_mt19937_rng_reproduces_reference_runs(backend):
    """rng='mt19937': every run is the reference's run for the same seed and initial tree; compare against the
    CPU oracle (pinned to the reference) driven like tnco/app/infinite_memory/sa.py:199-209."""
    import numpy as np
    from oracle import sa_oracle as so
    from tnco_b200.app import Optimizer
    from tnco_b200.engine import pack_leaf_bits, random_trees
    ts, ni = regular_network(24, 9)
    rows = index_rows(ts, ni)
    opt = Optimizer(seed=11, rng='mt19937', tree_builder='host')  # same initial trees as random_trees below
    tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=200, n_runs=3)
    seeds = random.Random(11).choices(range(2**32), k=3)
    inds = list(dict.fromkeys(x for xs in tn.ts_inds for x in xs))
    pos = {x: k for k, x in enumerate(inds)}
    lb = pack_leaf_bits([[pos[x] for x in xs] for xs in tn.ts_inds], len(inds))
    P, A, B = random_trees(lb, len(inds), np.array(seeds, np.uint64))
    costs = []
    for k, s in enumerate(seeds):
        nb = np.zeros((len(P[k]), lb.shape[1]), np.uint32)
        nb[:len(lb)] = lb
        for z in range(len(lb), len(P[k])):
            nb[z] = nb[A[k][z]] ^ nb[B[k][z]]
        oc = so.Chain(P[k], A[k], B[k], nb, len(inds), seed=s)
        oc.run([0 + n * (100 / 200) for n in range(200)])
        costs.append(Decimal('%.6g' % oc.min_total_cost))
    assert sorted(costs) == [r.cost for r in res]


@pytest.mark.parametrize('max_width', [None, 12])
def test_hyper_index_network_through_the_app(backend, max_width):
    """Networks with hyper-indices and open indices (what the reference's own random test networks look like,
    tnco/testing/utils.py:183-359): the returned path, replayed symbolically with the hyper-count rule of
    tnco/ctree.py:169-189, costs what the result says and respects max_width (test_contraction.py:147-181,315-352)."""
    from helpers import hyper_network
    from tnco_b200.app import Optimizer
    from tnco_b200.tn import get_hyper_count
    ts, ni, out = hyper_network(30, 4)
    rows = index_rows(ts, ni)
    for x in out:
        rows[x].append('*')
    tn, res = Optimizer(seed=5, max_width=max_width).optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=150, n_runs=6)
    assert len(res) == 6 and [r.cost for r in res] == sorted(r.cost for r in res)
    for r in res:
        hc = dict(get_hyper_count(tn.ts_inds))
        for x in tn.output_inds:
            hc[x] += 1
        slices = frozenset(r.slices) if max_width is not None else frozenset()
        work = [frozenset(x) for x in tn.ts_inds]
        total = 0
        for x, y in r.path:
            x, y = sorted((x, y))
            ty, tx = work.pop(y), work.pop(x)
            assert tx & ty
            total += 2**len(tx | ty | slices)
            new = set(tx ^ ty)
            for s in tx & ty:
                hc[s] -= 1
                if hc[s] > 0:
                    new.add(s)
            if max_width is not None:
                assert len(frozenset(new) - slices) <= max_width
            work.append(frozenset(new))
        assert len(work) == 1 and work[0] == frozenset(tn.output_inds)
        assert abs(math.log2(total) - math.log2(float(r.cost))) < 1e-4


@pytest.mark.parametrize('max_width', [None, 14])
def test_per_index_dims_through_the_app(backend, max_width):
    """Indices of different (power-of-two) dimensions: cost = sum over steps of the PRODUCT of the dimensions of
    x | y [| slices], width = sum of log2 dims (cost_model/simple.hpp:46-54, finite_width/cost_model/simple.hpp:47-55)."""
    import numpy as np
    from tnco_b200.app import Optimizer
    ts, ni = regular_network(26, 8)
    dims = np.random.default_rng(1).choice([2, 4, 8], size=ni).tolist()
    rows = index_rows(ts, ni)
    for i, d in enumerate(dims):
        rows[i][0] = d
    tn, res = Optimizer(seed=3, max_width=max_width).optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=150, n_runs=5)
    assert [r.cost for r in res] == sorted(r.cost for r in res)
    for r in res:
        slices = frozenset(r.slices) if max_width is not None else frozenset()
        work = [frozenset(x) for x in tn.ts_inds]
        total = 0
        for x, y in r.path:
            x, y = sorted((x, y))
            ty, tx = work.pop(y), work.pop(x)
            assert tx & ty
            total += math.prod(tn.dims[i] for i in tx | ty | slices)
            new = tx ^ ty
            if max_width is not None:
                assert sum(math.log2(tn.dims[i]) for i in new - slices) <= max_width
            work.append(new)
        assert abs(math.log2(total) - math.log2(float(r.cost))) < 1e-4
    # any positive integer dimensions (sequential products like the reference's dims-vector cost model)
    for i in range(0, ni, 3):
        rows[i][0] = 3
    tn, res = Optimizer(seed=3).optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=60,
                                         n_runs=3)
    for r in res:
        work = [frozenset(x) for x in tn.ts_inds]
        total = 0
        for x, y in r.path:
            x, y = sorted((x, y))
            ty, tx = work.pop(y), work.pop(x)
            total += math.prod(tn.dims[i] for i in tx | ty)
            work.append(tx ^ ty)
        assert abs(math.log2(total) - math.log2(float(r.cost))) < 1e-4


def test_verbose_progress_surface(backend, capsys):
    """verbose >= 2: the reference shows per-run status / log2_total_cost while runs are active
    (tnco/parallel.py:229-317, infinite_memory/sa.py:208-209); here every twentieth of the anneal reports the batch's
    status and best cost to stderr, and the result does not depend on being watched."""
    from tnco_b200.app import Optimizer
    ts, ni = regular_network(30, 5)
    rows = index_rows(ts, ni)
    kw = dict(betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=100, n_runs=4)
    _, quiet = Optimizer(seed=7).optimize(rows, **kw)
    capsys.readouterr()
    _, loud = Optimizer(seed=7, verbose=2).optimize(rows, **kw)
    err = capsys.readouterr().err
    lines = [l for l in err.splitlines() if l.startswith('[tnco_b200 rank 0]')]
    assert len(lines) == 20 and 'sweep 100/100' in lines[-1] and 'best log2 cost' in lines[0]
    assert [r.path for r in quiet] == [r.path for r in loud] and [r.cost for r in quiet] == [r.cost for r in loud]
