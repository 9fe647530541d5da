"""GPU parity (run on the B200 box: pytest -m gpu): the CUDA engine, through the C-ABI, against
 (i) golden vectors produced by the unmodified reference C++ core (tests/golden), in TNB_RNG_MT19937 mode
     (same seeds => bit-identical trees, costs, slices as the reference, checkpoint by checkpoint);
 (ii) the CPU oracle on a replayed, oracle-recorded draw stream (TNB_RNG_REPLAY);
 (iii) the CPU oracle's full-tree cost on random trees (tnb_eval_cost), tolerance 1e-12 relative
      (actual agreement is bit-exact for dim=2).
Every tile shape the engine can pick is exercised via TNB_TILE."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_sparse, random_tree, regular_network
from oracle import sa_oracle as so

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


def _engine(g, rng, tile=None, every=None):
    from tnco_b200.engine import Engine, pack_index_set
    if tile:
        os.environ['TNB_TILE'] = str(tile)
    else:
        os.environ.pop('TNB_TILE', None)
    mw = float(g['max_width'])
    mw = None if mw < 0 else mw
    n = (g['parent'].shape[0] + 1) // 2
    e = Engine()
    dims = g['dims'] if 'dims' in g.files and len(g['dims']) else None   # per-index dims fixtures (dim == 0)
    sp, n_projs = golden_sparse(g)
    e.set_network(g['bits'][:n], int(g['n_inds']), dim=int(g['dim']) or 2, dims=dims,
                  output_bits=pack_index_set(g['output_inds'].tolist(), int(g['n_inds'])), sparse_bits=sp,
                  n_projs=n_projs)
    assert e.hyper == bool(len(g['output_inds']))
    e.set_mode(max_width=mw, update_slices_every=int(g['every']) if every is None else every, rng=rng)
    if 'max_new' in g.files and int(g['max_new']):
        e.set_new_slices(int(g['max_new']))
    e.set_chains(g['parent'][None], g['child0'][None], g['child1'][None], [int(g['seed'])])
    os.environ.pop('TNB_TILE', None)
    return e, mw


def _check_against_golden(g, e, mw):
    n_sweeps = int(g['n_sweeps'])
    e.set_betas([100.0 * s / n_sweeps for s in range(n_sweeps)])
    t, m = e.costs()
    assert (e.bits(0) == g['bits']).all()   # index sets derived on the device (hyper-count rule) == the input tree's
    assert np.log2(t[0]) == float(g['init_log2_total'])
    if mw is not None:
        assert (e.slices()[0] == g['init_slices']).all()
    for k, s in enumerate(g['checkpoints'].tolist()):
        e.run(s + 1)
        p, a, b = e.trees()
        assert (p[0] == g['cp_parent'][k]).all() and (a[0] == g['cp_child0'][k]).all() and (
            b[0] == g['cp_child1'][k]).all(), s
        t, m = e.costs()
        assert np.log2(t[0]) == g['cp_log2_total'][k], s
        assert np.log2(m[0]) == g['cp_log2_min'][k], s
        if mw is not None:
            assert (e.slices()[0] == g['cp_slices'][k]).all(), s
            assert (e.slices(True)[0] == g['cp_min_slices'][k]).all(), s
    e.run(n_sweeps)
    p, a, b = e.trees(True)
    assert (p[0] == g['best_parent']).all() and (a[0] == g['best_child0']).all() and (
        b[0] == g['best_child1']).all()
    assert (e.bits(0) == g['final_bits']).all()


@pytest.mark.parametrize('name', CASES)
def test_mt19937_mode_matches_reference_golden(name):
    from tnco_b200.engine import RNG_MT19937
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    e, mw = _engine(g, RNG_MT19937)
    _check_against_golden(g, e, mw)


@pytest.mark.parametrize('tile', [8, 16, 32])
@pytest.mark.parametrize('name', ['reg64_inf', 'reg64_fw50', 'reg40_d3_inf'])
def test_every_tile_shape_matches_golden(name, tile):
    from tnco_b200.engine import RNG_MT19937
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    e, mw = _engine(g, RNG_MT19937, tile=tile)
    assert e.config()['tile'] == tile
    _check_against_golden(g, e, mw)


@pytest.mark.parametrize('name', ['reg64_inf', 'reg100_fw30', 'reg300_inf', 'hyper64_inf', 'hyper64_fw40',
                                  'dims64_fw45', 'dimshyper48_fw50', 'sparse64_inf', 'sparse100_fw30',
                                  'sparsehyper64_fw40', 'reg100_fw30_ns4', 'hyper64_fw40_ns2', 'gdims64_fw40_ns2'])
def test_replay_of_recorded_draw_stream_is_bit_exact(name):
    """north_star: replaying a reference-recorded proposal / uniform-draw sequence yields identical trees.
    The stream is recorded by the oracle (itself pinned to the reference) while it runs the same sweeps."""
    from tnco_b200.engine import RNG_REPLAY
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    mw = float(g['max_width'])
    mw = None if mw < 0 else mw
    n_sweeps = 400
    dims = g['dims'] if 'dims' in g.files and len(g['dims']) else None
    sp, n_projs = golden_sparse(g)
    oc = so.Chain(g['parent'], g['child0'], g['child1'], g['bits'], int(g['n_inds']), dim=int(g['dim']) or 2,
                  dims=dims, max_width=mw, seed=int(g['seed']), sparse_bits=sp, n_projs=n_projs,
                  max_number_new_slices=int(g['max_new']) if 'max_new' in g.files else 0)
    # the constructor's slicer draws precede the recording; re-create the full stream from the seed instead
    betas = [100.0 * s / n_sweeps for s in range(n_sweeps)]
    oc.run(betas, update_slices_every=int(g['every']))
    words = oc.counters()['words_drawn']
    stream = so.mt_stream(int(g['seed']), words + 40 * int(g['n_inds']) + 64 + 3 * len(g['parent']))  # tail >= the engine's per-sweep reserve
    e, _ = _engine(g, RNG_REPLAY)
    e.set_stream(stream[None])
    e.set_betas(betas)
    e.run(n_sweeps)
    pr = e.progress()
    assert pr['sweeps'][0] == n_sweeps
    assert pr['words'][0] == words
    for x, y in zip(e.trees(), oc.tree()):
        assert (x[0] == y).all()
    for x, y in zip(e.trees(True), oc.tree(True)):
        assert (x[0] == y).all()
    t, m = e.costs()
    assert t[0] == oc.total_cost and m[0] == oc.min_total_cost
    assert (e.bits(0) == oc.bits()).all()
    c = oc.counters()
    assert pr['proposals'][0] == c['proposals'] and pr['accepts'][0] == c['accepts']
    if mw is not None:
        assert (e.slices()[0] == oc.slices()).all() and (e.slices(True)[0] == oc.slices(True)).all()
        assert pr['width_rejects'][0] == c['width_rejects']


@pytest.mark.parametrize('n,dim', [(8, 2), (64, 2), (180, 2), (64, 3), (1000, 2)])
def test_eval_cost_matches_oracle(n, dim):
    from tnco_b200.engine import Engine
    ts, ni = regular_network(n, 100 + n)
    trees = [random_tree(ts, ni, s) for s in range(6)]
    e = Engine()
    e.set_network(trees[0][3][:n], ni, dim=dim)
    P, A, B = (np.stack([t[k] for t in trees]) for k in range(3))
    seq, pc, mw = e.eval_cost(P, A, B)
    sl = np.zeros((6, (ni + 31) // 32), np.uint32)
    sl[:, 0] = 0b1011
    seq_s, pc_s, mw_s = e.eval_cost(P, A, B, slices=sl)
    for i, t in enumerate(trees):
        o_seq, o_mw, o_pc = so.tree_cost(t[1], t[2], t[3], ni, dim=dim)
        assert abs(seq[i] - o_seq) <= 1e-12 * o_seq and abs(pc[i] - o_pc) <= 1e-12 * o_pc
        assert abs(mw[i] - o_mw) <= 1e-12 * max(o_mw, 1)
        if dim == 2:
            assert seq[i] == o_seq and pc[i] == o_pc and mw[i] == o_mw
        o_seq, o_mw, o_pc = so.tree_cost(t[1], t[2], t[3], ni, dim=dim, slices=sl[i])
        assert abs(seq_s[i] - o_seq) <= 1e-12 * o_seq and abs(pc_s[i] - o_pc) <= 1e-12 * o_pc
        assert abs(mw_s[i] - o_mw) <= 1e-12 * max(o_mw, 1)


def test_philox_chains_are_valid_and_deterministic():
    """Production RNG: no reference trajectory exists; check invariants the reference's tests check
    (tests/test_utils.py:600-748): cached total == independent recomputation, min <= total, leaves fixed,
    determinism for equal seeds, different seeds differ."""
    from tnco_b200.engine import Engine, random_trees
    ts, ni = regular_network(120, 7)
    lb = so_leaf = None
    from helpers import leaf_bits
    lb = leaf_bits(ts, ni)
    seeds = np.arange(64, dtype=np.uint64) + 11
    p, a, b = random_trees(lb, ni, seeds)
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni)
        e.set_mode()
        e.set_chains(p, a, b, seeds)
        e.set_betas(np.linspace(0, 100, 500, endpoint=False))
        e.run(250)
        e.run(500)
        t, m = e.costs()
        P, A, B = e.trees()
        bP, bA, bB = e.trees(True)
        seq, pc, _ = e.eval_cost(P, A, B)
        assert (pc == t).all() or np.allclose(np.log2(pc), np.log2(t), atol=1e-9)
        bseq, bpc, _ = e.eval_cost(bP, bA, bB)
        assert np.allclose(np.log2(bpc), np.log2(m), atol=1e-9)
        assert (m <= t).all()
        for c in (0, 17, 63):
            nb = e.bits(c)
            o_seq, _, o_pc = so.tree_cost(A[c], B[c], nb, ni)
            assert np.isclose(np.log2(o_pc), np.log2(t[c]), atol=1e-9)
            assert (nb[:120] == lb).all()
        outs.append((t.copy(), m.copy(), P.copy()))
        assert e.counters()['sweeps'] == 64 * 500
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][2] == outs[1][2]).all()
    assert len(set(outs[0][1].tolist())) > 8
    assert np.log2(outs[0][1]).mean() < np.log2(seq).mean() + 1e-9


@pytest.mark.parametrize('dim', [2, 3])
def test_philox_sparse_index_chains_are_valid(dim):
    """Sparse-index cost model under the production RNG (table-cost kernels): the cached totals equal an independent
    evaluation by the oracle's SimpleCostModelSparseInds restatement, and differ from the plain model's."""
    from helpers import leaf_bits
    from tnco_b200.engine import RNG_MT19937, RNG_PHILOX, Engine, EngineError, pack_index_set
    ts, ni = regular_network(100, 9)
    lb = leaf_bits(ts, ni)
    sparse = np.random.default_rng(3).choice(ni, size=15, replace=False).tolist()
    sp = pack_index_set(sparse, ni)
    seeds = np.arange(48, dtype=np.uint64) + 5
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, dim=dim, sparse_bits=sp, n_projs=8).set_mode().generate_chains(seeds)
        e.set_betas(np.linspace(0, 100, 400, endpoint=False))
        t0, _ = e.costs()
        e.run(400)
        t, m = e.costs()
        P, A, B = e.trees()
        seq, pc, _ = e.eval_cost(P, A, B)
        assert np.allclose(np.log2(pc), np.log2(t), atol=1e-9) and (m <= t).all()
        assert np.log2(m).mean() < np.log2(t0).mean() - 1
        for c in (0, 11, 47):
            oc = so.Chain(P[c], A[c], B[c], e.bits(c), ni, dim=dim, sparse_bits=sp, n_projs=8)
            assert abs(np.log2(oc.total_cost) - np.log2(t[c])) < 1e-9
            assert so.Chain(P[c], A[c], B[c], e.bits(c), ni, dim=dim).total_cost >= oc.total_cost
        if rep == 0:  # same initial trees under the plain model cost more: the cap is active
            e2 = Engine()
            e2.set_network(lb, ni, dim=dim).set_mode().generate_chains(seeds)
            p0 = e2.costs()[0]
            assert (p0 >= t0).all() and (p0 > t0).any()
            e2.close()
        outs.append((t.copy(), P.copy()))
        e.close()
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
    # with max_width: MT19937 mode re-slices with the reference's slicer verbatim; the production generator with the
    # production re-slicer in its sparse form (two popcounts per node, the capped width model of
    # finite_width/cost_model/simple_sparse_inds.hpp:39-77) -- either way every sliced width fits, for the current and
    # the best tree, and the cached totals equal an independent evaluation
    for rng in (RNG_PHILOX, RNG_MT19937):
        e = Engine()
        e.set_network(lb, ni, dim=dim, sparse_bits=sp, n_projs=8).set_mode(max_width=12 * np.log2(dim), rng=rng)
        e.generate_chains(seeds[:6] if rng == RNG_MT19937 else seeds).set_betas(np.linspace(0, 100, 200, endpoint=False))
        t0, _ = e.costs()
        e.run(200)
        t, m = e.costs()
        P, A, B = e.trees()
        _, pc, mw = e.eval_cost(P, A, B, slices=e.slices())
        assert np.allclose(np.log2(pc), np.log2(t), atol=1e-9) and (mw <= np.float32(12 * np.log2(dim)) + 1e-6).all()
        bP, bA, bB = e.trees(True)
        bseq, _, bmw = e.eval_cost(bP, bA, bB, slices=e.slices(True))
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9) and (bmw <= np.float32(12 * np.log2(dim)) + 1e-6).all()
        assert (m <= t).all() and np.log2(m).mean() < np.log2(t0).mean()
        if rng == RNG_PHILOX:
            sl = e.slices(True)
            assert (sl != 0).any() and e.progress()['width_rejects'].sum() > 0
        e.close()
    with pytest.raises((EngineError, ValueError), match='n_projs'):
        Engine().set_network(lb, ni, sparse_bits=sp, n_projs=0)


def test_mode_and_resume_argument_errors():
    """Philox kernels are compiled for the app's mode; core-object options and resume states have their own rules."""
    from helpers import leaf_bits
    from tnco_b200._lib import PROB_GREEDY
    from tnco_b200.engine import RNG_MT19937, Engine, EngineError, random_trees
    ts, ni = regular_network(20, 3)
    lb = leaf_bits(ts, ni)
    seeds = np.array([1, 2], np.uint64)
    p, a, b = random_trees(lb, ni, seeds)
    e = Engine()
    # greedy and always acceptance are limits of the production kernel's threshold test (1/beta = 0 / +inf)
    from tnco_b200._lib import PROB_ALWAYS
    e.set_network(lb, ni).set_mode(prob=PROB_GREEDY).set_chains(p, a, b, seeds).set_betas([1.0] * 40)
    t0, _ = e.costs()
    prev = t0
    for s_ in (10, 20, 40):
        e.run(s_)
        t1, m1 = e.costs()
        assert (t1 <= prev).all() and (m1 == t1).all()   # never uphill: the current tree is always the best one
        prev = t1
    assert (prev < t0).any()
    e.set_mode(prob=PROB_ALWAYS).set_chains(p, a, b, seeds).set_betas([50.0] * 20)
    e.run(20)
    c_ = e.counters()
    assert c_['accepts'] == c_['proposals'] > 0
    e.set_prob(0)                                          # back to Metropolis-Hastings on the same chains
    e.run(20)
    e.set_mode(disable_shared_inds=True).set_chains(p, a, b, seeds)
    with pytest.raises((EngineError, ValueError), match='Metropolis'):
        e.costs()
    # greedy acceptance on the stream kernels: costs never go up
    e.set_mode(prob=PROB_GREEDY, rng=RNG_MT19937).set_chains(p, a, b, seeds).set_betas([0.0] * 30)
    t0, _ = e.costs()
    e.run(30)
    assert (e.costs()[0] <= t0).all()
    # resume states: only right after set_chains, generator states only in MT19937 mode, slices only with max_width
    with pytest.raises((EngineError, ValueError), match='already constructed'):
        e.set_resume(mt_state=np.zeros((2, 625), np.uint32))
    e.set_chains(p, a, b, seeds)
    with pytest.raises((EngineError, ValueError), match='max_width'):
        e.set_resume(slices=np.zeros((2, (ni + 31) // 32), np.uint32))
    bad = np.zeros((2, 625), np.uint32)
    bad[:, 624] = 700
    with pytest.raises((EngineError, ValueError), match='generator state'):
        e.set_resume(mt_state=bad)
    e.set_mode().set_chains(p, a, b, seeds)
    with pytest.raises((EngineError, ValueError), match='MT19937'):
        e.set_resume(mt_state=np.zeros((2, 625), np.uint32))
    e.close()


def _check_tree_valid(P, A, B, n):
    """Reference tree invariants (include/tnco/tree.hpp:57-139): leaves first, root last, consistent links."""
    N = 2 * n - 1
    assert P[N - 1] == -1 and (P[:N - 1] >= n).all()
    assert (A[:n] == -1).all() and (B[:n] == -1).all()
    seen = np.zeros(N, int)
    for z in range(n, N):
        assert P[A[z]] == z and P[B[z]] == z and A[z] != B[z]
        seen[A[z]] += 1
        seen[B[z]] += 1
    assert (seen[:N - 1] == 1).all()


def test_philox_skip_slices_are_never_sliced():
    """skip_slices (finite_width/greedy/utils.hpp:76-79) under the production generator: the periodic re-slicer never
    takes a skipped index, the slices it does take still bring every tensor under max_width (the skipped ones are a
    minority here), and cached totals equal an independent evaluation with the slices."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine, random_trees
    ts, ni = regular_network(100, 11)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(32, dtype=np.uint64) + 3
    p, a, b = random_trees(lb, ni, seeds)
    skip = np.zeros((ni + 31) // 32, np.uint32)
    for x in range(0, ni, 5):
        skip[x // 32] |= np.uint32(1 << (x % 32))
    res = []
    for sk in (None, skip):
        e = Engine()
        e.set_network(lb, ni).set_mode(max_width=7.0, update_slices_every=4)
        if sk is not None:
            e.set_skip_slices(sk)
        e.set_chains(p, a, b, seeds).set_betas(np.linspace(0, 40, 300, endpoint=False))
        e.run(300)
        t, m = e.costs()
        P, A, B = e.trees()
        S = e.slices()
        seq, pc, mw = e.eval_cost(P, A, B, slices=S)
        assert np.allclose(np.log2(seq), np.log2(t), atol=1e-9)
        assert (mw <= 7.0).all() and S.any()
        res.append(S.copy())
        e.close()
    assert (res[0] & skip).any()          # without the option those indices do get sliced ...
    assert not (res[1] & skip).any()      # ... with it, never


@pytest.mark.parametrize('n,max_width,tile', [(64, 10, None), (150, 22, None), (150, 22, 16), (300, 30, None)])
def test_philox_finite_width_chains_are_valid(n, max_width, tile):
    """Production finite-width path (fast re-slicer): what the reference's is_valid() checks
    (finite_width/greedy/optimizer.hpp:406-423) -- every sliced width <= max_width for the current and the
    best tree, cached totals equal an independent evaluation with the slices -- plus determinism."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine, random_trees
    ts, ni = regular_network(n, 3 + n)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(48, dtype=np.uint64) + 5
    p, a, b = random_trees(lb, ni, seeds)
    outs = []
    for rep in range(2):
        if tile:
            os.environ['TNB_TILE'] = str(tile)
        e = Engine()
        e.set_network(lb, ni)
        e.set_mode(max_width=max_width, update_slices_every=10)
        e.set_chains(p, a, b, seeds)
        os.environ.pop('TNB_TILE', None)   # (the tile shape is fixed when the chains are created)
        if tile:
            assert e.config()['tile'] == tile
        e.set_betas(np.linspace(0, 100, 400, endpoint=False))
        t0, _ = e.costs()
        e.run(200)
        e.run(400)
        t, m = e.costs()
        P, A, B = e.trees()
        S = e.slices()
        seq, pc, mw = e.eval_cost(P, A, B, slices=S)
        assert np.allclose(np.log2(seq), np.log2(t), atol=1e-9)
        assert (mw <= max_width).all()
        bP, bA, bB = e.trees(True)
        bS = e.slices(True)
        bseq, bpc, bmw = e.eval_cost(bP, bA, bB, slices=bS)
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9)
        assert (bmw <= max_width).all()
        assert (m <= t * (1 + 1e-12)).all()
        for c in (0, 13, 47):
            _check_tree_valid(P[c], A[c], B[c], n)
            nb = e.bits(c)
            assert (nb[:n] == lb).all()
            o_seq, o_mw, o_pc = so.tree_cost(A[c], B[c], nb, ni, slices=S[c])
            assert np.isclose(np.log2(o_seq), np.log2(t[c]), atol=1e-9) and o_mw <= max_width
        pr = e.progress()
        assert (pr['sweeps'] == 400).all() and pr['width_rejects'].sum() > 0
        outs.append((t.copy(), m.copy(), P.copy(), S.copy()))
        e.close()
    assert all((x == y).all() for x, y in zip(outs[0], outs[1]))
    assert np.log2(outs[0][1]).mean() < np.log2(t0).mean()


@pytest.mark.parametrize('n,method', [(2, 0), (3, 1), (64, 0), (64, 1), (180, 0), (1000, 0)])
def test_device_generated_trees_are_valid(n, method):
    """tnb_generate_chains: every chain gets a valid tree whose contractions all share an index
    (check_shared_inds, include/tnco/ctree.hpp:101-152), deterministic per seed, different across seeds."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    if n <= 3:
        ts, ni = [[k for k in range(n - 1) if k in (t - 1, t)] for t in range(n)], n - 1   # a path graph
    else:
        ts, ni = regular_network(n, 50 + n)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(24, dtype=np.uint64) * 7 + 3
    trees = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni).set_mode()
        e.generate_chains(seeds, method=method)
        P, A, B = e.trees()
        t, m = e.costs()
        seq, pc, _ = e.eval_cost(P, A, B)
        assert (pc == t).all()
        for c in range(len(seeds)):
            _check_tree_valid(P[c], A[c], B[c], n)
        nb = e.bits(5)
        for z in range(n, 2 * n - 1):
            assert (nb[A[5][z]] & nb[B[5][z]]).any(), 'contracted pair shares no index'
        # the same trees through the validating entry point
        e2 = Engine()
        e2.set_network(lb, ni).set_mode()
        e2.set_chains(P, A, B, seeds)
        assert (e2.costs()[0] == t).all()
        e2.close()
        trees.append((P.copy(), t.copy()))
        e.set_betas(np.linspace(0, 100, 50, endpoint=False))
        e.run(50)
        assert (e.costs()[1] <= t).all()
        e.close()
    assert (trees[0][0] == trees[1][0]).all()
    if n >= 64:
        assert len({tuple(r) for r in trees[0][0].tolist()}) > 12
    if method == 0 and n >= 64:  # greedy is much better than random merges
        e = Engine()
        e.set_network(lb, ni).set_mode()
        e.generate_chains(seeds, method=1)
        assert np.log2(trees[0][1]).mean() < np.log2(e.costs()[0]).mean()
        e.close()


@pytest.mark.parametrize('n,method', [(12, 0), (48, 0), (48, 1), (90, 0)])
def test_device_generated_trees_for_hyper_index_networks(n, method):
    """tnb_generate_chains on networks WITH hyper-indices and open indices: valid trees whose contractions share an
    index under the hyper-count rule (the validating entry point tnb_set_chains re-derives it on the host and accepts
    them, and its index sets equal the oracle-side construction), deterministic, and better than random for greedy."""
    from tnco_b200.engine import Engine, random_trees
    ts, ni, out, lb, ob = _hyper_case(n, 500 + n)
    seeds = np.arange(20, dtype=np.uint64) * 5 + 2
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, output_bits=ob).set_mode()
        assert e.hyper
        e.generate_chains(seeds, method=method)
        P, A, B = e.trees()
        t, m = e.costs()
        for c in range(len(seeds)):
            _check_tree_valid(P[c], A[c], B[c], n)
        e2 = Engine()
        e2.set_network(lb, ni, output_bits=ob).set_mode()
        e2.set_chains(P, A, B, seeds)       # host-side check_shared_inds + hyper-count rule
        assert (e2.costs()[0] == t).all() and (e2.bits(3) == e.bits(3)).all()
        e2.close()
        e.set_betas(np.linspace(0, 100, 60, endpoint=False))
        e.run(60)
        assert (e.costs()[1] <= t).all()
        outs.append((P.copy(), t.copy()))
        e.close()
    assert (outs[0][0] == outs[1][0]).all()
    assert len({tuple(r) for r in outs[0][0].tolist()}) > (3 if n < 20 else 10)
    if method == 0 and n >= 48:
        hp, ha, hb = random_trees(lb, ni, seeds, method=1, output_bits=ob)
        e = Engine()
        e.set_network(lb, ni, output_bits=ob).set_mode()
        e.set_chains(hp, ha, hb, seeds)
        assert np.log2(outs[0][1]).mean() < np.log2(e.costs()[0]).mean()
        e.close()


def test_device_generated_trees_reject_disconnected_network():
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    lb = leaf_bits([[0], [0], [1], [1]], 2)
    e = Engine()
    e.set_network(lb, 2).set_mode()
    with pytest.raises(ValueError, match='not connected'):
        e.generate_chains([1, 2, 3])
    e.close()


@pytest.mark.parametrize('max_width', [None, 20])
def test_split_layout_gives_identical_results(max_width):
    """TNB_LAYOUT_SPLIT (headers and index sets apart, picked automatically for HBM-sized batches) and
    TNB_LAYOUT_INTERLEAVED are two placements of the same state: same seeds -> identical chains."""
    from helpers import leaf_bits
    from tnco_b200._lib import LAYOUT_INTERLEAVED, LAYOUT_SPLIT
    from tnco_b200.engine import Engine
    ts, ni = regular_network(140, 77)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(40, dtype=np.uint64) + 9
    outs = []
    for layout in (LAYOUT_INTERLEAVED, LAYOUT_SPLIT):
        e = Engine()
        e.set_network(lb, ni).set_mode(max_width=max_width, layout=layout)
        e.generate_chains(seeds)
        assert e.config()['layout'] == layout
        e.set_betas(np.linspace(0, 100, 300, endpoint=False))
        e.run(300)
        outs.append((e.costs(), e.trees(), e.trees(True), e.slices(True), e.bits(7), e.progress()['proposals']))
        e.close()
    a, b = outs
    assert (a[0][0] == b[0][0]).all() and (a[0][1] == b[0][1]).all()
    for k in (1, 2):
        assert all((x == y).all() for x, y in zip(a[k], b[k]))
    assert (a[3] == b[3]).all() and (a[4] == b[4]).all() and (a[5] == b[5]).all()


@pytest.mark.parametrize('n,tile', [(64, 4), (64, 8), (64, 32), (180, 16), (180, 32), (300, 32)])
def test_shared_memory_resident_layout_gives_identical_results(n, tile):
    """TNB_LAYOUT_SMEM (chain state resident in shared memory for the whole launch, north_star item 4) runs the same
    chains as the in-place layouts: same seeds -> identical trees, costs, best trees, index sets, counters -- over two
    launches (state goes home to global memory in between)."""
    from helpers import leaf_bits
    from tnco_b200._lib import LAYOUT_INTERLEAVED, LAYOUT_SMEM
    from tnco_b200.engine import Engine
    ts, ni = regular_network(n, 70 + n)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(70, dtype=np.uint64) + 9
    outs = []
    for layout in (LAYOUT_INTERLEAVED, LAYOUT_SMEM):
        os.environ['TNB_TILE'] = str(tile)
        try:
            e = Engine()
            e.set_network(lb, ni).set_mode(layout=layout)
            e.generate_chains(seeds)
        finally:
            os.environ.pop('TNB_TILE', None)
        cfg = e.config()
        assert cfg['layout'] == layout and cfg['tile'] == tile, cfg
        e.set_betas(np.linspace(0, 100, 400, endpoint=False))
        e.run(150)
        e.run(400)
        outs.append((e.costs(), e.trees(), e.trees(True), e.bits(7), e.bits(69), e.progress()['proposals'],
                     e.progress()['accepts']))
        e.close()
    a, b = outs
    assert (a[0][0] == b[0][0]).all() and (a[0][1] == b[0][1]).all()
    for k in (1, 2):
        assert all((x == y).all() for x, y in zip(a[k], b[k]))
    assert all((a[k] == b[k]).all() for k in (3, 4, 5, 6))


def _hyper_case(n, seed):
    from helpers import hyper_network, leaf_bits
    from tnco_b200.engine import pack_index_set
    ts, ni, out = hyper_network(n, seed)
    return ts, ni, out, leaf_bits(ts, ni), pack_index_set(out, ni)


@pytest.mark.parametrize('n,max_width', [(40, None), (90, None), (90, 18)])
def test_philox_hyper_index_networks_are_valid(n, max_width):
    """Networks WITH hyper-indices through the production path (HYPER kernels, host tree builder): cached totals
    equal an independent evaluation by the oracle on the engine's own index sets, index sets obey the hyper-count
    rule (they depend on the subtree only), sliced widths fit, leaves fixed, determinism."""
    from tnco_b200.engine import Engine, random_trees
    ts, ni, out, lb, ob = _hyper_case(n, 5 + n)
    seeds = np.arange(40, dtype=np.uint64) + 3
    for method in (0, 1):
        p, a, b = random_trees(lb, ni, seeds, method=method, output_bits=ob)
        for c in (0, 11):
            _check_tree_valid(p[c], a[c], b[c], n)
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, output_bits=ob).set_mode(max_width=max_width)
        assert e.hyper and e.config()['tile'] == 32
        e.set_chains(p, a, b, seeds)
        e.set_betas(np.linspace(0, 100, 300, endpoint=False))
        t0, _ = e.costs()
        e.run(300)
        t, m = e.costs()
        P, A, B = e.trees()
        S = e.slices() if max_width is not None else None
        seq, pc, mw = e.eval_cost(P, A, B, slices=S)
        assert np.allclose(np.log2(seq), np.log2(t), atol=1e-9)
        if max_width is not None:
            assert (mw <= max_width).all()
        bP, bA, bB = e.trees(True)
        bseq, _, bmw = e.eval_cost(bP, bA, bB, slices=e.slices(True) if max_width is not None else None)
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9)
        for c in (0, 17, 39):
            _check_tree_valid(P[c], A[c], B[c], n)
            nb = e.bits(c)
            assert (nb[:n] == lb).all()
            # hyper-count rule: inds(z) = held below z AND (held outside z, or output)
            below = nb[:n].copy()
            full = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
            full[:n] = lb
            cnt_total = np.zeros(ni, int)
            for x in range(n):
                for i in so_positions(lb[x]):
                    cnt_total[i] += 1
            sub = [None] * (2 * n - 1)
            for x in range(n):
                sub[x] = {i: 1 for i in so_positions(lb[x])}
            order = []
            stack = [2 * n - 2]
            while stack:
                z = stack.pop()
                if z >= n:
                    order.append(z)
                    stack += [A[c][z], B[c][z]]
            for z in reversed(order):
                d = dict(sub[A[c][z]])
                for i, k in sub[B[c][z]].items():
                    d[i] = d.get(i, 0) + k
                sub[z] = d
                want = sorted(i for i, k in d.items() if k < cnt_total[i] or i in out)
                assert so_positions(nb[z]) == want, (c, z)
            o_seq, o_mw, o_pc = so.tree_cost(A[c], B[c], nb, ni, slices=None if S is None else S[c])
            assert np.isclose(np.log2(o_seq), np.log2(t[c]), atol=1e-9)
        outs.append((t.copy(), m.copy(), P.copy()))
        e.close()
    assert all((x == y).all() for x, y in zip(outs[0], outs[1]))
    assert np.log2(outs[0][1]).mean() < np.log2(t0).mean()


def so_positions(row):
    return [w * 32 + b for w, v in enumerate(np.asarray(row).tolist()) for b in range(32) if (v >> b) & 1]


@pytest.mark.parametrize('n,max_width,hyper', [(80, None, False), (80, 30, False), (60, 26, True)])
def test_philox_per_index_dims(n, max_width, hyper):
    """Per-index dimensions (powers of two): production path against the oracle's dims-vector cost model
    (infinite_memory/cost_model/simple.hpp:46-54, finite_width/cost_model/simple.hpp:47-55) on the engine's own trees
    and index sets; sliced widths (sum of log2 dims) fit; whole indices are sliced; determinism."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine, random_trees
    if hyper:
        ts, ni, out, lb, ob = _hyper_case(n, 31)
    else:
        ts, ni = regular_network(n, 31)
        lb, ob, out = leaf_bits(ts, ni), None, []
    dims = np.random.default_rng(5).choice([2, 4, 8, 16], size=ni).astype(np.uint64)
    l2 = np.log2(dims.astype(float))
    seeds = np.arange(32, dtype=np.uint64) + 2
    p, a, b = random_trees(lb, ni, seeds, output_bits=ob)
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, dims=dims, output_bits=ob).set_mode(max_width=max_width)
        e.set_chains(p, a, b, seeds)
        e.set_betas(np.linspace(0, 100, 300, endpoint=False))
        t0, _ = e.costs()
        e.run(300)
        t, m = e.costs()
        P, A, B = e.trees()
        S = e.slices() if max_width is not None else None
        seq, pc, mw = e.eval_cost(P, A, B, slices=S)
        assert np.allclose(np.log2(seq), np.log2(t), atol=1e-9)
        for c in (0, 9, 31):
            nb = e.bits(c)
            assert (nb[:n] == lb).all()
            o_seq, o_mw, o_pc = so.tree_cost(A[c], B[c], nb, ni, dims=dims, slices=None if S is None else S[c])
            assert np.isclose(np.log2(o_seq), np.log2(t[c]), atol=1e-9)
            assert np.isclose(o_mw, mw[c])
            if max_width is not None:
                assert o_mw <= max_width
                # max log2 width recomputed from the index sets, slices removed
                keep = [i for i in range(ni) if not (int(S[c][i >> 5]) >> (i & 31)) & 1]
                w = max(sum(l2[i] for i in keep if (int(row[i >> 5]) >> (i & 31)) & 1) for row in nb)
                assert w <= max_width
        bP, bA, bB = e.trees(True)
        bseq, _, bmw = e.eval_cost(bP, bA, bB, slices=e.slices(True) if max_width is not None else None)
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9)
        if max_width is not None:
            assert (bmw <= max_width).all() and e.progress()['width_rejects'].sum() > 0
        outs.append((t.copy(), m.copy(), P.copy()))
        e.close()
    assert all((x == y).all() for x, y in zip(outs[0], outs[1]))
    assert np.log2(outs[0][1]).mean() < np.log2(t0).mean()


def test_per_index_dims_must_be_powers_of_two():
    """(Name kept for the emulation suite.)  Dimensions that are neither all equal nor all powers of two run the
    reference's sequential product / sum loops: any positive integers are accepted; zero is not."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    ts, ni = regular_network(10, 1)
    dims = np.full(ni, 2, np.uint64)
    dims[3] = 3
    e = Engine()
    e.set_network(leaf_bits(ts, ni), ni, dims=dims)
    e.set_network(leaf_bits(ts, ni), ni, dims=np.full(ni, 3, np.uint64))   # uniform: any integer dimension
    dims[3] = 0
    with pytest.raises(ValueError, match='positive'):
        e.set_network(leaf_bits(ts, ni), ni, dims=dims)
    e.close()


@pytest.mark.parametrize('n,max_width,sparse', [(60, None, False), (60, 16.0, False), (60, 16.0, True)])
def test_philox_general_per_index_dims(n, max_width, sparse):
    """General per-index dimensions under the production RNG (table-cost kernels, sequential cost / width loops):
    cached totals equal the oracle's dims-vector evaluation of the engine's own trees; sliced widths fit."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine, pack_index_set
    ts, ni = regular_network(n, 41)
    lb = leaf_bits(ts, ni)
    dims = np.random.default_rng(6).choice([2, 3, 5, 6, 7], size=ni).astype(np.uint64)
    sp = pack_index_set(list(range(0, ni, 7)), ni) if sparse else None
    seeds = np.arange(24, dtype=np.uint64) + 3
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, dims=dims, sparse_bits=sp, n_projs=9 if sparse else None).set_mode(max_width=max_width)
        e.generate_chains(seeds).set_betas(np.linspace(0, 100, 250, endpoint=False))
        t0, _ = e.costs()
        e.run(250)
        t, m = e.costs()
        P, A, B = e.trees()
        S = e.slices() if max_width is not None else None
        seq, pc, mw = e.eval_cost(P, A, B, slices=S)
        assert np.allclose(np.log2(pc), np.log2(t), atol=1e-9) and (m <= t).all()
        if max_width is not None:
            assert (mw <= np.float32(max_width) + 1e-5).all() and e.progress()['width_rejects'].sum() > 0
        for c in (0, 7, 23):
            oc = so.Chain(P[c], A[c], B[c], e.bits(c), ni, dims=dims, sparse_bits=sp, n_projs=9 if sparse else None,
                          max_width=max_width)
            if max_width is None:
                assert abs(np.log2(oc.total_cost) - np.log2(t[c])) < 1e-9
        outs.append((t.copy(), P.copy()))
        e.close()
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
    assert np.log2(outs[0][0]).mean() < np.log2(t0).mean()


def test_packed_tree_readback_and_cached_engine():
    """tnb_get_trees_packed == tnb_get_trees (both current and best trees); cached_engine hands out one engine per
    device and survives back-to-back reconfiguration."""
    from helpers import leaf_bits
    from tnco_b200.engine import cached_engine, release_cached_engines
    ts, ni = regular_network(50, 3)
    lb = leaf_bits(ts, ni)
    e = cached_engine(0)
    assert cached_engine(0) is e
    for rep in range(2):
        e.set_network(lb, ni).set_mode(max_width=None if rep == 0 else 12)
        e.generate_chains(np.arange(20, dtype=np.uint64) + rep)
        e.set_betas(np.linspace(0, 100, 100, endpoint=False))
        e.run(100)
        for best in (False, True):
            p, a, b = e.trees(best=best)
            w = e.trees_packed(best=best)
            assert ((w & 0xffff) == a[:, 50:]).all() and ((w >> 16) == b[:, 50:]).all()
            w2 = e.trees_packed(best=best, chain0=7, n=3)
            assert (w2 == w[7:10]).all()
    release_cached_engines()
    assert cached_engine(0) is not e
    release_cached_engines()


def test_tile_shape_follows_the_batch_size():
    """Lanes per chain are chosen per batch: the widest tile whose batch fits one wave of resident warps, else the
    narrowest that holds the index set -- and the choice never changes results (same seeds, Philox)."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    ts, ni = regular_network(64, 0)      # 96 indices: 3 words, tiles of 4 .. 32 lanes are possible
    lb = leaf_bits(ts, ni)
    seeds = np.arange(48, dtype=np.uint64) + 1
    outs = {}
    for tile in (4, 8, 16, 32):
        os.environ['TNB_TILE'] = str(tile)
        e = Engine()
        e.set_network(lb, ni).set_mode()
        e.generate_chains(seeds)
        os.environ.pop('TNB_TILE', None)
        assert e.config()['tile'] == tile
        e.set_betas(np.linspace(0, 100, 200, endpoint=False))
        e.run(200)
        outs[tile] = (e.costs()[1].copy(), e.trees_packed(best=True).copy())
        e.close()
    for tile in (8, 16, 32):
        assert (outs[tile][0] == outs[4][0]).all() and (outs[tile][1] == outs[4][1]).all()
    e = Engine()
    e.set_network(lb, ni).set_mode()
    e.generate_chains(seeds)
    assert e.config()['tile'] == 32          # a small batch: maximum parallelism
    e.close()


def test_minima_stay_exact_when_chains_wander_at_small_beta():
    """Regression: on a slow beta ramp the chains first wander up to costs around 2^100 and come back; the running
    total of the production kernel then carries the rounding of the large terms and once produced totals of the
    wrong sign that were recorded as unbeatable minima.  Minima must stay positive and equal to the exact cost of
    the recorded best tree."""
    from tnco_b200 import networks
    from tnco_b200.engine import Engine, pack_leaf_bits
    ts, ni = networks.grid_rqc(6, 6, 12)
    lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(96, dtype=np.uint64) + 1
    e = Engine()
    e.set_network(lb, ni).set_mode()
    e.generate_chains(seeds)
    n = 1000000
    e.set_betas(np.array([k * (100.0 / n) for k in range(4000)]))   # the first 4000 sweeps of a 10^6-sweep ramp
    peak = 0.0
    for upto in (500, 1500, 4000):
        e.run(upto)
        t, m = e.costs()
        peak = max(peak, float(np.log2(t.max())))
        assert (m > 0).all() and (t > 0).all()
        bP, bA, bB = e.trees(True)
        bseq, _, _ = e.eval_cost(bP, bA, bB)
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9)
    assert peak > 70          # the scenario really occurred: totals far above what fp64 sums exactly
    e.close()


def _ring_network(n):
    """n tensors in a ring: tensor t holds indices t and (t + 1) % n -- n indices, all of dimension 2."""
    return [[t, (t + 1) % n] for t in range(n)], n


def _caterpillar(n, n_chains):
    """The deepest tree: ((((0,1),2),3),...) -- a sweep from leaf 0 walks n - 2 levels."""
    N = 2 * n - 1
    P, A, B = (np.full((n_chains, N), -1, np.int32) for _ in range(3))
    for k in range(n - 1):
        z = n + k
        a, b = (0, 1) if k == 0 else (z - 1, k + 1)
        A[:, z], B[:, z] = a, b
        P[:, a], P[:, b] = z, z
    return P, A, B


def test_networks_of_1024_indices_take_two_words_per_lane():
    """The one-word-per-lane production kernels pack popcounts into 10-bit fields (no popcount reaches 1024 there):
    a network of exactly 1024 indices fits 32 words but must run with two words per lane."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    for n in (1023, 1024):
        ts, ni = _ring_network(n)
        lb = leaf_bits(ts, ni)
        seeds = np.arange(16, dtype=np.uint64) + 3
        for mw in (None, 3.0):
            e = Engine()
            e.set_network(lb, ni).set_mode(max_width=mw)
            e.generate_chains(seeds)
            assert e.config()['words_per_lane'] == (2 if n == 1024 else 1)
            e.set_betas(np.linspace(0, 50, 100, endpoint=False))
            e.run(100)
            t, m = e.costs()
            P, A, B = e.trees()
            bP, bA, bB = e.trees(True)
            S = e.slices() if mw is not None else None
            bS = e.slices(True) if mw is not None else None
            _, pc, w = e.eval_cost(P, A, B, slices=S)
            _, bpc, bw = e.eval_cost(bP, bA, bB, slices=bS)
            assert np.allclose(np.log2(pc), np.log2(t), atol=1e-9) and np.allclose(np.log2(bpc), np.log2(m), atol=1e-9)
            if mw is not None:
                assert (bw <= mw).all() and (w <= mw).all()
            e.close()


@pytest.mark.parametrize('max_width', [None, 3.0])
def test_best_tree_snapshots_on_very_deep_trees(max_width):
    """The incremental best-tree snapshot lists the nodes a sweep walked from a 64-entry ring; sweeps with more
    levels (here up to 198, from caterpillar trees) must fall back to the full copy, and either way the recorded
    best tree has to cost exactly min_total."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine
    ts, ni = _ring_network(200)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(32, dtype=np.uint64) + 5
    P0, A0, B0 = _caterpillar(200, 32)
    e = Engine()
    e.set_network(lb, ni).set_mode(max_width=max_width)
    e.set_chains(P0, A0, B0, seeds)
    e.set_betas(np.linspace(0.5, 100, 300, endpoint=False))
    levels = []
    for until in (1, 20, 100, 300):
        c0 = e.counters()
        e.run(until)
        c1 = e.counters()
        levels.append((c1['proposals'] - c0['proposals']) / max(c1['sweeps'] - c0['sweeps'], 1))
        t, m = e.costs()
        bP, bA, bB = e.trees(True)
        _, bpc, bw = e.eval_cost(bP, bA, bB, slices=e.slices(True) if max_width is not None else None)
        assert np.allclose(np.log2(bpc), np.log2(m), atol=1e-9)
        assert (m <= t * (1 + 1e-12)).all()
    assert levels[0] > 64     # the first sweeps are longer than the ring ...
    assert levels[-1] < 64    # ... the later ones are not
    e.close()


@pytest.mark.parametrize('dim', [3, 5])
def test_philox_uniform_dimension_other_than_two_with_max_width(dim):
    """A uniform dimension d != 2 keeps widths proportional to popcounts, so the table-cost production kernel takes
    the production re-slicer too (costs from the d^k table, full re-cost instead of the incremental one): every
    sliced width <= max_width for the current and the best tree, cached totals equal an independent evaluation,
    results are deterministic, and the anneal improves on the initial trees."""
    from helpers import leaf_bits
    from tnco_b200.engine import Engine, random_trees
    ts, ni = regular_network(120, 77)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(40, dtype=np.uint64) + 9
    p, a, b = random_trees(lb, ni, seeds)
    max_width = 14 * float(np.log2(dim))
    outs = []
    for rep in range(2):
        e = Engine()
        e.set_network(lb, ni, dim=dim).set_mode(max_width=max_width, update_slices_every=5)
        e.set_chains(p, a, b, seeds)
        e.set_betas(np.linspace(0, 60, 300, endpoint=False))
        t0, _ = e.costs()
        e.run(150)
        e.run(300)
        t, m = e.costs()
        P, A, B = e.trees()
        seq, pc, mw = e.eval_cost(P, A, B, slices=e.slices())
        assert np.allclose(np.log2(seq), np.log2(t), atol=1e-9) and (mw <= np.float32(max_width) + 1e-5).all()
        bP, bA, bB = e.trees(True)
        bseq, _, bmw = e.eval_cost(bP, bA, bB, slices=e.slices(True))
        assert np.allclose(np.log2(bseq), np.log2(m), atol=1e-9) and (bmw <= np.float32(max_width) + 1e-5).all()
        assert (m <= t * (1 + 1e-12)).all() and np.log2(m).mean() < np.log2(t0).mean()
        pr = e.progress()
        assert pr['width_rejects'].sum() > 0 and (pr['sweeps'] == 300).all()
        outs.append((t.copy(), m.copy(), P.copy(), e.slices().copy()))
        e.close()
    assert all((x == y).all() for x, y in zip(outs[0], outs[1]))
