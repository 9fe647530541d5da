"""north_star correctness part 3: the best-cost distribution over seeds of the production engine (Philox RNG,
log-free threshold acceptance, fast re-slicer) is statistically no worse than the reference's at equal sweep
counts.  The reference distribution comes from the CPU oracle (bit-identical to the reference core, see
tests/test_oracle.py) run on the same network from the same family of initial trees."""
import numpy as np
import pytest
from scipy import stats

from helpers import leaf_bits, regular_network
from oracle import sa_oracle as so

pytestmark = pytest.mark.gpu


def _reference_distribution(lb, ni, n, seeds, n_sweeps, max_width):
    from tnco_b200.engine import random_trees
    P, A, B = random_trees(lb, ni, np.asarray(seeds, np.uint64))
    betas = [100.0 * s / n_sweeps for s in range(n_sweeps)]
    out = []
    for k, s in enumerate(seeds):
        nb = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
        nb[:n] = lb
        for z in range(n, 2 * n - 1):
            nb[z] = nb[A[k][z]] ^ nb[B[k][z]]
        oc = so.Chain(P[k], A[k], B[k], nb, ni, seed=int(s), max_width=max_width)
        oc.run(betas, update_slices_every=10)
        out.append(oc.log2_min_total_cost)
    return np.array(out), (P, A, B)


@pytest.mark.parametrize('n,max_width,n_sweeps', [(64, None, 1500), (100, None, 1500), (100, 14, 1500)])
def test_best_cost_distribution_is_no_worse_than_the_reference(n, max_width, n_sweeps):
    from tnco_b200.engine import Engine, random_trees
    ts, ni = regular_network(n, 1234 + n)
    lb = leaf_bits(ts, ni)
    ref, _ = _reference_distribution(lb, ni, n, np.arange(48) + 1, n_sweeps, max_width)
    seeds = np.arange(384, dtype=np.uint64) + 1000
    P, A, B = random_trees(lb, ni, seeds)          # same initial-tree generator as the reference arm
    e = Engine()
    e.set_network(lb, ni).set_mode(max_width=max_width, update_slices_every=10)
    e.set_chains(P, A, B, seeds)
    e.set_betas([100.0 * s / n_sweeps for s in range(n_sweeps)])
    e.run(n_sweeps)
    ours = np.log2(e.costs()[1])
    e.close()
    # one-sided Mann-Whitney U: H1 = "ours is stochastically LARGER (worse) than the reference"
    p_worse = stats.mannwhitneyu(ours, ref, alternative='greater').pvalue
    assert p_worse > 1e-3, (ours.mean(), ref.mean(), p_worse)
    # and the means agree within 4 standard errors + 1 % (guards against a silently broken acceptance rule)
    se = np.sqrt(ours.var() / len(ours) + ref.var() / len(ref))
    assert ours.mean() <= ref.mean() + 4 * se + 0.01 * abs(ref.mean()), (ours.mean(), ref.mean(), se)
    assert ours.mean() >= ref.mean() - 6 * se - 0.03 * abs(ref.mean()), (ours.mean(), ref.mean(), se)


# ---- the benchmarked networks, at the benchmark's sweep count, against the UNMODIFIED reference core.
# tests/golden/stat_*.json hold 256 reference runs each (scripts/make_golden_stats.py, oracle/_ref driven like core_).
@pytest.mark.parametrize('case,chains', [('c1_1e4', 1024), ('c2_1e4', 1024), ('c3_1e4', 1024), ('c4_1e4', 1024),
                                         ('c4_3e4', 768)])
def test_equal_sweep_distribution_on_the_benchmarked_networks(case, chains):
    import json
    import os

    from helpers import GOLDEN
    from tnco_b200 import networks
    from tnco_b200.engine import Engine, pack_leaf_bits, random_trees
    fx = json.load(open(os.path.join(GOLDEN, f'stat_{case}.json')))
    ref = np.array(fx['log2_min_total_cost'])
    assert len(ref) >= 256 and fx['n_sweeps'] >= 10000
    ts, ni = eval('networks.' + fx['network'])
    lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(chains, dtype=np.uint64) + 100001
    P, A, B = random_trees(lb, ni, seeds)          # same initial-tree generator as the reference arm
    n_sweeps = fx['n_sweeps']
    e = Engine()
    e.set_network(lb, ni).set_mode(max_width=fx['max_width'], update_slices_every=fx['update_slices'])
    e.set_chains(P, A, B, seeds)
    e.set_betas([100.0 * s / n_sweeps for s in range(n_sweeps)])
    e.run(n_sweeps)
    t, m = e.costs()
    # the returned minima are real: re-evaluate every best tree (+ slices) from scratch
    bp, ba, bb = e.trees(best=True)
    sl = e.slices(best=True) if fx['max_width'] is not None else None
    seq, pc, mw = e.eval_cost(bp, ba, bb, slices=sl)
    assert np.allclose(np.log2(pc), np.log2(m), atol=1e-9)
    if fx['max_width'] is not None:
        assert (mw <= fx['max_width']).all()
    e.close()
    ours = np.log2(m)
    p_worse = stats.mannwhitneyu(ours, ref, alternative='greater').pvalue
    se = np.sqrt(ours.var() / len(ours) + ref.var() / len(ref))
    print(f'{case}: ours mean {ours.mean():.4f} min {ours.min():.4f} | reference mean {ref.mean():.4f} '
          f'min {ref.min():.4f} | p(ours worse) {p_worse:.3g} | se {se:.4f}')
    assert p_worse > 1e-3, (ours.mean(), ref.mean(), p_worse)
    assert ours.mean() <= ref.mean() + 4 * se, (ours.mean(), ref.mean(), se)


# ---- the production re-slicer on the table-cost kernels (uniform dimension 3; the sparse-index width model) against
# the reference's slicer, which the same kernels run verbatim when TNB_VERBATIM_RESLICER is set (bit-exact against
# the reference in MT19937 mode): same generator, same initial trees, equal sweep counts.
@pytest.mark.parametrize('case', ['d3', 'd2_sparse', 'd3_sparse'])
def test_production_reslicer_is_no_worse_than_the_reference_slicer(case):
    import os

    from tnco_b200 import networks
    from tnco_b200.engine import Engine, pack_index_set, pack_leaf_bits, random_trees
    ts, ni = networks.regular_graph(200, 0)
    lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(1536, dtype=np.uint64) + 1
    P, A, B = random_trees(lb, ni, seeds)
    dim = 3 if case.startswith('d3') else 2
    sparse = case.endswith('sparse')
    kw = dict(sparse_bits=pack_index_set(np.random.default_rng(3).choice(ni, size=40, replace=False).tolist(), ni),
              n_projs=16) if sparse else {}
    max_width = (18.0 if sparse else 20.0) * float(np.log2(dim))
    out = {}
    for verbatim in (False, True):
        if verbatim:
            os.environ['TNB_VERBATIM_RESLICER'] = '1'
        try:
            e = Engine()
            e.set_network(lb, ni, dim=dim, **kw).set_mode(max_width=max_width, update_slices_every=10)
            e.set_chains(P, A, B, seeds)
        finally:
            os.environ.pop('TNB_VERBATIM_RESLICER', None)
        e.set_betas(np.linspace(0, 100, 2000, endpoint=False))
        e.run(2000)
        t, m = e.costs()
        bp, ba, bb = e.trees(best=True)
        _, pc, mw = e.eval_cost(bp, ba, bb, slices=e.slices(best=True))
        assert np.allclose(np.log2(pc), np.log2(m), atol=1e-9) and (mw <= np.float32(max_width) + 1e-5).all()
        out[verbatim] = np.log2(m)
        e.close()
    ours, ref = out[False], out[True]
    p_worse = stats.mannwhitneyu(ours, ref, alternative='greater').pvalue
    se = np.sqrt(ours.var() / len(ours) + ref.var() / len(ref))
    print(f'{case}: production re-slicer mean {ours.mean():.4f} | reference slicer mean {ref.mean():.4f} | '
          f'p(ours worse) {p_worse:.3g} | se {se:.4f}')
    assert p_worse > 1e-3 and ours.mean() <= ref.mean() + 4 * se, (ours.mean(), ref.mean(), p_worse, se)
