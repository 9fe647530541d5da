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
