"""Trajectory parity of the PRODUCTION kernel (the instantiation bench.py times: Philox RNG, running total,
threshold-form acceptance, fast re-slicer) on the benchmarked networks C2 / C3 / C4 (and C1 / a sub-warp tile).

The kernel records every decision it takes (tnb_set_trace: leaf of every sweep, D/E choice, width gate, accept bit,
delta, total, the event's uniform, candidate slices and keep decision of every re-slice).  The decisions are turned
into the raw 32-bit draw stream the reference would have had to see to take them, and the CPU oracle -- the plain-C
restatement of Optimizer::update() that is pinned to the reference, oracle/sa_oracle.c -- is run on that stream from
the same initial tree ("may additionally take the recorded accept bit").  Required:

  * the oracle consumes exactly the synthesized stream (same coins drawn, same width gates: the draw order is
    data-dependent, so any disagreement about index sets desynchronises it);
  * identical trees, identical index sets of every node, identical contraction cost of every node (bit-exact: with
    d = 2 every cost is an exact power of two), identical delta of every proposal (bit-exact), best tree identical;
  * the kernel's running total equals the reference's partial_cost.back() within 1e-9 relative at every proposal
    (the reference sums in tree order, the kernel keeps a running sum that is re-based every 64 sweeps);
  * the kernel's accept bit equals the reference's rule u <= pow(1 + delta/total, -beta) (prob/mh.hpp:45-59) evaluated
    in fp64 on the recorded (delta, total, u, beta), except where u is within 2e-5 relative of the boundary (the
    kernel evaluates the same rule solved for delta with an fp32 exp2) -- and such proposals are < 0.1 % of all;
  * every re-slice keep / discard decision equals the reference's criterion (strictly cheaper, finite_width/greedy/
    optimizer.hpp:362-375) evaluated by the oracle on the candidate slices, except at exact-sum ties (1e-12).
"""
import os

import numpy as np
import pytest

from oracle import sa_oracle as so

pytestmark = pytest.mark.gpu

NETS = {
    'C1': ('regular_graph(64, 0)', None),
    'C2': ('grid_rqc(6, 6, 12)', None),
    'C3': ('sycamore(14)', None),
    'C4': ('sycamore(20)', 32.0),
    'C4w24': ('sycamore(20)', 24.0),
    'reg200_fw': ('regular_graph(200, 3)', 20.0),
}


def synthesize_stream(rec):
    """Kernel decision records -> the raw 32-bit words std::mt19937 would have had to produce."""
    kind = rec['w0'] & 3
    words = []
    for k, w0, w3 in zip(kind.tolist(), rec['w0'].tolist(), rec['w3'].tolist()):
        if k == 0:
            words.append(w3)                       # leaf = word % n_leaves (optimizer.hpp:103)
        elif k == 1:
            if w0 & 32:
                words.append(1 if (w0 & 4) else 0)  # coin = word % 2; 1 -> D = children[0]
            if w0 & 8:                              # width gate passed: the uniform is drawn (two words)
                words += [0, 0] if (w0 & 16) else [0xFFFFFFFF, 0xFFFFFFFF]   # u = 0 accepts, u = 1 - 2^-53 rejects
    return np.array(words, np.uint32)


def run_and_replay(net, max_width, n_sweeps, chains=3, tile=None, every=10, beta1=100.0, seed0=77):
    from tnco_b200 import networks
    from tnco_b200.engine import Engine, pack_leaf_bits, random_trees
    ts, ni = eval('networks.' + net)
    lb = pack_leaf_bits(ts, ni)
    n = lb.shape[0]
    seeds = np.arange(chains, dtype=np.uint64) + seed0
    P, A, B = random_trees(lb, ni, seeds)
    betas = np.array([beta1 * s / n_sweeps for s in range(n_sweeps)])
    if tile:
        os.environ['TNB_TILE'] = str(tile)
    try:
        e = Engine()
        e.set_network(lb, ni).set_mode(max_width=max_width, update_slices_every=every)
        e.set_chains(P, A, B, seeds)
    finally:
        os.environ.pop('TNB_TILE', None)
    if tile:
        assert e.config()['tile'] == tile
    e.set_trace(chains, cap_records=n_sweeps * (n + 2), cap_reslices=n_sweeps // every + 2)
    e.set_betas(betas)
    init_slices = e.slices() if max_width is not None else None
    init_bits = [e.bits(c) for c in range(chains)]
    # two launches: the trace and the running total must carry over
    e.run(n_sweeps // 3)
    e.run(n_sweeps)
    T, M = e.costs()
    gp, ga, gb = e.trees()
    bp, ba, bb = e.trees(best=True)
    stats = dict(proposals=0, gated=0, near=0, reslices=0, kept=0)
    pr = e.progress()
    for c in range(chains):
        rec, cand = e.trace(c)
        kind = rec['w0'] & 3
        prop = rec[kind == 1]
        rs = rec[kind == 2]
        assert (kind == 0).sum() == n_sweeps
        # the counters (proposals come from the generator's event index, width rejects from the gate passes)
        assert pr['sweeps'][c] == n_sweeps and pr['proposals'][c] == len(prop)
        assert pr['accepts'][c] == int(((prop['w0'] >> 4) & 1).sum())
        assert pr['width_rejects'][c] == int((((prop['w0'] >> 3) & 1) == 0).sum())
        oc = so.Chain(P[c], A[c], B[c], init_bits[c], ni, max_width=max_width, seed=0,
                      init_slices=None if max_width is None else init_slices[c])
        words = synthesize_stream(rec)
        oc.set_replay(words)
        oc.trace(len(prop) + 8)
        if max_width is not None:
            oc.set_forced_slices(cand, rs['w3'].astype(np.uint8))
        oc.run(betas, update_slices_every=every if max_width is not None else 0)
        used, over = oc.replay_state()
        assert not over and used == len(words), (used, len(words))
        # ---- state after the run
        for x, y in zip((gp[c], ga[c], gb[c]), oc.tree()):
            assert (x == y).all()
        assert (e.bits(c) == oc.bits()).all()
        occ, opc = oc.costs()
        assert (e.node_costs(c) == occ).all()
        assert abs(T[c] - oc.total_cost) <= 1e-9 * oc.total_cost
        assert abs(M[c] - oc.min_total_cost) <= 1e-9 * oc.min_total_cost
        for x, y in zip((bp[c], ba[c], bb[c]), oc.tree(best=True)):
            assert (x == y).all()
        if max_width is not None:
            assert (e.slices()[c] == oc.slices()).all() and (e.slices(best=True)[c] == oc.slices(best=True)).all()
        # ---- every proposal
        ot = oc.traced()
        assert len(ot) == len(prop)
        w0 = prop['w0']
        assert ((w0 >> 16) == ot['B']).all() and (prop['w3'] == ot['A']).all()
        assert (((w0 >> 2) & 1) == ot['pick0']).all()
        assert (((w0 >> 3) & 1) == ot['gate']).all()
        assert (((w0 >> 5) & 1) == ot['coin']).all()
        assert (((w0 >> 4) & 1) == ot['acc']).all()      # (forced through u; a refused downhill move would show here)
        g = ot['gate'] == 1
        assert (prop['d0'][g] == ot['delta'][g]).all()   # bit-exact deltas
        assert (np.abs(prop['d1'] - ot['total']) <= 1e-9 * ot['total']).all()
        # ---- the acceptance rule itself, in fp64, on the recorded numbers
        sweep_of = np.cumsum(kind == 0)[kind == 1] - 1
        beta = betas[sweep_of][g]
        nl2u = (prop['w1'][g]).view(np.float32).astype(np.float64)
        u = np.exp2(-nl2u)
        delta, total = ot['delta'][g], ot['total'][g]
        with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
            p = np.where(delta <= 0, 1.0, np.power(1.0 + delta / total, -beta))
        ref_acc = u <= p
        ker_acc = ot['acc'][g] == 1
        near = np.abs(u - p) <= 2e-5 * p
        bad = (ref_acc != ker_acc) & ~near
        assert not bad.any(), (int(bad.sum()), delta[bad][:4], total[bad][:4], u[bad][:4], p[bad][:4], beta[bad][:4])
        assert (ker_acc[delta <= 0]).all()
        stats['proposals'] += len(prop)
        stats['gated'] += int(g.sum())
        stats['near'] += int((near & (delta > 0)).sum())
        # ---- re-slices
        if max_width is not None:
            k, log = oc.reslice_log()
            assert k == len(rs) == len(cand)
            keep = rs['w3'] == 1
            own = log[:, 0] < log[:, 1]
            tie = np.abs(log[:, 0] - log[:, 1]) <= 1e-12 * log[:, 1]
            assert ((keep == own) | tie).all()
            assert (np.abs(rs['d1'] - log[:, 1]) <= 1e-9 * log[:, 1]).all()
            ch = ~(cand == 0).all(axis=1) & (rs['d0'] != rs['d1'])
            assert (np.abs(rs['d0'][ch] - log[ch, 0]) <= 1e-9 * log[ch, 0]).all()
            stats['reslices'] += k
            stats['kept'] += int(keep.sum())
    e.close()
    assert stats['near'] <= max(2, 1e-3 * stats['gated']), stats
    return stats


@pytest.mark.parametrize('name,n_sweeps', [('C1', 3000), ('C2', 1500), ('C3', 800), ('C4', 600), ('C4w24', 400),
                                           ('reg200_fw', 600)])
def test_production_kernel_decisions_replay_bit_exact_through_the_oracle(name, n_sweeps):
    net, mw = NETS[name]
    st = run_and_replay(net, mw, n_sweeps)
    assert st['proposals'] > 5 * n_sweeps
    if mw is not None:
        assert st['reslices'] > 0


@pytest.mark.parametrize('tile', [4, 8, 16, 32])
def test_production_kernel_replay_every_tile_shape(tile):
    run_and_replay('regular_graph(64, 0)', None, 800, chains=9, tile=tile)


@pytest.mark.parametrize('tile', [16, 32])
def test_production_kernel_replay_finite_width_sub_warp(tile):
    run_and_replay('regular_graph(150, 5)', 12.0, 500, chains=5, tile=tile)


def test_production_kernel_replay_low_beta_and_every_sweep_reslice():
    # slow ramp: many uphill acceptances, totals wander (running-total guard), re-slice after every sweep
    run_and_replay('grid_rqc(5, 5, 10)', 14.0, 700, chains=4, every=1, beta1=8.0)
