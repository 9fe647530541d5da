#!/usr/bin/env python3
"""Best log2 FLOPs reached in a fixed wall-clock budget: our GPU engine vs the reference CPU SA on the host cores
(BASELINE.json metric part 2).

    python scripts/anneal60.py CFG [--budget 60] [--max-width W] [--chains N]

Both arms anneal beta 0 -> 100 once, with n_steps calibrated so that the anneal fills the budget.
GPU arm: wall clock includes tree construction on the device, cache construction, all sweeps and the read-back of
the best costs / trees (one-time CUDA context creation is excluded and reported).  CPU arm: the compiled reference
core (oracle/_ref) driven like tnco/app/*/sa.py `core_`, one run per host core in joblib-loky processes.
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def _cpu_worker(args):
    P, A, B, nb, ni, seed, n_sweeps, mw, every, budget = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import RefChain
    rc = RefChain(P, A, B, nb, ni, seed=seed, max_width=mw)
    opt, mh = rc.opt, rc.mh
    t0 = time.perf_counter()
    done = 0
    for n in range(n_sweeps):
        mh.beta = n * (100.0 / n_sweeps)
        if mw is None:
            opt.update(mh)
        else:
            opt.update(mh, update_slices=(n % every == 0))
        done = n + 1
        if (n & 255) == 0 and time.perf_counter() - t0 > budget:  # the reference's timeout flag (parallel.py:243-248)
            break
    return time.perf_counter() - t0, done, opt.log2_min_total_cost


def cpu_arm(lb, ni, mw, budget, every=10):
    from joblib import Parallel, delayed
    from tnco_b200.engine import random_trees
    cores = os.cpu_count() or 1
    seeds = np.arange(cores, dtype=np.uint64) + 1
    P, A, B = random_trees(lb, ni, seeds)
    n = lb.shape[0]
    nbs = []
    for k in range(cores):
        nb = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
        nb[:n] = lb
        for z in range(n, 2 * n - 1):
            nb[z] = nb[A[k][z]] ^ nb[B[k][z]]
        nbs.append(nb)
    with Parallel(n_jobs=cores, backend='loky') as par:
        cal = par(delayed(_cpu_worker)((P[k], A[k], B[k], nbs[k], ni, int(seeds[k]), 3000, mw, every, 1e9))
                  for k in range(cores))
        rate = 3000 / max(c[0] for c in cal)   # sweeps/s of the slowest run with every core busy
        # second pass: a whole beta ramp of ~5 s (sweeps get cheaper as the trees improve, a short ramp underestimates)
        n2 = max(3000, int(rate * 5.0))
        cal = par(delayed(_cpu_worker)((P[k], A[k], B[k], nbs[k], ni, int(seeds[k]), n2, mw, every, 1e9))
                  for k in range(cores))
        rate = n2 / max(c[0] for c in cal)
        n_sweeps = max(1000, int(rate * budget * 0.97))
        t0 = time.perf_counter()
        res = par(delayed(_cpu_worker)((P[k], A[k], B[k], nbs[k], ni, int(seeds[k]), n_sweeps, mw, every, budget))
                  for k in range(cores))
        wall = time.perf_counter() - t0
    return dict(cores=cores, runs=cores, n_sweeps=n_sweeps, sweeps_done=[r[1] for r in res], wall_s=round(wall, 2),
                in_loop_s=round(max(r[0] for r in res), 2), best_log2_flops=min(r[2] for r in res),
                mean_best_log2_flops=float(np.mean([r[2] for r in res])))


def gpu_arm(lb, ni, mw, budget, n_chains, every=10):
    from tnco_b200.engine import Engine
    seeds = np.arange(n_chains, dtype=np.uint64) + 1
    t0 = time.perf_counter()
    e = Engine()
    ctx_s = time.perf_counter() - t0
    e.set_network(lb, ni).set_mode(max_width=mw, update_slices_every=every)
    # calibration run (short anneal, same batch size)
    e.generate_chains(seeds)
    e.set_betas(np.linspace(0, 100, 300, endpoint=False))
    e.run(300)
    ms, _ = e.timing()
    n2 = max(300, int(300 / (ms * 1e-3) * 3.0))   # second pass: a whole beta ramp of ~3 s
    e.generate_chains(seeds)
    e.set_betas(np.linspace(0, 100, n2, endpoint=False))
    e.timing()
    e.run(n2)
    ms, _ = e.timing()
    n_sweeps = max(1000, int(n2 / (ms * 1e-3) * budget * 0.96))
    # the timed anneal
    t0 = time.perf_counter()
    e.generate_chains(seeds)
    e.set_betas(np.array([k * (100.0 / n_sweeps) for k in range(n_sweeps)]))
    done, chunk = 0, max(1, n_sweeps // 16)
    while done < n_sweeps and time.perf_counter() - t0 < budget:
        done = min(n_sweeps, done + chunk)
        e.run(done)
    t, m = e.costs()
    k = int(np.argmin(m))
    bp, ba, bb = e.trees(best=True, chain0=k, n=1)
    sl = e.slices(best=True, chain0=k, n=1) if mw is not None else None
    wall = time.perf_counter() - t0
    c = e.counters()
    # independent check of the winner: full-tree evaluation of the returned tree (+ slices)
    seq, pc, w = e.eval_cost(bp, ba, bb, slices=sl)
    kms, _ = e.timing()
    out = dict(chains=n_chains, n_sweeps=n_sweeps, sweeps_done=done, wall_s=round(wall, 2), context_s=round(ctx_s, 2),
               kernel_s=round(kms * 1e-3, 2), proposals=c['proposals'], proposals_per_s=c['proposals'] / wall,
               best_log2_flops=float(np.log2(m.min())), mean_best_log2_flops=float(np.log2(m).mean()),
               best_recomputed_log2_flops=float(np.log2(seq[0])), best_max_width=float(w[0]), config=e.config())
    e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('cfg')
    ap.add_argument('--budget', type=float, default=60.0)
    ap.add_argument('--max-width', type=float, default=None)
    ap.add_argument('--chains', type=int, default=4096)
    ap.add_argument('--skip-cpu', action='store_true')
    a = ap.parse_args()
    from tnco_b200 import networks
    from tnco_b200.engine import pack_leaf_bits
    ts, ni = networks.CONFIGS[a.cfg]['make']()
    lb = pack_leaf_bits(ts, ni)
    out = dict(cfg=a.cfg, tensors=len(ts), indices=ni, max_width=a.max_width, budget_s=a.budget)
    out['gpu'] = gpu_arm(lb, ni, a.max_width, a.budget, a.chains)
    if not a.skip_cpu:
        out['cpu_reference'] = cpu_arm(lb, ni, a.max_width, a.budget)
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
