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


from bench import _cpu_worker, cpu_arm  # noqa: E402,F401  (defined there: bench.py times the same arm)


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
