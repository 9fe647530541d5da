#!/usr/bin/env python3
"""Quick throughput probe of the sweep kernel on every BASELINE config (dev tool, not the bench)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnco_b200 import networks  # noqa: E402
if os.environ.get('TNB_LIB'):   # A/B builds of the library (dev tool only)
    from tnco_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ['TNB_LIB'])
from tnco_b200.engine import Engine, pack_leaf_bits, random_trees  # noqa: E402


def probe(cfg, n_chains, n_sweeps, max_width=None, tile=None, trees='greedy', layout=0):
    ts, ni = networks.CONFIGS[cfg]['make']()
    lb = pack_leaf_bits(ts, ni)
    seeds = np.arange(n_chains, dtype=np.uint64) + 1
    t0 = time.time()
    n_init = min(n_chains, 512)
    p, a, b = random_trees(lb, ni, seeds[:n_init], method=0 if trees == 'greedy' else 1)
    reps = (n_chains + n_init - 1) // n_init
    p, a, b = (np.tile(x, (reps, 1))[:n_chains] for x in (p, a, b))
    t_trees = time.time() - t0
    if tile:
        os.environ['TNB_TILE'] = str(tile)
    e = Engine()
    e.set_network(lb, ni)
    e.set_mode(max_width=max_width, layout=layout, update_slices_every=int(os.environ.get('TNB_EVERY', '10')))
    e.set_chains(p, a, b, seeds)
    os.environ.pop('TNB_TILE', None)   # (the tile shape is fixed when the chains are created)
    e.set_betas(np.linspace(0, 100, n_sweeps, endpoint=False))
    t, m = e.costs()
    init_log2 = float(np.log2(t).mean())
    e.timing()
    e.run(n_sweeps // 10)  # warm-up part of the anneal
    ms0, _ = e.timing()
    c0 = e.counters()
    e.run(n_sweeps)
    ms, nl = e.timing()
    c1 = e.counters()
    t, m = e.costs()
    props = c1['proposals'] - c0['proposals']
    out = dict(cfg=cfg, n_chains=n_chains, n_sweeps=n_sweeps, max_width=max_width, **e.config(),
               tree_gen_s=round(t_trees, 3), init_log2=round(init_log2, 2),
               best_log2=round(float(np.log2(m).min()), 3), mean_best_log2=round(float(np.log2(m).mean()), 3),
               ms=round(ms, 2), proposals=props, proposals_per_s=props / (ms * 1e-3),
               sweeps_per_s=(c1['sweeps'] - c0['sweeps']) / (ms * 1e-3),
               accept=round((c1['accepts'] - c0['accepts']) / max(props, 1), 3),
               levels_per_sweep=round(props / max(c1['sweeps'] - c0['sweeps'], 1), 2))
    print(json.dumps(out), flush=True)
    e.close()
    return out


if __name__ == '__main__':
    which = sys.argv[1:] or ['C1', 'C2', 'C3', 'C4', 'C5']
    for cfg in which:
        if cfg == 'SMEM':   # shared-memory-resident layout (3) against the in-place layout (1), small networks
            for tile in (4, 8, 16, 32):
                for layout in (1, 3):
                    probe('C1', 32768, 2000, tile=tile, layout=layout)
            for tile in (16, 32):
                for layout in (1, 3):
                    probe('C2', 4096, 2000, tile=tile, layout=layout)
            for layout in (1, 3):
                probe('C2', 16384, 1000, tile=16, layout=layout)
        if cfg == 'C1':
            probe('C1', 32768, 2000)
            probe('C1', 32768, 2000, tile=32)
        if cfg == 'C2':
            probe('C2', 4096, 2000)
            probe('C2', 4096, 2000, tile=32)
            probe('C2', 16384, 1000)
        if cfg == 'C3':
            probe('C3', 8192, 1000)
        if cfg == 'C4':
            probe('C4', 4096, 500, max_width=32)
        if cfg.startswith('C4x'):   # C4 with another chain count, e.g. C4x3552
            probe('C4', int(cfg[3:]), 500, max_width=32)
        if cfg == 'C5':
            probe('C5', 4096, 500)
