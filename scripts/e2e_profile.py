"""Where the end-to-end time of Optimizer.optimize goes (dev tool).  argv: cfg n_runs n_steps [max_width]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnco_b200 import networks  # noqa: E402
from tnco_b200.app import Optimizer  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
n_runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
mw = float(sys.argv[4]) if len(sys.argv) > 4 else None
ts, ni = networks.CONFIGS[cfg]['make']()
rows = [[2] for _ in range(ni)]
for t, xs in enumerate(ts):
    for x in xs:
        rows[x].append(f't{t}')
for rep in range(3):
    opt = Optimizer(method='sa', seed=rep, max_width=mw)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=n_steps, n_runs=n_runs)
    pr.disable()
    dt = time.perf_counter() - t0
    st = opt.last_stats
    print(f'rep {rep}: wall {dt * 1e3:.1f} ms, kernel {st["kernel_ms"]:.1f} ms, trees+init {st["tree_gen_s"] * 1e3:.1f} ms, '
          f'proposals {st["proposals"]}, e2e rate {st["proposals"] / dt:.3e}, best cost {res[0].cost}')
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
