"""The N>1 path on CPU: world_size=2 over gloo.  Chains are sharded with no data-path collective; the only
exchange is the min-reduction of the best cost + broadcast of the winning tree and the gather of per-run
results (tnco_b200/dist.py).  The engine is the kernel-logic emulation (tests/emu) because this box has no GPU."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

WORKER = r'''
import ctypes, json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np
import torch.distributed as dist
from tnco_b200 import _lib
_lib._LIB = _lib.bind(ctypes.CDLL(os.path.join({root!r}, 'tests', 'emu', 'libtnb_emu.so')))
from tnco_b200 import dist as tdist
from tnco_b200.app import Optimizer
from helpers import regular_network

world = int(os.environ.get('WORLD_SIZE', '1'))
if world > 1:
    dist.init_process_group('gloo')
rank = tdist.world()[0]
assert tdist.shard(10, 0, 3) == (0, 3) and tdist.shard(10, 2, 3) == (6, 10)
# global_best: ONE packed-key min-reduction + ONE broadcast.  Rank r owns chains [2r, 2r+2) of 2*world and offers
# cost 10-r for its chain 2r+1 with payload [r]*4
tdist.set_shard_total(2 * world)
best, payload, owner, chain = tdist.global_best(10.0 - rank, 2 * rank + 1, np.full(4, rank, np.int32))
assert chain == 2 * (world - 1) + 1
assert tdist.pack_key(3.0, 5) < tdist.pack_key(3.5, 1) and tdist.pack_key(3.0, 1) < tdist.pack_key(3.0, 5)
assert tdist.pack_key(float('inf'), 0) > tdist.pack_key(1e300, 2**24 - 1)
rows_all = tdist.all_gather_rows(np.arange(*tdist.shard(7)).reshape(-1, 1), 7)
ts, ni = regular_network(20, 3)
rows = [[2] for _ in range(ni)]
for t, xs in enumerate(ts):
    for x in xs:
        rows[x].append('t%d' % t)
opt = Optimizer(seed=9, max_width={mw}, sync_every=20, gather_paths='all')
tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=60, n_runs=6)
# top-k gathering: costs of all runs everywhere, trees of the k best + this rank's own runs only
opt_k = Optimizer(seed=9, max_width={mw}, gather_paths=2)
_, res_k = opt_k.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=60, n_runs=6)
topk = dict(costs=[str(r.cost) for r in res_k], paths=[], missing=0)
for r in res_k:
    try:
        topk['paths'].append(r.path)
    except RuntimeError as ex:
        assert 'another rank' in str(ex)
        topk['paths'].append(None)
        topk['missing'] += 1
# fewer runs than ranks: the rank without runs must neither crash nor hang the collectives
_, res_1 = Optimizer(seed=3, max_width={mw}).optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False,
                                                     n_steps=30, n_runs=1)
one = dict(cost=str(res_1[0].cost), path=res_1[0].path, n=len(res_1))
# seed=None: every rank must end up with the same (fused) network and the same run seeds
tn_n, res_n = Optimizer(seed=None, max_width={mw}).optimize(rows, betas=(0, 100), n_steps=20, n_runs=4)
noseed = dict(n_tensors=len(tn_n), costs=[str(r.cost) for r in res_n], best_path=res_n[0].path)
# timeout together with the periodic exchange: the stop decision is collective (no mismatched collectives / hang)
opt_t = Optimizer(seed=5, max_width={mw}, sync_every=5)
_, res_t = opt_t.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=200000, n_runs=4,
                          timeout=0.5 + 0.4 * rank)
timed = dict(n=len(res_t), exchanges=len(opt_t.last_stats.get('global_best_history', [])))
out = dict(rank=rank, world=world, best=best, payload=payload.tolist(), owner=owner, gathered=rows_all.ravel().tolist(),
           costs=[str(r.cost) for r in res], paths=[r.path for r in res],
           slices=[sorted(r.slices) for r in res] if {mw} is not None else None,
           local_sweeps=opt.last_stats['sweeps'], history=opt.last_stats.get('global_best_history'),
           topk=topk, one=one, noseed=noseed, timed=timed)
open(os.path.join({outdir!r}, 'out_%d_%d.json' % (world, rank)), 'w').write(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
'''


def run(world, mw, tmp_path):
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'tests', 'emu')], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    script = tmp_path / f'worker_{world}_{mw}.py'
    script.write_text(WORKER.format(root=ROOT, mw=mw, outdir=str(tmp_path)))
    env = dict(os.environ, OMP_NUM_THREADS='1')
    if world == 1:
        cmd = [sys.executable, str(script)]
    else:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 2000 + (7 if mw else 0)),
               str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads((tmp_path / f'out_{world}_{r}.json').read_text()) for r in range(world)]


@pytest.mark.parametrize('mw', [None, 8])
def test_two_ranks_over_gloo_match_single_process(mw, tmp_path):
    single = run(1, mw, tmp_path)[0]
    two = sorted(run(2, mw, tmp_path), key=lambda r: r['rank'])
    assert [r['rank'] for r in two] == [0, 1] and all(r['world'] == 2 for r in two)
    for r in two:
        assert r['best'] == 9.0 and r['owner'] == 1 and r['payload'] == [1, 1, 1, 1]   # min over ranks, owner's tree
        assert r['gathered'] == list(range(7))
        # results do not depend on the number of ranks (seeds and Philox counters use global chain ids)
        assert r['costs'] == single['costs'] and r['paths'] == single['paths'] and r['slices'] == single['slices']
        # periodic min-reduction: every 20 sweeps, same value on both ranks, non-increasing
        assert [h[0] for h in r['history']] == [20, 40, 60] and r['history'] == two[0]['history']
        assert all(a[1] >= b[1] for a, b in zip(r['history'], r['history'][1:]))
    assert two[0]['local_sweeps'] + two[1]['local_sweeps'] == single['local_sweeps'] == 6 * 60
    # gather_paths=2: all costs on both ranks; the two best trees everywhere; the rest only where they ran
    for r in two:
        assert r['topk']['costs'] == single['costs']
        assert r['topk']['paths'][:2] == single['paths'][:2]
        assert all(p is None or p == q for p, q in zip(r['topk']['paths'], single['paths']))
    assert single['topk']['missing'] == 0 and two[0]['topk']['missing'] + two[1]['topk']['missing'] == 4
    # n_runs = 1 < world size
    assert all(r['one'] == single['one'] for r in two) and single['one']['n'] == 1
    # seed=None: ranks agree with each other (not with the single process, which drew its own seed)
    assert two[0]['noseed'] == two[1]['noseed']
    # collective stop
    assert two[0]['timed']['exchanges'] == two[1]['timed']['exchanges'] >= 1 and two[0]['timed']['n'] == 4
