#!/usr/bin/env python3
"""Benchmark of the SA hot path (BASELINE.json metric: SA proposals/sec).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation, host cores

Workload (config.workload): BASELINE.json configs[1] -- 2D-grid 6x6 random circuit depth 12 tensor network
(180 tensors, 324 indices, bond dim 2), unconstrained SA, betas 0 -> 100, 4096 chains per GPU.
One STEP = one full anneal of the whole batch: `--sweeps` leaf->root sweeps of every chain from fresh initial
trees (built on the device, outside the timed region).  `value` = proposals/s with chain state resident in HBM, timed with CUDA events on the engine's stream
around the sweep kernel (L2 flushed before every step), max over ranks.  `e2e` = the same metric through the
public API `Optimizer(method='sa').optimize(...)` with host buffers: H2D of network / seeds / schedule, initial
trees and cache construction on the device, sweeps, D2H of the best costs and trees, result objects, wall clock.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the default (the configuration `metric` is quoted on that fits one GPU)
    'C2': dict(text='C2: 2D-grid 6x6 random circuit depth 12 TN (180 tensors, 324 indices, d=2), unconstrained SA',
               make=lambda nw: nw.grid_rqc(6, 6, 12), max_width=None),
    # BASELINE.json configs[3] / north_star target: Sycamore-53 m=20, memory-constrained (max width 2^32)
    'C4': dict(text='C4: Sycamore-style 53-qubit m=20 TN (430 tensors, 807 indices, d=2), memory-constrained SA, '
                    'max_width=32, update_slices=10', make=lambda nw: nw.sycamore(20), max_width=32.0),
}
WORKLOAD = WORKLOADS['C2']['text']
_SEL = {'name': 'C2'}


# ------------------------------------------------------------------------------------------ helpers
def measured_peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[])
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def workload():
    from tnco_b200 import networks
    from tnco_b200.engine import pack_leaf_bits
    ts, ni = WORKLOADS[_SEL['name']]['make'](networks)
    return ts, ni, pack_leaf_bits(ts, ni)


def index_rows(ts, ni):
    rows = [[2] for _ in range(ni)]
    for t, xs in enumerate(ts):
        for x in xs:
            rows[x].append(f't{t}')
    return rows


def bytes_per_proposal(W, levels_per_sweep, p_acc):
    """Algorithmic bytes per proposal, SURVEY.md 8(d): one new sibling bitset per level, the leaf pair once
    per sweep, the write-back of inds[B] on accept, plus 64 B of cost / topology scalars."""
    return 4.0 * W * (1.0 + 2.0 / max(levels_per_sweep, 1e-9) + p_acc) + 64.0


# ------------------------------------------------------------------------------------------ CPU reference arm
def _ref_worker(args):
    """One run of the reference, driven exactly like `core_` (tnco/app/infinite_memory/sa.py:199-209)."""
    kind, P, A, B, nb, ni, seed, n_sweeps, count, mw = args
    import numpy as np  # noqa
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    status = np.zeros(1)
    log2c = np.zeros(1, np.float32)
    if kind == 'reference':
        from helpers import RefChain
        rc = RefChain(P, A, B, nb, ni, seed=seed, max_width=mw)
        opt, mh = rc.opt, rc.mh
        t0 = time.perf_counter()
        for n in range(n_sweeps):
            mh.beta = n * (100.0 / n_sweeps)
            if mw is None:
                opt.update(mh)
            else:
                opt.update(mh, update_slices=(n % 10 == 0))   # finite_width/sa.py:228
            status[0] = n / n_sweeps
            log2c[0] = opt.log2_min_total_cost
        dt = time.perf_counter() - t0
        best = opt.log2_min_total_cost
    else:
        from oracle import sa_oracle as so
        oc = so.Chain(P, A, B, nb, ni, seed=seed, max_width=mw)
        t0 = time.perf_counter()
        oc.run([n * (100.0 / n_sweeps) for n in range(n_sweeps)], update_slices_every=10)
        dt = time.perf_counter() - t0
        best = oc.log2_min_total_cost
    props = 0
    if count:  # exact proposal count from the bit-identical restatement (untimed)
        from oracle import sa_oracle as so
        oc = so.Chain(P, A, B, nb, ni, seed=seed, max_width=mw)
        oc.run([n * (100.0 / n_sweeps) for n in range(n_sweeps)], update_slices_every=10)
        props = oc.counters()['proposals']
    return dt, props, best


def cpu_reference_rate(n_sweeps, n_runs=None, repeats=1):
    """proposals/s of the reference CPU SA on all host cores (joblib loky, one run per core)."""
    from joblib import Parallel, delayed
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import ref_core
    from tnco_b200.engine import random_trees
    kind = 'reference' if ref_core() is not None else 'port'
    ts, ni, lb = workload()
    cores = os.cpu_count() or 1
    n_runs = n_runs or cores
    seeds = np.arange(n_runs, dtype=np.uint64) + 1
    P, A, B = random_trees(lb, ni, seeds)
    n = lb.shape[0]
    nbs = []
    for k in range(n_runs):
        nb = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
        nb[:n] = lb
        for z in range(n, 2 * n - 1):
            nb[z] = nb[A[k][z]] ^ nb[B[k][z]]
        nbs.append(nb)
    out = []
    with Parallel(n_jobs=cores, backend='loky') as par:
        par(delayed(_ref_worker)((kind, P[k], A[k], B[k], nbs[k], ni, int(seeds[k]), 50, False, WORKLOADS[_SEL['name']]['max_width']))
            for k in range(n_runs))  # pool warm-up
        for rep in range(repeats):
            t0 = time.perf_counter()
            res = par(delayed(_ref_worker)((kind, P[k], A[k], B[k], nbs[k], ni, int(seeds[k]), n_sweeps, True, WORKLOADS[_SEL['name']]['max_width']))
                      for k in range(n_runs))
            wall = time.perf_counter() - t0
            props = sum(r[1] for r in res)
            in_loop = max(r[0] for r in res)
            out.append(dict(wall_s=wall, in_loop_s=in_loop, proposals=props, best_log2=min(r[2] for r in res)))
    return kind, cores, n_runs, out


def calibrate_ref_sweeps(target_s):
    kind, cores, n_runs, out = cpu_reference_rate(2000, repeats=1)
    rate = 2000 / max(out[0]['in_loop_s'], 1e-6)
    return max(2000, int(rate * target_s))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_sweeps = calibrate_ref_sweeps(4.0)
    kind, cores, n_runs, out = cpu_reference_rate(n_sweeps, repeats=args.warmup + args.steps)
    timed = out[args.warmup:]
    props = sum(o['proposals'] for o in timed)
    secs = sum(o['in_loop_s'] for o in timed)
    value = props / secs
    line = dict(metric='SA proposals/sec', value=value, unit='proposals/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload=WORKLOADS[_SEL['name']]['text'], chains=n_runs, sweeps_per_step=n_sweeps, betas=[0, 100]),
                cpu_baseline=dict(value=value, unit='proposals/s', cores=cores, kind=kind,
                                  sample=f'{n_runs} runs (one per core, joblib loky) x {n_sweeps} sweeps per step, '
                                         'in-loop time of the slowest run'),
                e2e=dict(value=value, unit='proposals/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                best_log2_flops=min(o['best_log2'] for o in timed))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from tnco_b200 import dist as tdist
    from tnco_b200.app import Optimizer
    from tnco_b200.engine import Engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ts, ni, lb = workload()
    n, W = lb.shape[0], lb.shape[1]
    C, S = args.chains, args.sweeps
    betas = np.array([n_ * (100.0 / S) for n_ in range(S)])
    seeds = (np.arange(C, dtype=np.uint64) + 1) + np.uint64(rank * C)

    eng = Engine(local)
    mw = WORKLOADS[_SEL['name']]['max_width']
    eng.set_network(lb, ni).set_mode(max_width=mw)
    eng.set_betas(betas)
    cfg = eng.config()

    def step():
        eng.generate_chains(seeds, chain_id0=rank * C)   # fresh initial trees, built on the device
        eng.costs()          # forces cache construction (init kernel) before the timed region
        eng.flush_l2()
        eng.timing()
        tdist.barrier()
        torch.cuda.synchronize()
        eng.run(S)           # the sweep kernel; timed inside with CUDA events on the engine's stream
        torch.cuda.synchronize()
        ms, nl = eng.timing()
        return ms, nl, eng.counters()

    for _ in range(args.warmup):
        step()
    tot_ms, launches, props, accs, sweeps = 0.0, 0, 0, 0, 0
    best = float('inf')
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            ms, nl, c = step()
            tot_ms += ms
            launches += nl
            props += c['proposals']
            accs += c['accepts']
            sweeps += c['sweeps']
            best = min(best, float(np.log2(eng.costs()[1]).min()))
    ms_max = tdist.all_reduce_max(tot_ms)
    props_all = tdist.all_reduce_sum(props)
    # the path's one exchange step: min-reduce of the best cost + broadcast of the winning tree (NCCL)
    t, m = eng.costs()
    k = int(np.argmin(m))
    bp, ba, bb = eng.trees(best=True, chain0=k, n=1)
    gbest, _, owner = tdist.global_best(float(m[k]), np.concatenate([bp[0], ba[0], bb[0]]))
    eng.close()

    # ---- e2e through the public API, host buffers in, result objects out
    rows = index_rows(ts, ni)
    e2e_props, e2e_s, e2e_parts = 0, 0.0, {}
    for i in range(args.e2e_warmup + args.e2e_steps):
        opt = Optimizer(method='sa', seed=1000 + i, max_width=mw)
        tdist.barrier()
        t0 = time.perf_counter()
        tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=S, n_runs=C * world)
        dt = tdist.all_reduce_max(time.perf_counter() - t0)
        if i >= args.e2e_warmup:
            e2e_props += tdist.all_reduce_sum(opt.last_stats['proposals'])
            e2e_s += dt
            e2e_parts = {k: round(1e3 * opt.last_stats.get(k, 0.0), 1) for k in ('engine_s', 'exchange_s', 'assemble_s')}
            e2e_parts['kernel_ms'] = round(opt.last_stats['kernel_ms'], 1)
            e2e_parts['wall_ms'] = round(1e3 * dt, 1)
    N = 2 * n - 1
    npad, ws = (N + 7) // 8 * 8, (W + 3) // 4 * 4
    h2d = C * 8 + S * 8 + n * ws * 4 + (ni + 1) * 8 + 2 * ni * 2   # seeds, betas, network (trees are built on the device)
    d2h = C * (npad * 2 + (n - 1) * 4 + 2 * 8 + 3 * 8)

    L = props / max(sweeps, 1)
    pacc = accs / max(props, 1)
    bpp = bytes_per_proposal(W, L, pacc)
    per_rank_rate = props / (tot_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = per_rank_rate * bpp / 1e9
    traffic, issue = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'sweep_kernel_traffic.json')))
        traffic = prof['dram_bytes_per_launch'] if _SEL['name'] == 'C2' else None
        # the roof that actually binds this L2-resident workload: warp-instruction issue (4 schedulers x 148 SMs,
        # one instruction per cycle each); instructions per proposal from the committed ncu capture
        ipp = prof['warp_instructions_per_proposal']
    except Exception:
        ipp = None
    line = dict(metric='SA proposals/sec', value=props_all / (ms_max * 1e-3), unit='proposals/s', n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms_max / args.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=WORKLOADS[_SEL['name']]['text'], chains_per_gpu=C, sweeps_per_step=S, betas=[0, 100],
                            rng='philox4x32-10', l2='flushed before every step (256 MiB memset)',
                            tile=cfg['tile'], words_per_lane=cfg['words_per_lane'],
                            state_bytes_per_chain=cfg['state_bytes_per_chain'], parallelism=f'chains sharded x{world}'),
                clocks=clk.summary(),
                e2e=dict(value=e2e_props / max(e2e_s, 1e-9), unit='proposals/s', h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h, api="Optimizer(method='sa').optimize(rows, betas=(0,100), fuse=False, n_steps, n_runs)",
                         last_step_ms=e2e_parts),
                gpu_launches=launches,
                roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                              traffic=traffic, kernel='sa_sweep_kernel', peak_source=peak_src,
                              bytes_per_proposal=bpp, levels_per_sweep=L, accept_ratio=pacc,
                              note=('chain state (%d MB) is L2-resident; the kernel is latency/issue bound, see DESIGN.md'
                                    if C * cfg['state_bytes_per_chain'] <= (96 << 20) else
                                    'chain state (%d MB) is HBM-resident (split layout), latency bound, see DESIGN.md')
                              % (C * cfg['state_bytes_per_chain'] // 1000000)),
                best_log2_flops=gbest and float(np.log2(gbest)), proposals_per_step=props / args.steps)
    if ipp and args.workload == 'C2':
        mhz = line['clocks'].get('sm_mhz') or line['clocks'].get('sm_max_mhz') or 1965.0
        peak_issue = 148 * 4 * mhz * 1e6
        line['roofline']['issue'] = dict(bound='warp-instruction issue', warp_instr_per_proposal=ipp,
                                         achieved=per_rank_rate * ipp / 1e9, peak=peak_issue / 1e9, unit='Ginstr/s',
                                         frac=per_rank_rate * ipp / peak_issue,
                                         source='instructions/proposal: ncu smsp__inst_executed.sum of the same kernel '
                                                '(profiles/sweep_kernel_traffic.json); rate and clock: this run')
    if world == 1 and not args.no_cpu_baseline:
        n_sw = calibrate_ref_sweeps(12.0)
        kind, cores, n_runs, out = cpu_reference_rate(n_sw, repeats=1)
        o = out[0]
        line['cpu_baseline'] = dict(value=o['proposals'] / o['in_loop_s'], unit='proposals/s', cores=cores, kind=kind,
                                    sample=f'{n_runs} runs (one per core, joblib loky) x {n_sw} sweeps of the same '
                                           f'network, in-loop time {o["in_loop_s"]:.1f} s (wall {o["wall_s"]:.1f} s)',
                                    best_log2_flops=o['best_log2'])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--chains', type=int, default=4096, help='chains per GPU')
    ap.add_argument('--sweeps', type=int, default=10000, help='sweeps per chain per step (n_steps of the anneal)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--e2e-warmup', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='C2', choices=sorted(WORKLOADS), help='C2 = BASELINE.json configs[1] (default)')
    args = ap.parse_args()
    _SEL['name'] = args.workload
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
