#!/usr/bin/env python3
"""Benchmark of the SA hot path (BASELINE.json metric: SA proposals/sec; best log2 FLOPs at a fixed 60 s).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation, host cores

Workload (config.workload): **C4**, BASELINE.json configs[3] and the north_star target -- Sycamore-style 53-qubit m=20
tensor network (430 tensors, 807 indices, bond dim 2), memory-constrained SA (max_width = 32, re-slicing every 10th
sweep), betas 0 -> 100, 4096 chains per GPU.  (Round 1 benchmarked C2; lines of the two rounds are not comparable.
C1 / C2 / C3 / C5 are measured in the same run and reported under "configs".)
One STEP = one full anneal of the whole batch: `--sweeps` leaf->root sweeps of every chain from fresh initial trees
(built on the device, outside the timed region).  `value` = proposals/s with chain state resident in HBM, timed with
CUDA events on the engine's stream around the sweep kernel (L2 flushed before every step), max over ranks.
`e2e` = the same metric through the public API `Optimizer(method='sa').optimize(...)` with host buffers: H2D of
network / seeds / schedule, initial trees and cache construction on the device, sweeps, D2H of the best costs and
trees, the sorted result list, the linear contraction paths of the 32 best runs and the JSON of the best one, wall
clock.  `roofline`: the larger of (algorithmic integer lane-operations/s) / (measured LOP3+POPC peak) and
(algorithmic bytes/s) / (measured bandwidth of the level the chain state lives in) -- peaks from
profiles/r02_int_peaks.json (scripts/peaks_microbench.cu) and MEASURED_PEAKS.json; `roofline.issue` is the
warp-instruction-issue utilisation (instructions per proposal from the committed ncu captures).
`best_log2_at_60s`: one beta 0 -> 100 anneal calibrated to fill `--anneal-budget` seconds of wall clock on the GPU(s)
(tree construction, sweeps, read-back included) next to the reference CPU SA given the same wall clock on all host
cores (N = 1 only; the reference arm reports its own).
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
    # chains / sweeps: per GPU and per step of the headline run (C4) or of the short per-config runs (the others)
    'C1': dict(text='C1: random 3-regular graph TN, 64 tensors (96 indices, d=2), unconstrained SA',
               make=lambda nw: nw.regular_graph(64, 0), max_width=None, chains=32768, sweeps=10000),
    'C2': dict(text='C2: 2D-grid 6x6 random circuit depth 12 TN (180 tensors, 324 indices, d=2), unconstrained SA',
               make=lambda nw: nw.grid_rqc(6, 6, 12), max_width=None, chains=4096, sweeps=10000),
    'C3': dict(text='C3: Sycamore-style 53-qubit m=14 TN (301 tensors, 549 indices, d=2), unconstrained SA',
               make=lambda nw: nw.sycamore(14), max_width=None, chains=12288, sweeps=4000),
    # BASELINE.json configs[3] / north_star target: Sycamore-53 m=20, memory-constrained (max width 2^32)
    'C4': dict(text='C4: Sycamore-style 53-qubit m=20 TN (430 tensors, 807 indices, d=2), memory-constrained SA, '
                    'max_width=32, update_slices=10', make=lambda nw: nw.sycamore(20), max_width=32.0, chains=4096,
               sweeps=10000),
    'C5': dict(text='C5: random 3-regular graph TN, 1000 tensors (1500 indices, d=2), unconstrained SA, HBM-resident '
                    'chain state', make=lambda nw: nw.regular_graph(1000, 0), max_width=None,
               # (its kernel holds 64 registers: 148 SMs x 32 resident single-warp blocks = 4736 chains fill one wave;
               #  4096 leave an eighth of the slots empty and run 6 % slower, profiles/r02_experiments.md)
               chains=4736, sweeps=2000),
}
_SEL = {'name': 'C4'}


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


def workload(name=None):
    from tnco_b200 import networks
    from tnco_b200.engine import pack_leaf_bits
    ts, ni = WORKLOADS[name or _SEL['name']]['make'](networks)
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
    if args.anneal_budget > 0:   # BASELINE metric part (ii) for this arm: one anneal filling the wall-clock budget
        ts, ni, lb = workload()
        c = cpu_arm(lb, ni, WORKLOADS[_SEL['name']]['max_width'], args.anneal_budget)
        line['best_log2_at_60s'] = dict(budget_s=args.anneal_budget, cpu_reference=c['best_log2_flops'], cpu_detail=c)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ fixed wall clock, CPU arm
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



# ------------------------------------------------------------------------------------------ roofline
def load_peaks():
    """Measured denominators: profiles/r02_int_peaks.json (scripts/peaks_microbench.cu on this pool's B200s) and the
    driver-written MEASURED_PEAKS.json (HBM copy bandwidth)."""
    hbm, hbm_src = measured_peak_gbs()
    try:
        pk = json.load(open(os.path.join(ROOT, 'profiles', 'r02_int_peaks.json')))
        src = 'measured (profiles/r02_int_peaks.json)'
    except Exception:
        pk = dict(popc=dict(lane_ops_per_clk_per_sm=16.0), lop3=dict(lane_ops_per_clk_per_sm=64.0),
                  l2_hit_ld128=dict(gb_per_s=17000.0), sms=148)
        src = 'fallback (B200_PROFILING.md: 16 POPC, 64 LOP3 lanes/clk/SM)'
    return dict(hbm_gbs=hbm, hbm_src=hbm_src, popc=pk['popc']['lane_ops_per_clk_per_sm'],
                lop3=pk['lop3']['lane_ops_per_clk_per_sm'], l2_gbs=pk['l2_hit_ld128']['gb_per_s'], sms=pk.get('sms', 148),
                src=src)


def roofline_of(name, W, finite, rate, L, pacc, state_bytes, mhz, proposals_per_launch):
    """per proposal 7*W integer lane-operations (5 logic + 2 POPC per word; finite width 11*W = 8 + 3)
    and 4*W*(1 + 2/L + p_acc) + 64 algorithmic bytes.  Reports the larger fraction as the bound."""
    pk = load_peaks()
    logic, popc = (8, 3) if finite else (5, 2)
    clk = (mhz or 1965.0) * 1e6
    # lane-operations/s the SMs can retire for this mix: (logic + popc) / (logic / LOP3 rate + popc / POPC rate)
    int_peak = (logic + popc) / (logic / pk['lop3'] + popc / pk['popc']) * pk['sms'] * clk
    int_ach = rate * (logic + popc) * W
    bpp = bytes_per_proposal(W, L, pacc)
    resident = 'l2' if state_bytes <= (96 << 20) else 'hbm'
    mem_peak = pk['l2_gbs'] if resident == 'l2' else pk['hbm_gbs']
    mem_ach = rate * bpp / 1e9
    i_frac, m_frac = int_ach / int_peak, mem_ach / mem_peak
    out = dict(kernel='sa_sweep_kernel', bytes_per_proposal=bpp, int_lane_ops_per_proposal=(logic + popc) * W,
               levels_per_sweep=L, accept_ratio=pacc, chain_state_bytes=int(state_bytes), residency=resident,
               int=dict(achieved=int_ach / 1e9, peak=int_peak / 1e9, unit='Glane-op/s', frac=i_frac,
                        mix=f'{logic} LOP3 + {popc} POPC per word', peak_source=pk['src']),
               mem=dict(achieved=mem_ach, peak=mem_peak, unit='GB/s', frac=m_frac,
                        peak_source=(pk['src'] + ' l2_hit_ld128') if resident == 'l2' else pk['hbm_src']))
    if m_frac >= i_frac:
        out.update(bound='hbm' if resident == 'hbm' else 'l2', achieved=mem_ach, peak=mem_peak, unit='GB/s', frac=m_frac,
                   peak_source=out['mem']['peak_source'])
    else:
        out.update(bound='int', achieved=int_ach / 1e9, peak=int_peak / 1e9, unit='Glane-op/s', frac=i_frac,
                   peak_source=pk['src'])
    traffic = None
    try:  # per-launch DRAM traffic and instruction counts of the same kernel from the committed ncu captures
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'r02_sweep_kernel_ncu.json')))[name]
        traffic = prof['dram_bytes_per_proposal'] * proposals_per_launch
        ipp = prof['warp_instructions_per_proposal']
        peak_issue = pk['sms'] * 4 * clk
        out['issue'] = dict(bound='warp-instruction issue (utilisation, not a roofline)', warp_instr_per_proposal=ipp,
                            achieved=rate * ipp / 1e9, peak=peak_issue / 1e9, unit='Ginstr/s', frac=rate * ipp / peak_issue,
                            source='ncu smsp__inst_executed.sum (profiles/r02_sweep_kernel_ncu.json); rate, clock: this run')
    except Exception:
        pass
    out['traffic'] = traffic
    return out


# ------------------------------------------------------------------------------------------ our arm
def index_rows_of(name):
    ts, ni, lb = workload(name)
    return index_rows(ts, ni), ts, ni, lb


def measure_kernel(name, eng_dev, rank, world, chains, sweeps, steps, warmup, clock=None):
    """Device-timed proposals/s of the sweep kernel on one workload (this rank's share)."""
    import torch
    from tnco_b200 import dist as tdist
    from tnco_b200.engine import Engine
    ts, ni, lb = workload(name)
    mw = WORKLOADS[name]['max_width']
    betas = np.array([n_ * (100.0 / sweeps) for n_ in range(sweeps)])
    seeds = (np.arange(chains, dtype=np.uint64) + 1) + np.uint64(rank * chains)
    eng = Engine(eng_dev)
    eng.set_network(lb, ni).set_mode(max_width=mw)
    eng.set_betas(betas)

    def step():
        eng.generate_chains(seeds, chain_id0=rank * chains)   # fresh initial trees, built on the device
        eng.costs()          # forces cache construction (init kernel) before the timed region
        eng.flush_l2()
        eng.timing()
        tdist.barrier()
        torch.cuda.synchronize()
        eng.run(sweeps)      # the sweep kernel; timed inside with CUDA events on the engine's stream
        torch.cuda.synchronize()
        ms, nl = eng.timing()
        return ms, nl, eng.counters()

    for _ in range(warmup):
        step()
    tot_ms, launches, props, accs, swp = 0.0, 0, 0, 0, 0
    best = float('inf')
    ctx = clock if clock is not None else _Null()
    with ctx:
        for _ in range(steps):
            ms, nl, c = step()
            tot_ms += ms
            launches += nl
            props += c['proposals']
            accs += c['accepts']
            swp += c['sweeps']
            best = min(best, float(eng.costs()[1].min()))
    cfg = eng.config()
    eng.close()
    ms_max = tdist.all_reduce_max(tot_ms)
    props_all = tdist.all_reduce_sum(props)
    best_all = -tdist.all_reduce_max(-best)
    return dict(value=props_all / (ms_max * 1e-3), ms_per_step=ms_max / steps, launches=launches, proposals=props,
                rank_rate=props / (tot_ms * 1e-3), L=props / max(swp, 1), pacc=accs / max(props, 1), cfg=cfg,
                best_log2=float(np.log2(best_all)), W=lb.shape[1], n=lb.shape[0], ni=ni, steps=steps)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


E2E_API = ("Optimizer(method='sa', max_width).optimize(rows, betas=(0,100), fuse=False, n_steps, n_runs); "
           "[r.path for r in res[:32]]; res[0].to_json()")


def measure_e2e(name, world, chains, sweeps, steps, warmup):
    """The same metric through the public API with host buffers, including what the reference's runs end with: the
    sorted result list, materialised linear paths (of the 32 best runs) and the JSON record of the best."""
    from tnco_b200 import dist as tdist
    from tnco_b200.app import Optimizer
    rows, ts, ni, lb = index_rows_of(name)
    mw = WORKLOADS[name]['max_width']
    n, W = lb.shape
    props, secs, parts = 0, 0.0, {}
    for i in range(warmup + steps):
        opt = Optimizer(method='sa', seed=1000 + i, max_width=mw)
        tdist.barrier()
        t0 = time.perf_counter()
        tn, res = opt.optimize(rows, betas=(0, 100), fuse=False, decompose_hyper_inds=False, n_steps=sweeps,
                               n_runs=chains * world)
        paths = [r.path for r in res[:32]]
        js = res[0].to_json()
        dt_local = time.perf_counter() - t0
        assert len(paths[0]) == n - 1 and len(js) > 10
        dt = tdist.all_reduce_max(dt_local)
        if i >= warmup:
            props += tdist.all_reduce_sum(opt.last_stats['proposals'])
            secs += dt
            parts = {k: round(1e3 * opt.last_stats.get(k, 0.0), 1) for k in ('engine_s', 'exchange_s', 'assemble_s')}
            parts['kernel_ms'] = round(opt.last_stats['kernel_ms'], 1)
            parts['wall_ms'] = round(1e3 * dt, 1)
    N = 2 * n - 1
    npad, ws = (N + 7) // 8 * 8, (W + 3) // 4 * 4
    h2d = chains * 8 + sweeps * 12 + n * ws * 4 + (ni + 1) * 8 + 2 * ni * 2   # seeds, betas (+1/beta), network
    d2h = chains * ((n - 1) * 4 + 8 + (ws * 4 if mw is not None else 0) + 3 * 8)  # packed best trees, minima, slices, counters
    return dict(value=props / max(secs, 1e-9), unit='proposals/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                api=E2E_API, last_step_ms=parts)


def anneal_gpu(name, dev, rank, world, chains, budget):
    """One beta 0 -> 100 anneal calibrated to fill `budget` seconds of wall clock (tree construction on the device, cache
    construction, sweeps, read-back of costs and the winning tree included; CUDA context creation excluded)."""
    from tnco_b200 import dist as tdist
    from tnco_b200.engine import Engine
    ts, ni, lb = workload(name)
    mw = WORKLOADS[name]['max_width']
    seeds = (np.arange(chains, dtype=np.uint64) + 1) + np.uint64(rank * chains)
    e = Engine(dev)
    e.set_network(lb, ni).set_mode(max_width=mw)
    n2 = 300
    for target_s in (None, 3.0):   # two calibration passes: sweeps get cheaper as the trees improve
        e.generate_chains(seeds, chain_id0=rank * chains)
        e.set_betas(np.linspace(0, 100, n2, endpoint=False))
        e.timing()
        e.run(n2)
        ms, _ = e.timing()
        rate = n2 / (ms * 1e-3)
        n2 = max(300, int(rate * 3.0))
    n_sweeps = max(1000, int(rate * budget * 0.95))
    n_sweeps = int(-tdist.all_reduce_max(-float(n_sweeps)))   # the slowest rank's calibration, so all ranks run alike
    tdist.barrier()
    t0 = time.perf_counter()
    e.generate_chains(seeds, chain_id0=rank * chains)
    e.set_betas(np.array([k * (100.0 / n_sweeps) for k in range(n_sweeps)]))
    e.run(n_sweeps, timeout_s=budget - (time.perf_counter() - t0) - 0.05)
    t, m = e.costs()
    k = int(np.argmin(m))
    bp, ba, bb = e.trees(best=True, chain0=k, n=1)
    sl = e.slices(best=True, chain0=k, n=1) if mw is not None else None
    wall = time.perf_counter() - t0
    c = e.counters()
    seq, pc, w = e.eval_cost(bp, ba, bb, slices=sl)   # independent re-evaluation of the winner
    done = e.reached
    e.close()
    best_all = -tdist.all_reduce_max(-float(m[k]))
    wall_all = tdist.all_reduce_max(wall)
    props_all = tdist.all_reduce_sum(c['proposals'])
    return dict(best_log2_flops=float(np.log2(best_all)), mean_best_log2_flops=float(np.log2(m).mean()), wall_s=round(wall_all, 2),
                chains_per_gpu=chains, n_sweeps=n_sweeps, sweeps_done=int(done), proposals=props_all,
                recomputed_log2_flops_rank0=float(np.log2(seq[0])), max_width_rank0=float(w[0]))


def anneal_cpu(name, budget):
    ts, ni, lb = workload(name)
    return cpu_arm(lb, ni, WORKLOADS[name]['max_width'], budget)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from tnco_b200 import dist as tdist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    name = args.workload
    C = args.chains or WORKLOADS[name]['chains']
    S = args.sweeps or WORKLOADS[name]['sweeps']
    clk = ClockSampler(local)
    m = measure_kernel(name, local, rank, world, C, S, args.steps, args.warmup, clock=clk)
    clocks = clk.summary()
    mhz = clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or 1965.0
    e2e = measure_e2e(name, world, C, S, args.e2e_steps, args.e2e_warmup) if args.e2e_steps > 0 else None
    finite = WORKLOADS[name]['max_width'] is not None
    cfg = m['cfg']
    line = dict(metric='SA proposals/sec', value=m['value'], unit='proposals/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=m['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload=WORKLOADS[name]['text'], chains_per_gpu=C, sweeps_per_step=S, betas=[0, 100],
                            rng='philox4x32-10', l2='flushed before every step (256 MiB memset)', tile=cfg['tile'],
                            words_per_lane=cfg['words_per_lane'], state_bytes_per_chain=cfg['state_bytes_per_chain'],
                            parallelism=f'chains sharded x{world}',
                            comparability='round 1 benchmarked C2 (now under "configs"); this line is C4'),
                clocks=clocks, e2e=e2e, gpu_launches=m['launches'],
                roofline=roofline_of(name, m['W'], finite, m['rank_rate'], m['L'], m['pacc'],
                                     C * cfg['state_bytes_per_chain'], mhz, m['proposals'] / m['steps']),
                best_log2_flops=m['best_log2'], proposals_per_step=m['proposals'] / m['steps'])
    # ---- the other BASELINE configs, short runs of the same two measurements
    if not args.no_configs:
        line['configs'] = {}
        for other in ('C1', 'C2', 'C3', 'C5'):
            if other == name:
                continue
            oc, osw = WORKLOADS[other]['chains'], WORKLOADS[other]['sweeps']
            om = measure_kernel(other, local, rank, world, oc, osw, 2, 3)
            oe = measure_e2e(other, world, oc, osw, 1, 1)
            line['configs'][other] = dict(
                workload=WORKLOADS[other]['text'], chains_per_gpu=oc, sweeps_per_step=osw, value=om['value'],
                ms_per_step=om['ms_per_step'], e2e=oe['value'], tile=om['cfg']['tile'],
                roofline=roofline_of(other, om['W'], False, om['rank_rate'], om['L'], om['pacc'],
                                     oc * om['cfg']['state_bytes_per_chain'], mhz, om['proposals'] / om['steps']),
                best_log2_flops=om['best_log2'])
    # ---- BASELINE metric part (ii): best log2 FLOPs at a fixed wall clock
    if args.anneal_budget > 0:
        g = anneal_gpu(name, local, rank, world, C, args.anneal_budget)
        line['best_log2_at_60s'] = dict(budget_s=args.anneal_budget, n_gpus=world, gpu=g['best_log2_flops'], gpu_detail=g,
                                        cpu_reference=None)
    if world == 1 and not args.no_cpu_baseline:
        n_sw = calibrate_ref_sweeps(12.0)
        kind, cores, n_runs, out = cpu_reference_rate(n_sw, repeats=1)
        o = out[0]
        line['cpu_baseline'] = dict(value=o['proposals'] / o['in_loop_s'], unit='proposals/s', cores=cores, kind=kind,
                                    sample=f'{n_runs} runs (one per core, joblib loky) x {n_sw} sweeps of the same '
                                           f'network, in-loop time {o["in_loop_s"]:.1f} s (wall {o["wall_s"]:.1f} s)',
                                    best_log2_flops=o['best_log2'])
        if args.anneal_budget > 0:
            c = anneal_cpu(name, args.anneal_budget)
            line['best_log2_at_60s']['cpu_reference'] = c['best_log2_flops']
            line['best_log2_at_60s']['cpu_detail'] = c
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
    ap.add_argument('--chains', type=int, default=0, help='chains per GPU (default: the workload\'s, 4096 for C4)')
    ap.add_argument('--sweeps', type=int, default=0, help='sweeps per chain per step (default: the workload\'s, 10^4 for C4)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--e2e-warmup', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the short C1/C2/C3/C5 runs')
    ap.add_argument('--anneal-budget', type=float, default=60.0, help='seconds of the fixed-wall-clock anneal (0 = skip)')
    ap.add_argument('--workload', default='C4', choices=sorted(WORKLOADS), help='C4 = north_star target (default)')
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
