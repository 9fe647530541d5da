"""Shared driver of ``Optimizer.optimize`` for both SA plugins (tnco/app/infinite_memory/sa.py:100-257 and
tnco/app/finite_width/sa.py:116-289): argument checks, beta ramp, seeds, per-component runs, result assembly.
The per-run loop of the reference (one process per run, one pybind call per sweep) is replaced by ONE batched
engine run over all ``n_runs`` chains on the GPU."""
from __future__ import annotations

import functools as fts
import operator as op
import time
from collections.abc import Sequence
from random import Random
from sys import stderr

import numpy as np

from .. import dist
from ..engine import (RNG_MT19937, RNG_PHILOX, cached_engine, TREES_GREEDY, TREES_RANDOM, merge_paths, pack_index_set,
                      pack_leaf_bits, random_trees, tree_to_path, unpack_bits)
from ..optimize.infinite_memory.cost_model import check_sparse
from ..tn import get_connected_components
from .app import cost_to_decimal, load_tn


def expand_betas(betas, n_steps):
    """tnco/app/infinite_memory/sa.py:139-156."""
    if n_steps is not None:
        if int(n_steps) != n_steps or n_steps <= 0:
            raise ValueError("'n_steps' must be a positive number.")
        n_steps = int(n_steps)
    if isinstance(betas, tuple) and len(betas) == 2:
        if n_steps is None:
            raise ValueError("'n_steps' must be provided if 'betas' has the format '(beta_min, beta_max)'.")
        if betas[0] == betas[1]:
            raise ValueError("'betas' must use the format '(beta_ini, beta_end)', with 'beta_ini != beta_end'.")
        step = (betas[1] - betas[0]) / n_steps  # more_itertools.numeric_range: start + n*step
        return np.array([betas[0] + n * step for n in range(n_steps)], np.float64)
    b = np.array(list(betas), np.float64)
    if n_steps is not None:
        b = b[:n_steps]
    if len(b) == 0:
        raise ValueError("'betas' is empty.")
    return b


def run_component(opt, comp, tn, imap, seeds, betas, *, finite, update_slices, deadline, stats, n_projs=None):
    """All runs of one connected component.  Returns per run: min cost (float), best tree -> path over all
    tensors of tn, slices (index names)."""
    all_ts, all_dims, outs = tn.ts_inds, tn.dims, tn.output_inds   # (properties that rebuild their value: read once)
    ts = [all_ts[t] for t in comp]
    inds = list(dict.fromkeys(x for xs in ts for x in xs))
    pos = {x: k for k, x in enumerate(inds)}
    dims = [int(all_dims[x]) for x in inds]
    uniform = all(d == dims[0] for d in dims)
    lb = pack_leaf_bits([[pos[x] for x in xs] for xs in ts], len(inds))
    out_bits = pack_index_set([pos[x] for x in inds if x in outs], len(inds))
    # sparse-index cost model (tnco/app/infinite_memory/sa.py:158-161): every component gets the network's sparse
    # indices (those it holds) and the same n_projs
    sparse = [pos[x] for x in inds if x in tn.sparse_inds] if tn.sparse_inds else []
    sp_bits = pack_index_set(sparse, len(inds)) if tn.sparse_inds else None
    rng_kind = RNG_MT19937 if opt.rng == 'mt19937' else RNG_PHILOX
    n_runs = len(seeds)
    multi = bool(opt.distributed) and dist.world()[1] > 1
    lo, hi = dist.shard(n_runs) if multi else (0, n_runs)
    my_seeds = np.asarray(seeds[lo:hi], np.uint64)
    n_local = hi - lo
    method = TREES_GREEDY if opt.init_trees == 'greedy' else TREES_RANDOM
    n_steps = len(betas)
    W32 = (len(inds) + 31) // 32
    n_int = len(ts) - 1
    dev = dist.local_device(opt.device)
    dist.set_device(dev)
    dist.set_shard_total(n_runs)
    t_eng = time.perf_counter()
    # a rank without runs (n_runs < world size) skips the engine and offers +inf / empty blocks to the collectives
    mins = np.zeros(0, np.float64)
    chw = np.zeros((0, n_int), np.uint32)
    sl = np.zeros((0, W32), np.uint32)
    eng = cached_engine(dev) if n_local > 0 else None
    try:
        if eng is not None:
            eng.set_network(lb, len(inds), dim=dims[0], dims=None if uniform else dims, output_bits=out_bits,
                            sparse_bits=sp_bits, n_projs=n_projs)
            eng.set_mode(max_width=opt.max_width if finite else None, update_slices_every=update_slices, rng=rng_kind)
            t0 = time.perf_counter()
            built = False
            if opt.tree_builder == 'device':
                try:
                    eng.generate_chains(my_seeds, chain_id0=lo, method=method)
                    built = True
                except ValueError as ex:   # (hyper-index networks with very few indices per tensor)
                    if 'not supported' not in str(ex):
                        raise
            if not built:  # host C++ threads
                P, A, B = random_trees(lb, len(inds), my_seeds, method=method, output_bits=out_bits)
                eng.set_chains(P, A, B, my_seeds, chain_id0=lo)
            stats['tree_gen_s'] += time.perf_counter() - t0
            eng.set_betas(betas)
        if multi and opt.sync_every:
            # periodic exchange: ONE min-reduction of the packed (best cost, chain id) key over ranks
            # + ONE broadcast of the winning tree every sync_every sweeps.  Reporting / early visibility only --
            # chains never read it, so run statistics stay the reference's.  Every rank makes the same number of
            # exchanges: the stop decision (deadline passed on ANY rank) is itself collective.
            done = 0
            while done < n_steps:
                late = 0.0 if deadline is None or time.perf_counter() < deadline else 1.0
                if dist.all_reduce_max(late) > 0.0:
                    break
                done = min(n_steps, done + int(opt.sync_every))
                if eng is not None:
                    eng.run(done)
                stats.setdefault('global_best_history', []).append((done, _exchange_best(eng, lo, finite, n_int, W32)[0]))
        elif eng is not None and opt.verbose >= 2:
            # Progress surface (reference: per-run `status` and `log2_total_cost` buffers shown by a progress bar when
            # verbose >= 2, tnco/parallel.py:229-317, infinite_memory/sa.py:208-209): the anneal runs in 20 pieces and
            # every piece reports the batch's status and best / mean log2 cost to stderr; the same records are kept
            # in stats['progress'].
            import sys
            t_run = time.perf_counter()
            for piece in range(1, 21):
                target = n_steps * piece // 20
                if target == 0:
                    continue
                eng.run(target, timeout_s=None if deadline is None else max(deadline - time.perf_counter(), 0.0))
                _, m_now = eng.costs()
                rec = dict(status=eng.reached / n_steps, sweeps=eng.reached, elapsed_s=time.perf_counter() - t_run,
                           log2_min_total_cost=float(np.log2(m_now.min())) if m_now.size and m_now.min() > 0 else float('nan'),
                           mean_log2_min_total_cost=float(np.mean(np.log2(np.maximum(m_now, 1e-300)))))
                stats.setdefault('progress', []).append(rec)
                print('[tnco_b200 rank %d] %5.1f %%  sweep %d/%d  best log2 cost %.4f  mean %.4f  %.1f s' % (
                    dist.world()[0], 100 * rec['status'], rec['sweeps'], n_steps, rec['log2_min_total_cost'],
                    rec['mean_log2_min_total_cost'], rec['elapsed_s']), file=sys.stderr, flush=True)
                if eng.reached < target:   # timed out
                    break
        elif eng is not None:
            # timeout: the reference polls a stop flag every sweep (sa.py:201); here the engine checks the wall clock
            # between internal launches (tnb_run_timed)
            eng.run(n_steps, timeout_s=None if deadline is None else max(deadline - time.perf_counter(), 0.0))
        if eng is not None:
            ms, nl = eng.timing()
            stats['kernel_ms'] += ms
            stats['launches'] += nl
            c = eng.counters()
            for k in ('proposals', 'accepts', 'sweeps'):
                stats[k] += c[k]
            _, mins = eng.costs()
            # best trees travel and wait in the engine's compact form: one word child0 | child1 << 16 per internal node
            chw = eng.trees_packed(best=True)
            sl = eng.slices(best=True) if finite else np.zeros((n_local, W32), np.uint32)
            stats['config'] = eng.config()
    except BaseException:
        if eng is not None:
            eng.close()  # do not hand a half-configured engine to the next call
        raise
    stats['engine_s'] = stats.get('engine_s', 0.0) + time.perf_counter() - t_eng
    t_x = time.perf_counter()
    owned = np.arange(lo, hi)
    have = {int(r): k for k, r in enumerate(owned)}    # global run id -> row of chw / sl
    if multi:
        # The one exchange step.  Trees stay on the rank that owns them; what travels is
        #   (1) the per-run minima, 8 B x n_runs (all_gather) -- every rank can order all runs;
        #   (2) the trees (+ slices) of the global top-k, filled in by their owners (one all_reduce over a
        #       [k][n_int + W32] block that is zero elsewhere); row 0 is the global best, i.e. min + broadcast in one.
        mins_all = dist.all_gather_rows(mins, n_runs)
        k = n_runs if opt.gather_paths == 'all' else min(int(opt.gather_paths), n_runs)
        top = np.argsort(mins_all, kind='stable')[:k]
        blk = np.zeros((k, n_int + W32), np.uint32)
        for j, r in enumerate(top.tolist()):
            if r in have:
                blk[j, :n_int] = chw[have[r]]
                blk[j, n_int:] = sl[have[r]]
        blk = dist.all_reduce_rows_sum(blk)
        rows = np.concatenate([np.arange(len(owned)), len(owned) + np.arange(k)])
        chw = np.concatenate([chw, blk[:, :n_int]], axis=0)
        sl = np.concatenate([sl, blk[:, n_int:]], axis=0)
        for j, r in enumerate(top.tolist()):
            have.setdefault(int(r), len(owned) + j)
        mins = mins_all
        stats['global_best'] = float(mins_all[top[0]]) if k else float('inf')
    stats['exchange_s'] = stats.get('exchange_s', 0.0) + time.perf_counter() - t_x
    return dict(mins=mins, chw=chw, slices=sl, have=have, inds=inds, comp=np.asarray(comp, np.int32),
                n_tensors=len(tn), whole=(len(comp) == len(tn)))


def _exchange_best(eng, lo, finite, n_int, W32):
    """Periodic exchange: packed-key min-reduction + broadcast of the winner's best tree (+ slices)."""
    if eng is None:
        return dist.global_best(float('inf'), 0, np.zeros(n_int + W32, np.int32))
    m = eng.costs()[1]
    kb = int(np.argmin(m))
    payload = np.concatenate([eng.trees_packed(best=True, chain0=kb, n=1)[0],
                              eng.slices(best=True, chain0=kb, n=1)[0] if finite else np.zeros(W32, np.uint32)])
    return dist.global_best(float(m[kb]), lo + kb, payload.view(np.int32))


def _pairs(a):
    """[k][2] int32 -> list of k 2-tuples (one C-level conversion)."""
    a = np.ascontiguousarray(a, np.int32)
    return a.view(np.dtype([('x', '<i4'), ('y', '<i4')])).reshape(a.shape[0]).tolist()


def _comp_path(pc, r):
    """Linear path of run r of one component over all tensors (ContractionTree.path(), ctree.py:350-388)."""
    n = len(pc['comp'])
    c0, c1 = np.full(2 * n - 1, -1, np.int32), np.full(2 * n - 1, -1, np.int32)
    row = _row(pc, r)
    c0[n:] = pc['chw'][row] & np.uint32(0xffff)
    c1[n:] = pc['chw'][row] >> np.uint32(16)
    return tree_to_path(c0, c1, n_tensors=pc['n_tensors'], tensors_pos=pc['comp'])


def _row(pc, r):
    """Row of run r's best tree on this rank; under torch.distributed only the rank's own runs and the gathered
    top-k (Optimizer(gather_paths=k | 'all')) are here."""
    try:
        return pc['have'][int(r)]
    except KeyError:
        raise RuntimeError(
            f'the contraction tree of run {r} lives on another rank: it is neither one of this rank\'s runs nor in '
            "the gathered top-k; construct the Optimizer with gather_paths='all' (or a larger k)") from None


def _comp_slices(pc, r):
    return frozenset(pc['inds'][i] for i in unpack_bits(pc['slices'][_row(pc, r)]))


class ResultList(Sequence):
    """``sorted(results)`` of the reference (sa.py:257) as a read-only sequence whose records come into being when
    they are indexed: ``optimize()`` orders all runs by cost but builds no per-run object up front (32768 records
    cost 12-64 ms of Python per call and rank).  Indexing, slicing (-> list), iteration, ``len``, ``==`` with lists."""

    def __init__(self, cls, field, order):
        self._cls, self._field, self._order, self._made = cls, field, np.asarray(order), {}

    def __len__(self):
        return len(self._order)

    def _one(self, i):
        rec = self._made.get(i)
        if rec is None:
            rec = self._made[i] = self._cls._from_source(self._field, int(self._order[i]))
        return rec

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(k) for k in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError('result index out of range')
        return self._one(i)

    def __eq__(self, other):
        return len(self) == len(other) and all(a == b for a, b in zip(self, other))

    def __repr__(self):
        return f'ResultList({len(self)} runs, best cost {self[0].cost if len(self) else None})'

    def __reduce__(self):
        return (list, (list(self),))


def optimize(opt, results_cls, tn, betas, n_steps, n_runs, n_projs, update_slices, timeout, finite,
             load_tn_options):
    rng = opt._rng
    if opt.distributed and dist.world()[1] > 1 and opt.seed is None:
        # seed=None draws from the OS on every process: ranks would fuse the network differently and pick different
        # run seeds.  Rank 0's draw is shared before anything random happens.
        shared = dist.broadcast_int(Random().getrandbits(62))
        rng = Random(shared)
        tn = load_tn(tn, atol=opt.atol, dtype=opt.dtype, backend=opt.backend, seed=shared, verbose=opt.verbose,
                     **load_tn_options)
    else:
        tn = opt._load_tn(tn, **load_tn_options)
    check_sparse(tn.sparse_inds, n_projs)  # the cost model's argument rules (sa.py:158-161 builds it first)
    betas = expand_betas(betas, n_steps)
    if int(n_runs) != n_runs or n_runs < 1:
        raise ValueError("'n_runs' must be a positive number.")
    seeds = rng.choices(range(2**32), k=int(n_runs))  # sa.py:237
    if opt.verbose == 1:
        print('# Optimizing ...', file=stderr, flush=True, end='')
    t_start = time.perf_counter()
    deadline = None if timeout is None else t_start + float(timeout)
    stats = dict(tree_gen_s=0.0, kernel_ms=0.0, launches=0, proposals=0, accepts=0, sweeps=0)
    comps = get_connected_components(tn.ts_inds)
    per_comp = []
    for comp in comps:
        if len(comp) < 2:  # trivial path (sa.py:179-183)
            per_comp.append(None)
            continue
        per_comp.append(run_component(opt, comp, tn, None, seeds, betas, finite=finite,
                                      update_slices=update_slices, deadline=deadline, stats=stats,
                                      n_projs=n_projs))
    runtime = time.perf_counter() - t_start
    R = int(n_runs)
    live = [pc for pc in per_comp if pc is not None]
    # total cost per run = sum of the 6-significant-digit Decimals the reference prints (sa.py:215-220)
    if len(live) == 1:
        # the printed cost is a monotone function of the exact one: ordering by the exact cost orders the Decimals
        # (equal Decimals come out in order of their exact costs instead of run order)
        keys = np.asarray(live[0]['mins'], np.float64)
    else:
        keys = np.array([float(sum(cost_to_decimal(pc['mins'][r]) for pc in live)) for r in range(R)]) if live \
            else np.zeros(R)
    order = np.argsort(keys, kind='stable')  # == sorted(results) on cost (app.py:83-87, sa.py:257)

    def field(name, r):
        """One field of the result record of run r, computed when first looked at."""
        if name == 'runtime_s':
            return runtime
        if name == 'disconnected_costs':
            return [0 if pc is None else cost_to_decimal(pc['mins'][r]) for pc in per_comp]
        if name == 'cost':
            return sum(field('disconnected_costs', r))
        if name == 'disconnected_paths':
            return [[] if pc is None else _pairs(_comp_path(pc, r)) for pc in per_comp]
        if name == 'path':
            if len(live) == 1 and live[0]['whole']:  # one component spanning the network: its path, pairs sorted
                return _pairs(np.sort(_comp_path(live[0], r), axis=1))
            cat = np.concatenate([_comp_path(pc, r) for pc in live], axis=0)[None] if live else \
                np.zeros((1, 0, 2), np.int32)
            # tn_utils.merge_contraction_paths (sa.py:230), in C++
            return _pairs(merge_paths(len(tn), [len(pc['comp']) - 1 for pc in live], cat)[0])
        if name == 'disconnected_slices':
            return [frozenset() if pc is None else _comp_slices(pc, r) for pc in per_comp]
        if name == 'slices':
            return fts.reduce(op.or_, field('disconnected_slices', r), frozenset())
        raise AttributeError(name)

    # One record per run, tens of thousands per call (x ranks): the list creates a record when it is first looked at.
    results = ResultList(results_cls, field, order)
    stats['assemble_s'] = time.perf_counter() - t_start - runtime
    if opt.verbose == 1:
        print(' Done!', file=stderr, flush=True)
    stats['runtime_s'] = runtime
    object.__setattr__(opt, 'last_stats', stats)
    return opt._dump_results(tn, results)
