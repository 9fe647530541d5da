"""Thin numpy-facing wrapper over the C-ABI (include/tnco_b200.h).  One Engine == one GPU.

Host code only marshals buffers; all optimisation work happens in the CUDA library.
"""
from __future__ import annotations

import atexit
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import (LAYOUT_AUTO, PROB_MH, RNG_MT19937, RNG_PHILOX, RNG_REPLAY, TREES_GREEDY,  # noqa: F401
                   TREES_RANDOM)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class EngineError(RuntimeError):
    pass


def pack_leaf_bits(ts_inds, n_inds):
    """Per-tensor lists of index positions -> [n][W32] uint32 bitsets (bit i -> word i//32, bit i%32)."""
    W = (int(n_inds) + 31) // 32
    out = np.zeros((len(ts_inds), W), np.uint32)
    for t, xs in enumerate(ts_inds):
        for x in xs:
            out[t, x >> 5] |= np.uint32(1 << (x & 31))
    return out


def unpack_bits(row):
    """uint32 bitset row -> sorted list of positions."""
    return [w * 32 + b for w, v in enumerate(np.asarray(row).tolist()) for b in range(32) if (v >> b) & 1]


# ------------------------------------------------------------------------------------------ host helpers
def pack_index_set(positions, n_inds):
    """Index positions -> [W32] uint32 bitset."""
    out = np.zeros((int(n_inds) + 31) // 32, np.uint32)
    for x in positions:
        out[x >> 5] |= np.uint32(1 << (x & 31))
    return out


def random_trees(leaf_bits, n_inds, seeds, method=TREES_GREEDY, n_threads=0, output_bits=None):
    """Initial contraction trees, one per seed: (parent, child0, child1), each [n_trees][2n-1] int32.
    ``output_bits``: [W32] bitset of the open indices (only matters for networks with hyper-indices)."""
    L = _lib.lib()
    lb = _c(leaf_bits, np.uint32)
    n = lb.shape[0]
    seeds = _c(seeds, np.uint64)
    T, N = len(seeds), 2 * n - 1
    p, a, b = (np.empty((T, N), np.int32) for _ in range(3))
    ob = None if output_bits is None else _c(output_bits, np.uint32)
    rc = L.tnb_random_trees_out(n, int(n_inds), _ptr(lb, C.c_uint32), _ptr(ob, C.c_uint32), T,
                                _ptr(seeds, C.c_uint64), int(method), int(n_threads), _ptr(p, C.c_int32),
                                _ptr(a, C.c_int32), _ptr(b, C.c_int32))
    if rc:
        raise ValueError(L.tnb_last_error(None).decode())
    return p, a, b


def tree_to_path(child0, child1, n_tensors=None, tensors_pos=None):
    """Tree(s) -> linear (einsum) path(s), [n_trees][n-1][2] (ContractionTree.path(), tnco/ctree.py:350-388).
    Leaf k is tensor ``tensors_pos[k]`` of a network of ``n_tensors`` tensors (default: identity)."""
    L = _lib.lib()
    tp = None if tensors_pos is None else _c(tensors_pos, np.int32)
    a, b = _c(child0, np.int32), _c(child1, np.int32)
    single = a.ndim == 1
    a2, b2 = np.atleast_2d(a), np.atleast_2d(b)
    T, N = a2.shape
    n = (N + 1) // 2
    out = np.empty((T, max(n - 1, 0), 2), np.int32)
    rc = L.tnb_tree_to_path(n, T, _ptr(a2, C.c_int32), _ptr(b2, C.c_int32), int(n_tensors or n),
                            _ptr(tp, C.c_int32), _ptr(out, C.c_int32))
    if rc:
        raise ValueError(L.tnb_last_error(None).decode())
    return out[0] if single else out


def merge_paths(n_tensors, lens, paths):
    """Batched merge_contraction_paths (tnco/utils/tn.py:334-401): ``paths`` is [n_runs][sum(lens)][2] with the
    per-component paths of a run concatenated; returns [n_runs][n_tensors-1][2]."""
    L = _lib.lib()
    lens = _c(lens, np.int32).reshape(-1)
    pth = _c(paths, np.int32).reshape(-1, int(lens.sum()), 2) if len(lens) else np.zeros((len(paths), 0, 2), np.int32)
    out = np.empty((pth.shape[0], max(int(n_tensors) - 1, 0), 2), np.int32)
    rc = L.tnb_merge_paths(int(n_tensors), pth.shape[0], len(lens), _ptr(lens, C.c_int32), _ptr(pth, C.c_int32),
                           _ptr(out, C.c_int32))
    if rc:
        raise ValueError(L.tnb_last_error(None).decode())
    return out


def path_to_tree(path, n_leaves):
    """Linear path -> (parent, child0, child1) in reference numbering (tnco/ctree.py:108-131,208-218)."""
    L = _lib.lib()
    pth = _c(path, np.int32).reshape(-1, 2)
    n = int(n_leaves)
    if len(pth) != n - 1:
        raise ValueError('a full path needs n_leaves-1 contractions')
    p, a, b = (np.empty(2 * n - 1, np.int32) for _ in range(3))
    rc = L.tnb_path_to_tree(n, _ptr(pth, C.c_int32), _ptr(p, C.c_int32), _ptr(a, C.c_int32), _ptr(b, C.c_int32))
    if rc:
        raise ValueError(L.tnb_last_error(None).decode())
    return p, a, b


def mt19937_stream(seed, n):
    out = np.empty(int(n), np.uint32)
    _lib.lib().tnb_mt19937_stream(int(seed) & 0xFFFFFFFF, int(n), _ptr(out, C.c_uint32))
    return out


def mt19937_state_str(seed, n_draws):
    """libstdc++ text form of std::mt19937(seed) after n_draws outputs == reference ``Optimizer.prng_state``."""
    st = np.empty(624, np.uint32)
    pos = C.c_int32(0)
    _lib.lib().tnb_mt19937_state(int(seed) & 0xFFFFFFFF, int(n_draws), _ptr(st, C.c_uint32), C.byref(pos))
    return ' '.join(map(str, st.tolist())) + ' ' + str(pos.value)


def parse_mt19937_state(text):
    """libstdc++ text form of a std::mt19937 (624 words, then the position) -> uint32[625]."""
    try:
        w = [int(x) for x in str(text).split()]
    except ValueError:
        w = []
    if len(w) != 625 or w[624] > 624 or any(x < 0 or x >= 2**32 for x in w):
        raise ValueError("'seed' is not a valid std::mt19937 state.")
    return np.array(w, np.uint32)


def mt19937_advance_str(state625, n_draws):
    """Text form of the generator `state625` after n_draws further outputs."""
    st = np.array(state625[:624], np.uint32)
    pos = C.c_int32(int(state625[624]))
    _lib.lib().tnb_mt19937_advance(_ptr(st, C.c_uint32), C.byref(pos), int(n_draws))
    return ' '.join(map(str, st.tolist())) + ' ' + str(pos.value)


# ------------------------------------------------------------------------------------------ engine
_ENGINES = {}


def cached_engine(device=0):
    """One long-lived Engine per device for back-to-back ``optimize()`` calls: creating / destroying the CUDA stream
    and (re)allocating device memory per call costs tens to hundreds of milliseconds.  The engine keeps its device
    memory until ``release_cached_engines()`` (also run at interpreter exit)."""
    e = _ENGINES.get(int(device))
    if e is not None and e._L is not _lib.lib():  # the bound library changed (tests switch to the emulation build)
        e.close()
        e = None
    if e is None or not getattr(e, '_h', None):
        e = _ENGINES[int(device)] = Engine(device)
    return e


def release_cached_engines():
    for e in list(_ENGINES.values()):
        e.close()
    _ENGINES.clear()


class Engine:
    """Batched SA chains on one B200."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.tnb_create(C.byref(h), int(device))
        if rc:
            raise EngineError(self._L.tnb_last_error(None).decode())
        self._h = h
        self.n = self.N = self.W = self.n_chains = 0

    def close(self):
        if getattr(self, '_h', None):
            self._L.tnb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            msg = self._L.tnb_last_error(self._h).decode()
            if any(k in msg for k in ('Precision is too low', 'invalid', 'not supported', 'not connected')):
                raise ValueError(msg)
            raise EngineError(msg)

    def set_network(self, leaf_bits, n_inds, dim=2, dims=None, output_bits=None, sparse_bits=None, n_projs=None):
        lb = _c(leaf_bits, np.uint32)
        self.n, self.n_inds = lb.shape[0], int(n_inds)
        self.N, self.W = 2 * self.n - 1, (self.n_inds + 31) // 32
        if lb.shape != (self.n, self.W):
            raise ValueError('leaf_bits must be [n_leaves][ceil(n_inds/32)]')
        d = None if dims is None else _c(dims, np.uint64)
        self._chk(self._L.tnb_set_network(self._h, self.n, self.n_inds, _ptr(lb, C.c_uint32), int(dim),
                                          _ptr(d, C.c_uint64)))
        if output_bits is not None:
            ob = _c(output_bits, np.uint32).reshape(-1)
            if ob.shape != (self.W,):
                raise ValueError('output_bits must be [ceil(n_inds/32)]')
            self._chk(self._L.tnb_set_output_inds(self._h, _ptr(ob, C.c_uint32)))
        if sparse_bits is not None:
            self.set_sparse_inds(sparse_bits, n_projs)
        return self

    def set_skip_slices(self, skip_bits):
        """Indices the greedy slicer never takes (core-object ``skip_slices``); ``None`` clears."""
        sb = None if skip_bits is None else _c(skip_bits, np.uint32).reshape(-1)
        if sb is not None and sb.shape != (self.W,):
            raise ValueError('skip_bits must be [ceil(n_inds/32)]')
        self._chk(self._L.tnb_set_skip_slices(self._h, _ptr(sb, C.c_uint32)))
        return self

    def set_sparse_inds(self, sparse_bits, n_projs):
        """Sparse-index cost model (SimpleCostModelSparseInds); ``sparse_bits=None`` returns to the simple one."""
        if sparse_bits is None:
            self._chk(self._L.tnb_set_sparse_inds(self._h, None, 0))
            return self
        sb = _c(sparse_bits, np.uint32).reshape(-1)
        if sb.shape != (self.W,):
            raise ValueError('sparse_bits must be [ceil(n_inds/32)]')
        if n_projs is None or int(n_projs) != n_projs or n_projs < 1:
            raise ValueError("'n_projs' must be a positive number.")
        self._chk(self._L.tnb_set_sparse_inds(self._h, _ptr(sb, C.c_uint32), int(n_projs)))
        return self

    @property
    def hyper(self):
        """True if the network (with its output indices) has hyper-indices."""
        return bool(self._L.tnb_is_hyper(self._h))

    def set_mode(self, max_width=None, update_slices_every=10, disable_shared_inds=False, prob=PROB_MH,
                 rng=RNG_PHILOX, layout=LAYOUT_AUTO):
        mw = -1.0 if (max_width is None or math.isinf(max_width)) else float(max_width)
        self.finite = mw >= 0
        self._chk(self._L.tnb_set_mode(self._h, mw, int(update_slices_every), int(bool(disable_shared_inds)),
                                       int(prob), int(rng), int(layout)))
        return self

    def set_new_slices(self, max_number_new_slices):
        """max_number_new_slices of the finite-width core object (stream modes only); 0 = off."""
        self._chk(self._L.tnb_set_new_slices(self._h, int(max_number_new_slices)))
        return self

    def set_prob(self, prob):
        self._chk(self._L.tnb_set_prob(self._h, int(prob)))
        return self

    def set_chains(self, parent, child0, child1, seeds, chain_id0=0):
        p, a, b = (np.atleast_2d(_c(x, np.int32)) for x in (parent, child0, child1))
        s = _c(seeds, np.uint64).reshape(-1)
        if p.shape != (len(s), self.N) or a.shape != p.shape or b.shape != p.shape:
            raise ValueError('trees must be [n_chains][2*n_leaves-1] with one seed per chain')
        self.n_chains = len(s)
        self._chk(self._L.tnb_set_chains(self._h, self.n_chains, _ptr(p, C.c_int32), _ptr(a, C.c_int32),
                                         _ptr(b, C.c_int32), _ptr(s, C.c_uint64), int(chain_id0)))
        return self

    def set_resume(self, mt_state=None, slices=None, best_trees=None, best_slices=None):
        """Continue from a saved state (tnb_set_resume): call right after set_chains."""
        ms = None if mt_state is None else _c(mt_state, np.uint32).reshape(self.n_chains, 625)
        sl = None if slices is None else _c(slices, np.uint32).reshape(self.n_chains, self.W)
        bs = None if best_slices is None else _c(best_slices, np.uint32).reshape(self.n_chains, self.W)
        bp = ba = bb = None
        if best_trees is not None:
            bp, ba, bb = (np.atleast_2d(_c(x, np.int32)) for x in best_trees)
            if bp.shape != (self.n_chains, self.N) or ba.shape != bp.shape or bb.shape != bp.shape:
                raise ValueError('best trees must be [n_chains][2*n_leaves-1]')
        self._chk(self._L.tnb_set_resume(self._h, _ptr(ms, C.c_uint32), _ptr(sl, C.c_uint32), _ptr(bp, C.c_int32),
                                         _ptr(ba, C.c_int32), _ptr(bb, C.c_int32), _ptr(bs, C.c_uint32)))
        return self

    def generate_chains(self, seeds, chain_id0=0, method=TREES_GREEDY):
        """One chain per seed; the initial trees are built on the device (tnb_generate_chains)."""
        s = _c(seeds, np.uint64).reshape(-1)
        self.n_chains = len(s)
        self._chk(self._L.tnb_generate_chains(self._h, self.n_chains, _ptr(s, C.c_uint64), int(chain_id0),
                                              int(method)))
        return self

    def set_stream(self, words):
        w = np.atleast_2d(_c(words, np.uint32))
        if w.shape[0] != self.n_chains:
            raise ValueError('one stream per chain')
        self._chk(self._L.tnb_set_stream(self._h, _ptr(w, C.c_uint32), w.shape[1]))
        return self

    TRACE_DTYPE = np.dtype([('w0', '<u4'), ('w1', '<u4'), ('w2', '<u4'), ('w3', '<u4'), ('d0', '<f8'), ('d1', '<f8')])

    def set_trace(self, n_chains, cap_records, cap_reslices=0):
        """Record the decisions of chains [0, n_chains) of the production kernels (tnb_set_trace); 0 = off."""
        self._trace_caps = (int(cap_records), int(cap_reslices))
        self._chk(self._L.tnb_set_trace(self._h, int(n_chains), int(cap_records), int(cap_reslices)))
        return self

    def trace(self, chain):
        """(records as a structured array of TRACE_DTYPE, candidate slices [n_reslices][W32]) of a traced chain."""
        cap, scap = self._trace_caps
        n, ns = C.c_uint64(0), C.c_uint32(0)
        self._chk(self._L.tnb_get_trace(self._h, int(chain), C.byref(n), None, C.byref(ns), None))
        if n.value > cap or ns.value > scap:
            raise EngineError(f'trace overflow: {n.value} records / {ns.value} re-slices, capacity {cap} / {scap}')
        rec = np.zeros(n.value, self.TRACE_DTYPE)
        sl = np.zeros((ns.value, self.W), np.uint32)
        self._chk(self._L.tnb_get_trace(self._h, int(chain), C.byref(n), rec.ctypes.data_as(C.c_void_p), C.byref(ns),
                                        _ptr(sl, C.c_uint32)))
        return rec, sl

    def node_costs(self, chain):
        o = np.empty(self.N, np.float64)
        self._chk(self._L.tnb_get_node_costs(self._h, int(chain), _ptr(o, C.c_double)))
        return o

    def set_betas(self, betas):
        b = _c(betas, np.float64).reshape(-1)
        self._chk(self._L.tnb_set_betas(self._h, _ptr(b, C.c_double), len(b)))
        return self

    def run(self, until_sweep, timeout_s=None):
        """Advance every chain to sweep `until_sweep`; with `timeout_s` the engine stops between its internal launches
        once the wall clock passes it (``self.reached`` = the sweep index actually reached)."""
        if timeout_s is None:
            self._chk(self._L.tnb_run(self._h, int(until_sweep)))
            self.reached = int(until_sweep)
            return self
        r = C.c_int64(0)
        self._chk(self._L.tnb_run_timed(self._h, int(until_sweep), float(timeout_s), C.byref(r)))
        self.reached = r.value
        return self

    def timing(self):
        ms, n = C.c_double(0), C.c_int64(0)
        self._chk(self._L.tnb_get_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def costs(self):
        t, m = np.empty(self.n_chains), np.empty(self.n_chains)
        self._chk(self._L.tnb_get_costs(self._h, _ptr(t, C.c_double), _ptr(m, C.c_double)))
        return t, m

    def trees(self, best=False, chain0=0, n=None):
        n = self.n_chains - chain0 if n is None else n
        p, a, b = (np.empty((n, self.N), np.int32) for _ in range(3))
        self._chk(self._L.tnb_get_trees(self._h, int(best), int(chain0), int(n), _ptr(p, C.c_int32),
                                        _ptr(a, C.c_int32), _ptr(b, C.c_int32)))
        return p, a, b

    def trees_packed(self, best=False, chain0=0, n=None):
        """[n][n_leaves-1] uint32, child0 | child1 << 16 of internal node n_leaves + i at column i."""
        n = self.n_chains - chain0 if n is None else n
        o = np.empty((n, max(self.n - 1, 0)), np.uint32)
        self._chk(self._L.tnb_get_trees_packed(self._h, int(best), int(chain0), int(n), _ptr(o, C.c_uint32)))
        return o

    def bits(self, chain):
        o = np.empty((self.N, self.W), np.uint32)
        self._chk(self._L.tnb_get_bits(self._h, int(chain), _ptr(o, C.c_uint32)))
        return o

    def slices(self, best=False, chain0=0, n=None):
        n = self.n_chains - chain0 if n is None else n
        o = np.empty((n, self.W), np.uint32)
        self._chk(self._L.tnb_get_slices(self._h, int(best), int(chain0), int(n), _ptr(o, C.c_uint32)))
        return o

    def progress(self):
        s = np.empty(self.n_chains, np.int64)
        p, a, w, d = (np.empty(self.n_chains, np.uint64) for _ in range(4))
        self._chk(self._L.tnb_get_progress(self._h, _ptr(s, C.c_int64), _ptr(p, C.c_uint64), _ptr(a, C.c_uint64),
                                           _ptr(w, C.c_uint64), _ptr(d, C.c_uint64)))
        return dict(sweeps=s, proposals=p, accepts=a, width_rejects=w, words=d)

    def counters(self):
        p, a, s = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._chk(self._L.tnb_get_counters(self._h, C.byref(p), C.byref(a), C.byref(s)))
        return dict(proposals=p.value, accepts=a.value, sweeps=s.value)

    def eval_cost(self, parent, child0, child1, slices=None):
        """(total cost summed in traversal order, partial_cost[root], max log2 width after slicing) per tree."""
        p, a, b = (np.atleast_2d(_c(x, np.int32)) for x in (parent, child0, child1))
        T = p.shape[0]
        sl = None if slices is None else np.atleast_2d(_c(slices, np.uint32))
        ts, tp, mw = np.empty(T), np.empty(T), np.empty(T)
        self._chk(self._L.tnb_eval_cost(self._h, T, _ptr(p, C.c_int32), _ptr(a, C.c_int32), _ptr(b, C.c_int32),
                                        _ptr(sl, C.c_uint32), _ptr(ts, C.c_double), _ptr(tp, C.c_double),
                                        _ptr(mw, C.c_double)))
        return ts, tp, mw

    def flush_l2(self):
        self._chk(self._L.tnb_flush_l2(self._h))

    def config(self):
        v = [C.c_int(0) for _ in range(4)]
        self._chk(self._L.tnb_get_config(self._h, *[C.byref(x) for x in v]))
        return dict(tile=v[0].value, words_per_lane=v[1].value, layout=v[2].value,
                    state_bytes_per_chain=v[3].value)


atexit.register(release_cached_engines)
