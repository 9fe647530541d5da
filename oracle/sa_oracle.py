"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle_sa.so (the plain-C CPU restatement).

Nothing under tnco_b200/ imports this module.  Users: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs (see oracle/sa_oracle.h for the scope statement).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'liboracle_sa.so')

PROB_MH, PROB_GREEDY, PROB_ALWAYS = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, ~1 s)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, 'sa_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, 'port'], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i32p, u32p, u64p, f64p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_double))
        L.ora_create_ex.restype = C.c_void_p
        L.ora_create_ex.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, u32p, C.c_uint64, u64p, C.c_int,
                                    C.c_float, C.c_uint32, C.c_int, u32p, C.c_uint64, u32p, C.POINTER(C.c_int)]
        L.ora_create_ex2.restype = C.c_void_p
        L.ora_create_ex2.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, u32p, C.c_uint64, u64p, C.c_int,
                                     C.c_float, C.c_uint32, C.c_int, u32p, C.c_uint64, u32p, u32p, C.POINTER(C.c_int)]
        L.ora_set_replay.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.ora_replay_state.argtypes = [C.c_void_p, u64p, C.POINTER(C.c_int)]
        L.ora_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ora_traced.restype = C.c_uint64
        L.ora_traced.argtypes = [C.c_void_p]
        L.ora_set_forced_slices.argtypes = [C.c_void_p, u32p, C.POINTER(C.c_uint8), C.c_uint64, f64p, C.c_uint64]
        L.ora_forced_used.restype = C.c_uint64
        L.ora_forced_used.argtypes = [C.c_void_p]
        L.ora_set_max_new_slices.argtypes = [C.c_void_p, C.c_int]
        L.ora_new_slice_counters.argtypes = [C.c_void_p, u64p, u64p]
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, u32p, C.c_uint64, u64p, C.c_int,
                                 C.c_float, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_update.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.ora_run.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int64, C.c_int, C.c_int64]
        L.ora_get_tree.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p]
        L.ora_get_bits.argtypes = [C.c_void_p, C.c_int, u32p]
        L.ora_get_slices.argtypes = [C.c_void_p, C.c_int, u32p]
        for f in ('ora_total_cost', 'ora_min_total_cost', 'ora_log2_total_cost',
                  'ora_log2_min_total_cost'):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_void_p]
        L.ora_get_costs.argtypes = [C.c_void_p, f64p, f64p]
        L.ora_prng_state.argtypes = [C.c_void_p, u32p, C.POINTER(C.c_int)]
        L.ora_counters.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, u64p]
        L.ora_record.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.ora_recorded.restype = C.c_uint64
        L.ora_recorded.argtypes = [C.c_void_p]
        L.ora_mt_stream.argtypes = [C.c_uint32, C.c_uint64, u32p]
        L.ora_tree_cost.restype = C.c_double
        L.ora_tree_cost.argtypes = [C.c_int, C.c_int, i32p, i32p, u32p, C.c_uint64, u64p, u32p, f64p, f64p]
        L.ora_get_contraction.restype = C.c_int
        L.ora_get_contraction.argtypes = [C.c_int, i32p, i32p, i32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Chain:
    """One SA chain == one reference ``Optimizer`` object (infinite_memory or finite_width.greedy)."""

    def __init__(self, parent, child0, child1, node_bits, n_inds, *, dim=2, dims=None, max_width=None,
                 seed=0, disable_shared_inds=False, sparse_bits=None, n_projs=None, skip_bits=None,
                 init_slices=None, max_number_new_slices=0):
        L = lib()
        self.parent0, c0, c1 = _i32(parent), _i32(child0), _i32(child1)
        self.N = len(self.parent0)
        self.n = (self.N + 1) // 2
        self.n_inds = int(n_inds)
        self.W = (self.n_inds + 31) // 32
        nb = np.ascontiguousarray(node_bits, dtype=np.uint32).reshape(self.N, self.W)
        dims_a = None if dims is None else np.ascontiguousarray(dims, dtype=np.uint64)
        err = C.c_int(0)
        self.finite = max_width is not None
        sp = None if sparse_bits is None else np.ascontiguousarray(sparse_bits, dtype=np.uint32).reshape(self.W)
        sk = None if skip_bits is None else np.ascontiguousarray(skip_bits, dtype=np.uint32).reshape(self.W)
        isl = None if init_slices is None else np.ascontiguousarray(init_slices, dtype=np.uint32).reshape(self.W)
        self._h = L.ora_create_ex2(self.n, self.n_inds, _p(self.parent0, C.c_int32), _p(c0, C.c_int32),
                                       _p(c1, C.c_int32), _p(nb, C.c_uint32), int(dim),
                                       None if dims_a is None else _p(dims_a, C.c_uint64), int(self.finite),
                                       float(max_width if self.finite else 0.0), int(seed) & 0xFFFFFFFF,
                                       int(disable_shared_inds), None if sp is None else _p(sp, C.c_uint32),
                                       int(n_projs or 0), None if sk is None else _p(sk, C.c_uint32),
                                       None if isl is None else _p(isl, C.c_uint32), C.byref(err))
        if not self._h:
            raise ValueError('Precision is too low.' if err.value == 2 else
                             "'n_projs' must be a positive number." if err.value == 3 else 'invalid input')
        self._rec = None
        if max_number_new_slices:
            L.ora_set_max_new_slices(self._h, int(max_number_new_slices))

    def new_slice_counters(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().ora_new_slice_counters(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def __del__(self):
        if getattr(self, '_h', None):
            lib().ora_destroy(self._h)
            self._h = None

    def update(self, beta, prob=PROB_MH, update_slices=True):
        lib().ora_update(self._h, prob, float(beta), int(bool(update_slices)))

    def run(self, betas, prob=PROB_MH, update_slices_every=0, sweep_offset=0):
        b = np.ascontiguousarray(betas, dtype=np.float64)
        lib().ora_run(self._h, prob, _p(b, C.c_double), len(b), int(update_slices_every), int(sweep_offset))

    def tree(self, best=False):
        p, a, b = (np.empty(self.N, np.int32) for _ in range(3))
        lib().ora_get_tree(self._h, int(best), _p(p, C.c_int32), _p(a, C.c_int32), _p(b, C.c_int32))
        return p, a, b

    def bits(self, best=False):
        o = np.empty((self.N, self.W), np.uint32)
        lib().ora_get_bits(self._h, int(best), _p(o, C.c_uint32))
        return o

    def slices(self, best=False):
        o = np.empty(self.W, np.uint32)
        lib().ora_get_slices(self._h, int(best), _p(o, C.c_uint32))
        return o

    def costs(self):
        cc, pc = np.empty(self.N), np.empty(self.N)
        lib().ora_get_costs(self._h, _p(cc, C.c_double), _p(pc, C.c_double))
        return cc, pc

    total_cost = property(lambda s: lib().ora_total_cost(s._h))
    min_total_cost = property(lambda s: lib().ora_min_total_cost(s._h))
    log2_total_cost = property(lambda s: lib().ora_log2_total_cost(s._h))
    log2_min_total_cost = property(lambda s: lib().ora_log2_min_total_cost(s._h))

    def prng_state(self):
        s = np.empty(624, np.uint32)
        pos = C.c_int(0)
        lib().ora_prng_state(self._h, _p(s, C.c_uint32), C.byref(pos))
        return s, pos.value

    def prng_state_str(self):
        """libstdc++ ``operator<<(ostream&, mt19937)`` text == reference ``Optimizer.prng_state``."""
        s, pos = self.prng_state()
        return ' '.join(map(str, s.tolist())) + ' ' + str(pos)

    def counters(self):
        v = [C.c_uint64(0) for _ in range(5)]
        lib().ora_counters(self._h, *[C.byref(x) for x in v])
        return dict(zip(('proposals', 'accepts', 'sweeps', 'words_drawn', 'width_rejects'),
                        (x.value for x in v)))

    TRACE_DTYPE = np.dtype([('B', '<i4'), ('A', '<i4'), ('pick0', 'u1'), ('gate', 'u1'), ('acc', 'u1'), ('coin', 'u1'),
                            ('pad', '<u4'), ('delta', '<f8'), ('total', '<f8'), ('u', '<f8'), ('p', '<f8')])

    def set_replay(self, words):
        """Feed every 32-bit draw from `words` (the reference's data-dependent draw order) instead of the generator."""
        self._rp = np.ascontiguousarray(words, dtype=np.uint32)
        lib().ora_set_replay(self._h, _p(self._rp, C.c_uint32), len(self._rp))

    def replay_state(self):
        k, o = C.c_uint64(0), C.c_int(0)
        lib().ora_replay_state(self._h, C.byref(k), C.byref(o))
        return k.value, bool(o.value)

    def trace(self, cap):
        self._tr = np.zeros(int(cap), self.TRACE_DTYPE)
        assert self.TRACE_DTYPE.itemsize == 48
        lib().ora_trace(self._h, self._tr.ctypes.data_as(C.c_void_p), int(cap))

    def traced(self):
        n = int(lib().ora_traced(self._h))
        if n > len(self._tr):
            raise RuntimeError('trace buffer overflow')
        return self._tr[:n]

    def set_forced_slices(self, candidates, keep):
        """k-th re-slice: candidate slices candidates[k], kept iff keep[k]; see reslice_log()."""
        self._fs = np.ascontiguousarray(candidates, dtype=np.uint32).reshape(-1, self.W)
        self._fk = np.ascontiguousarray(keep, dtype=np.uint8).reshape(-1)
        assert len(self._fs) == len(self._fk)
        self._rl = np.zeros((max(len(self._fk), 1), 2), np.float64)
        lib().ora_set_forced_slices(self._h, _p(self._fs, C.c_uint32), self._fk.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    len(self._fk), _p(self._rl, C.c_double), len(self._rl))

    def reslice_log(self):
        """(re-slices consumed, [k][2] = reference criterion: cost under the candidate, cost under the current)."""
        k = int(lib().ora_forced_used(self._h))
        return k, self._rl[:min(k, len(self._rl))]

    def record(self, cap):
        self._rec = np.zeros(int(cap), np.uint32)
        lib().ora_record(self._h, _p(self._rec, C.c_uint32), int(cap))

    def recorded(self):
        n = int(lib().ora_recorded(self._h))
        if n > len(self._rec):
            raise RuntimeError('record buffer overflow')
        return self._rec[:n].copy()


def mt_stream(seed, n):
    out = np.empty(int(n), np.uint32)
    lib().ora_mt_stream(int(seed) & 0xFFFFFFFF, int(n), _p(out, C.c_uint32))
    return out


def tree_cost(child0, child1, node_bits, n_inds, dim=2, dims=None, slices=None):
    """(total cost summed in traverse order, max log2 width after slicing, partial_cost[root])."""
    c0, c1 = _i32(child0), _i32(child1)
    N = len(c0)
    W = (n_inds + 31) // 32
    nb = np.ascontiguousarray(node_bits, dtype=np.uint32).reshape(N, W)
    dims_a = None if dims is None else np.ascontiguousarray(dims, dtype=np.uint64)
    sl = None if slices is None else np.ascontiguousarray(slices, dtype=np.uint32)
    mw, pcr = C.c_double(0), C.c_double(0)
    t = lib().ora_tree_cost((N + 1) // 2, n_inds, _p(c0, C.c_int32), _p(c1, C.c_int32), _p(nb, C.c_uint32),
                            int(dim), None if dims_a is None else _p(dims_a, C.c_uint64),
                            None if sl is None else _p(sl, C.c_uint32), C.byref(mw), C.byref(pcr))
    return t, mw.value, pcr.value


def get_contraction(child0, child1):
    c0, c1 = _i32(child0), _i32(child1)
    n = (len(c0) + 1) // 2
    tr = np.empty((max(n - 1, 0), 3), np.int32)
    k = lib().ora_get_contraction(n, _p(c0, C.c_int32), _p(c1, C.c_int32), _p(tr, C.c_int32))
    return tr[:k]
