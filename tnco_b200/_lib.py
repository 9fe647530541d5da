"""ctypes binding of libtnco_b200.so (the C-ABI declared in include/tnco_b200.h).

The library is hand-written CUDA for sm_100a; there is no CPU fallback: if the shared library is missing
the import of any compute entry point fails loudly, and ``tnb_create`` fails when no B200 is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtnco_b200.so')

PROB_MH, PROB_GREEDY, PROB_ALWAYS = 0, 1, 2
RNG_PHILOX, RNG_MT19937, RNG_REPLAY = 0, 1, 2
LAYOUT_AUTO, LAYOUT_INTERLEAVED, LAYOUT_SPLIT, LAYOUT_SMEM = 0, 1, 2, 3
TREES_GREEDY, TREES_RANDOM = 0, 1

i32p, u32p, u64p, i64p, f64p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                C.POINTER(C.c_int64), C.POINTER(C.c_double))
intp = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol of include/tnco_b200.h
SIGNATURES = {
    'tnb_version': (C.c_int, []),
    'tnb_last_error': (C.c_char_p, [C.c_void_p]),
    'tnb_random_trees': (C.c_int, [C.c_int, C.c_int, u32p, C.c_int, u64p, C.c_int, C.c_int, i32p, i32p, i32p]),
    'tnb_random_trees_out': (C.c_int, [C.c_int, C.c_int, u32p, u32p, C.c_int, u64p, C.c_int, C.c_int, i32p, i32p, i32p]),
    'tnb_tree_to_path': (C.c_int, [C.c_int, C.c_int, i32p, i32p, C.c_int, i32p, i32p]),
    'tnb_merge_paths': (C.c_int, [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]),
    'tnb_path_to_tree': (C.c_int, [C.c_int, i32p, i32p, i32p, i32p]),
    'tnb_mt19937_stream': (None, [C.c_uint32, C.c_uint64, u32p]),
    'tnb_mt19937_state': (None, [C.c_uint32, C.c_uint64, u32p, i32p]),
    'tnb_mt19937_advance': (None, [u32p, i32p, C.c_uint64]),
    'tnb_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    'tnb_destroy': (None, [C.c_void_p]),
    'tnb_set_network': (C.c_int, [C.c_void_p, C.c_int, C.c_int, u32p, C.c_uint64, u64p]),
    'tnb_set_output_inds': (C.c_int, [C.c_void_p, u32p]),
    'tnb_is_hyper': (C.c_int, [C.c_void_p]),
    'tnb_set_sparse_inds': (C.c_int, [C.c_void_p, u32p, C.c_uint64]),
    'tnb_set_skip_slices': (C.c_int, [C.c_void_p, u32p]),
    'tnb_set_mode': (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'tnb_set_prob': (C.c_int, [C.c_void_p, C.c_int]),
    'tnb_set_new_slices': (C.c_int, [C.c_void_p, C.c_int]),
    'tnb_set_update_slices': (C.c_int, [C.c_void_p, C.c_int]),
    'tnb_set_chains': (C.c_int, [C.c_void_p, C.c_int, i32p, i32p, i32p, u64p, C.c_uint64]),
    'tnb_generate_chains': (C.c_int, [C.c_void_p, C.c_int, u64p, C.c_uint64, C.c_int]),
    'tnb_set_resume': (C.c_int, [C.c_void_p, u32p, u32p, i32p, i32p, i32p, u32p]),
    'tnb_set_stream': (C.c_int, [C.c_void_p, u32p, C.c_uint64]),
    'tnb_set_trace': (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32]),
    'tnb_get_trace': (C.c_int, [C.c_void_p, C.c_int, u64p, C.c_void_p, u32p, u32p]),
    'tnb_get_node_costs': (C.c_int, [C.c_void_p, C.c_int, f64p]),
    'tnb_set_betas': (C.c_int, [C.c_void_p, f64p, C.c_int64]),
    'tnb_run': (C.c_int, [C.c_void_p, C.c_int64]),
    'tnb_run_timed': (C.c_int, [C.c_void_p, C.c_int64, C.c_double, i64p]),
    'tnb_get_timing': (C.c_int, [C.c_void_p, f64p, i64p]),
    'tnb_get_costs': (C.c_int, [C.c_void_p, f64p, f64p]),
    'tnb_get_trees': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]),
    'tnb_get_trees_packed': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, u32p]),
    'tnb_get_bits': (C.c_int, [C.c_void_p, C.c_int, u32p]),
    'tnb_get_slices': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, u32p]),
    'tnb_get_progress': (C.c_int, [C.c_void_p, i64p, u64p, u64p, u64p, u64p]),
    'tnb_get_counters': (C.c_int, [C.c_void_p, u64p, u64p, u64p]),
    'tnb_eval_cost': (C.c_int, [C.c_void_p, C.c_int, i32p, i32p, i32p, u32p, f64p, f64p, f64p]),
    'tnb_flush_l2': (C.c_int, [C.c_void_p]),
    'tnb_get_config': (C.c_int, [C.c_void_p, intp, intp, intp, intp]),
    'tnb_group_create': (C.c_int, [C.POINTER(C.c_void_p), intp, C.c_int]),
    'tnb_group_destroy': (None, [C.c_void_p]),
    'tnb_group_size': (C.c_int, [C.c_void_p]),
    'tnb_group_engine': (C.c_void_p, [C.c_void_p, C.c_int]),
    'tnb_group_last_error': (C.c_char_p, [C.c_void_p]),
    'tnb_group_set_network': (C.c_int, [C.c_void_p, C.c_int, C.c_int, u32p, C.c_uint64, u64p, u32p]),
    'tnb_group_set_mode': (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'tnb_group_set_betas': (C.c_int, [C.c_void_p, f64p, C.c_int64]),
    'tnb_group_generate_chains': (C.c_int, [C.c_void_p, C.c_int, u64p, C.c_int]),
    'tnb_group_run': (C.c_int, [C.c_void_p, C.c_int64, C.c_double, i64p]),
    'tnb_group_get_costs': (C.c_int, [C.c_void_p, f64p, f64p]),
    'tnb_group_get_counters': (C.c_int, [C.c_void_p, u64p, u64p, u64p]),
    'tnb_group_get_best': (C.c_int, [C.c_void_p, f64p, i64p, i32p, i32p, i32p, u32p]),
}

_LIB = None


def bind(cdll):
    for name, (res, args) in SIGNATURES.items():
        f = getattr(cdll, name)  # AttributeError if the symbol is missing
        f.restype = res
        f.argtypes = args
    return cdll


def lib():
    """The loaded product library.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build the CUDA extension first (make -C tnco_b200/csrc, or '
                '`python -c "import __graft_entry__ as g; g.build()"`). tnco_b200 has no CPU fallback.')
        _LIB = bind(C.CDLL(LIB_PATH))
    return _LIB
