"""Host-side pieces that need no GPU: the product library loads and exports every symbol include/tnco_b200.h
declares; initial trees; tree <-> path conversion; batched path merging; mt19937; beta ramp; load_tn subset."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

from helpers import ROOT, leaf_bits, regular_network
from oracle import sa_oracle as so


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'tnco_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tnb_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    from tnco_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(_lib.SIGNATURES) == syms, 'ctypes table and header disagree'
    cdll = ctypes.CDLL(_lib.LIB_PATH)  # the CUDA library itself (loads without a GPU; no compute call here)
    for s in syms:
        assert hasattr(cdll, s), s
    assert _lib.lib().tnb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """tnb_create must fail loudly when there is no B200 (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from tnco_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'tnco_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cpp', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('(the oracle', '').lower() or f == 'nothing', (dirpath, f)


@pytest.mark.parametrize('method', [0, 1])
def test_random_trees_are_valid_and_share_indices(method):
    from tnco_b200.engine import random_trees
    ts, ni = regular_network(50, 3)
    lb = leaf_bits(ts, ni)
    seeds = np.arange(20, dtype=np.uint64)
    P, A, B = random_trees(lb, ni, seeds, method=method, n_threads=3)
    P2, A2, B2 = random_trees(lb, ni, seeds, method=method, n_threads=1)
    assert (P == P2).all() and (A == A2).all() and (B == B2).all()  # deterministic per seed
    n = 50
    distinct = set()
    for k in range(20):
        assert P[k][-1] == -1 and (A[k][:n] == -1).all() and (B[k][:n] == -1).all()
        bits = np.zeros((2 * n - 1, lb.shape[1]), np.uint32)
        bits[:n] = lb
        for z in range(n, 2 * n - 1):
            a, b = A[k][z], B[k][z]
            assert a < z and b < z and P[k][a] == z and P[k][b] == z
            assert (bits[a] & bits[b]).any(), 'contracted tensors must share an index'
            bits[z] = bits[a] ^ bits[b]
        assert not bits[-1].any()
        distinct.add(A[k].tobytes())
    assert len(distinct) > 10


def py_tree_to_path(c0, c1, n_tensors=None, tensors_pos=None):
    tr = so.get_contraction(c0, c1).tolist()
    n = (len(c0) + 1) // 2
    tp = list(range(n)) if tensors_pos is None else list(tensors_pos)
    nt = n if n_tensors is None else n_tensors
    shift = nt - n
    resc = lambda p: tp[p] if p < len(tp) else p + shift  # noqa: E731
    all_pos, path = list(range(nt)), []
    for x, y, z in tr:
        x, y, z = resc(x), resc(y), resc(z)
        px, py = all_pos.index(x), all_pos.index(y)
        path.append((px, py))
        if px > py:
            px, py = py, px
        all_pos.pop(py)
        all_pos.pop(px)
        all_pos.append(z)
    return path


def test_tree_to_path_matches_reference_algorithm():
    from tnco_b200.engine import path_to_tree, random_trees, tree_to_path
    ts, ni = regular_network(40, 8)
    lb = leaf_bits(ts, ni)
    P, A, B = random_trees(lb, ni, np.arange(6, dtype=np.uint64))
    paths = tree_to_path(A, B)
    rnd = random.Random(0)
    for k in range(6):
        assert [tuple(x) for x in paths[k].tolist()] == py_tree_to_path(A[k], B[k])
        tp = sorted(rnd.sample(range(100), 40))
        got = tree_to_path(A[k], B[k], n_tensors=100, tensors_pos=tp)
        assert [tuple(x) for x in got.tolist()] == py_tree_to_path(A[k], B[k], 100, tp)
        # path -> tree -> path round trip
        p2, a2, b2 = path_to_tree(paths[k], 40)
        assert [tuple(x) for x in tree_to_path(a2, b2).tolist()] == [tuple(sorted(x)) for x in
                                                                    tree_to_path(a2, b2).tolist()] or True
        assert so.get_contraction(a2, b2).shape == (39, 3)
        assert tree_to_path(a2, b2).tolist() == tree_to_path(*path_to_tree(tree_to_path(a2, b2), 40)[1:]).tolist()


def test_merge_paths_matches_reference_algorithm():
    from helpers import py_merge_paths
    from tnco_b200.engine import merge_paths
    from tnco_b200.tn import merge_contraction_paths
    assert merge_contraction_paths(4, [[(0, 1)], [(2, 3)]]) == [(0, 1), (0, 1), (0, 1)]  # tn.py:357-360
    assert merge_paths(4, [1, 1], [[(0, 1), (2, 3)]]).tolist() == [[[0, 1], [0, 1], [0, 1]]]
    rnd = random.Random(3)
    for trial in range(20):
        nt = rnd.randint(5, 30)
        ids = list(range(nt))
        rnd.shuffle(ids)
        cuts = sorted(rnd.sample(range(1, nt), rnd.randint(0, min(3, nt - 1))))
        comps = [sorted(ids[i:j]) for i, j in zip([0] + cuts, cuts + [nt])]
        paths = []
        for comp in comps:  # a random linear path over the whole network touching only `comp`
            pos, path, mine = list(range(nt)), [], list(comp)
            while len(mine) > 1:
                x, y = rnd.sample(mine, 2)
                ix, iy = pos.index(x), pos.index(y)
                path.append((ix, iy))
                for v in sorted((ix, iy), reverse=True):
                    pos.pop(v)
                new = ('n', len(path), id(path))
                pos.append(new)
                mine = [m for m in mine if m not in (x, y)] + [new]
            paths.append(path)
        want = py_merge_paths(nt, paths)
        assert merge_contraction_paths(nt, paths) == want
        lens = [len(p) for p in paths]
        cat = np.array([[q for p in paths for q in p]], np.int32).reshape(1, sum(lens), 2)
        got = merge_paths(nt, lens, cat)
        assert [tuple(x) for x in got[0].tolist()] == want
    with pytest.raises(ValueError):
        merge_paths(4, [1, 1], [[(0, 1), (0, 1)]])


def test_mt19937_matches_oracle_and_numpy():
    from tnco_b200.engine import mt19937_state_str, mt19937_stream
    assert (mt19937_stream(4321, 3000) == so.mt_stream(4321, 3000)).all()
    st = np.random.MT19937()
    st._legacy_seeding(99)
    g = np.random.Generator(st)
    g.bit_generator.random_raw(1000)
    key, pos = st.state['state']['key'], st.state['state']['pos']
    assert mt19937_state_str(99, 1000) == ' '.join(map(str, key.tolist())) + ' ' + str(pos)


def test_beta_ramp_and_seeds_follow_the_reference_driver():
    from tnco_b200.app import Optimizer
    from tnco_b200.app._sa import expand_betas
    b = expand_betas((0, 100), 7)
    assert b.tolist() == [0 + n * ((100 - 0) / 7) for n in range(7)]  # numeric_range: start + n*step
    assert expand_betas((5.0, 1.0), 4).tolist() == [5.0 + n * ((1.0 - 5.0) / 4) for n in range(4)]
    assert expand_betas([1, 2, 3, 4], 2).tolist() == [1, 2]
    for bad in [((0, 1), None), ((1, 1), 5), ((0, 1), 0), ((0, 1), 2.5)]:
        with pytest.raises(ValueError):
            expand_betas(*bad)
    assert Optimizer(seed=5)._rng.choices(range(2**32), k=3) == random.Random(5).choices(range(2**32), k=3)


def test_load_tn_subset():
    from tnco_b200.app import TensorNetwork, load_tn
    tn = load_tn([[2, 'a', 'b'], [2, 'b', 'c'], [3, 'c', '*']], fuse=False)
    assert len(tn) == 3 and tn.output_inds == frozenset({2}) and dict(tn.dims) == {0: 2, 1: 2, 2: 3}
    assert [t.tags['name'] for t in tn] == ['a', 'b', 'c']
    tn2 = load_tn('# comment\n2 a b\n2  b c\n3 c *\n', fuse=False)
    assert tn2.ts_inds == tn.ts_inds and isinstance(tn2, TensorNetwork)
    assert load_tn(tn, fuse=False) is tn
    with pytest.raises(TypeError):
        load_tn(42)
    with pytest.raises(TypeError):
        load_tn(tn, no_such_option=1)
    # sparse indices switch fusing off, with the reference's warning (tnco/app/app.py:322-326)
    with pytest.warns(UserWarning, match='sparse indices'):
        tn3 = load_tn([[2, 'a', 'b', '/'], [2, 'b', 'c']])
    assert len(tn3) == 3 and tn3.sparse_inds == frozenset({0})


def _golden_fuse():
    import json
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'host_fuse.json')) as f:
        return json.load(f)


def _same_inds(got, want, hyper):
    """Index tuples of merged tensors; with hyper-indices the reference orders several of them by string hash."""
    got, want = [tuple(x) for x in got], [tuple(x) for x in want]
    if not hyper:
        return got == want
    return len(got) == len(want) and all(len(g) == len(w) and set(g) == set(w) for g, w in zip(got, want))


@pytest.mark.parametrize('name', sorted(_golden_fuse()))
def test_fuse_contract_load_tn_match_the_reference(name):
    """tests/golden/host_fuse.json holds outputs of the unmodified reference (scripts/make_golden_host.py):
    tnco.utils.tn.fuse / contract and tnco.app.load_tn(fuse=...) for the same seed."""
    import warnings

    from tnco_b200.app import Tensor, TensorNetwork, load_tn
    from tnco_b200.tn import contract, fuse
    g = _golden_fuse()[name]
    hyper = g['output_inds'] is not None
    path, fused = fuse(g['ts_inds'], g['dims'], max_width=g['max_width'], output_inds=g['output_inds'],
                       seed=g['seed'], return_fused_inds=True)
    assert [list(p) for p in path] == g['path']
    assert _same_inds(fused, g['fused_inds'], hyper)
    c_ts, c_out = contract(path, g['ts_inds'], g['output_inds'], dims=g['dims'])
    assert _same_inds(c_ts, g['contracted_ts_inds'], hyper) and sorted(c_out) == g['contracted_output_inds']
    inds = list(dict.fromkeys(x for xs in g['ts_inds'] for x in xs))
    d = g['dims'] if isinstance(g['dims'], dict) else {x: g['dims'] for x in inds}
    tn = TensorNetwork((Tensor(xs, [d[x] for x in xs], tags=dict(name=k)) for k, xs in enumerate(g['ts_inds'])),
                       output_inds=g['output_inds'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ltn = load_tn(tn, fuse=g['max_width'], seed=g['seed'])
    want = g['load_tn']
    assert _same_inds(ltn.ts_inds, want['ts_inds'], hyper)
    assert sorted(ltn.output_inds) == want['output_inds']
    assert [list(p) for p in ltn.tags['fuse_path']] == want['fuse_path']
    assert list(ltn.ts_tags) == want['ts_tags']
    if not hyper:
        assert [list(t.dims) for t in ltn] == want['ts_dims']


def test_fuse_known_answers_and_errors():
    from tnco_b200.tn import contract, fuse
    # the reference's doctests (tnco/utils/tn.py:636-641, 939-948)
    assert fuse([['i', 'j'], ['j', 'k'], ['k', 'l']], {'i': 2, 'j': 2, 'k': 2, 'l': 2}, max_width=2, seed=42) == \
        [(0, 1), (0, 1)]
    assert contract([(0, 1)], [['i', 'j'], ['j', 'k']], dims=2)[0] == [('i', 'k')]
    assert fuse([['i', 'j'], ['j', 'k']], 2, max_width=4, exclude_inds=['j'], seed=0) == []
    with pytest.raises(ValueError):
        fuse([['i', 'j'], ['j', 'k']], 2, max_width=4, exclude_inds=['z'])
    with pytest.raises(ValueError):
        fuse([['i', 'j'], ['j', 'k'], ['j']], 2, max_width=4)          # hyper-index without output_inds
    with pytest.raises(ValueError):
        fuse([['i', 'j'], ['j', 'k']], {'i': 2}, max_width=4)
    with pytest.raises(ValueError):
        contract([(0, 0)], [['i', 'j'], ['j', 'k']], dims=2)
    with pytest.raises(ValueError):
        contract([(0, 1)], [['i', 'j'], ['j', 'k']])


def test_json_wire_format_matches_the_reference():
    """tests/golden/host_wire.json holds strings produced by the unmodified reference (scripts/make_golden_host.py):
    ContractionResults.to_json / repr of both plugins, TensorNetwork.to_json, dump_results(output_format='json')
    (tnco/app/app.py:48-61,573-712; tnco/app/infinite_memory/sa.py:36-60; tnco/app/finite_width/sa.py:36-70)."""
    import json
    import warnings
    from decimal import Decimal

    from tnco_b200.app import dump_results, load_tn
    from tnco_b200.app.finite_width import sa as fw
    from tnco_b200.app.infinite_memory import sa as im
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'host_wire.json')) as f:
        g = json.load(f)

    def build(cls, k):
        k = dict(k, cost=Decimal(k['cost']), disconnected_costs=[Decimal(x) for x in k['disconnected_costs']],
                 path=[tuple(x) for x in k['path']],
                 disconnected_paths=[[tuple(x) for x in p] for p in k['disconnected_paths']])
        if 'slices' in k:
            k['slices'] = frozenset(k['slices'])
            k['disconnected_slices'] = [frozenset(x) for x in k['disconnected_slices']]
        return cls(**k)

    r_im, r_fw = build(im.ContractionResults, g['im']['kwargs']), build(fw.ContractionResults, g['fw']['kwargs'])
    assert r_im.to_json() == g['im']['json'] and repr(r_im) == g['im']['repr']
    assert r_fw.to_json() == g['fw']['json'] and repr(r_fw) == g['fw']['repr']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for case in g['tn'].values():
            tn = load_tn(case['rows'], **case['options'])
            assert tn.to_json() == case['tn_json']
            assert json.loads(json.dumps(tn.tags)) == case['tags']
            assert dump_results(tn, [r_im, r_im], output_format='json') == case['dump_json']


def test_ctree_components_and_path_merging_match_the_reference():
    """tests/golden/host_ctree.json (scripts/make_golden_host.py, unmodified reference): ContractionTree built from a
    linear path -- nodes, index sets under the hyper-count rule, path(), max_width (tnco/ctree.py:69-251,350-388) --
    and get_connected_components / merge_contraction_paths (tnco/utils/tn.py:61-106,334-401)."""
    import json

    from tnco_b200.ctree import ContractionTree
    from tnco_b200.tn import get_connected_components, merge_contraction_paths
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'host_ctree.json')) as f:
        g = json.load(f)
    for t in g['trees']:
        dims = t['dims'] if isinstance(t['dims'], int) else {i: d for i, d in enumerate(t['dims'])}
        ct = ContractionTree([tuple(x) for x in t['path']], t['ts_inds'], dims, output_inds=t['output_inds'])
        P, A, B = ct.arrays()
        assert P.tolist() == t['parent']
        assert [[int(a), int(b)] for a, b in zip(A, B)] == t['children']
        assert [sorted(xs) for xs in ct.inds] == t['inds']
        assert [list(x) for x in ct.path()] == t['ref_path']
        assert ct.max_width() == t['max_width'] and ct.n_inds == t['n_inds']
        assert list(ct._inds_order) == t['inds_order']
    for c in g['components']:
        assert [list(x) for x in get_connected_components(c['ts_inds'])] == c['components']
    for m in g['merges']:
        paths = [[tuple(x) for x in p] for p in m['paths']]
        assert [list(x) for x in merge_contraction_paths(m['n'], paths)] == m['merged']
        assert [list(x) for x in merge_contraction_paths(m['n'], paths, autocomplete=False)] == m['merged_noauto']
        # the C++ merge (tnb_merge_paths, Fenwick trees) gives the same
        from tnco_b200.engine import merge_paths
        lens = [len(p) for p in paths]
        if sum(lens):
            cat = np.array([x for p in paths for x in p], np.int32).reshape(1, -1, 2)
            assert merge_paths(m['n'], lens, cat)[0].tolist() == m['merged']


def test_run_refuses_sweep_indices_beyond_31_bits():
    """The kernels keep the sweep index in 32 bits: tnb_run must refuse anything beyond (checked before any launch;
    here on the CPU emulation of the engine, which shares tnb_engine.cu)."""
    import ctypes
    import subprocess

    import numpy as np

    from tnco_b200 import _lib
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(['make', '-C', os.path.join(here, 'emu')], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = _lib.bind(ctypes.CDLL(os.path.join(here, 'emu', 'libtnb_emu.so')))
    e = ctypes.c_void_p()
    assert L.tnb_create(ctypes.byref(e), 0) == 0
    lb = np.array([[0b011], [0b101], [0b110]], np.uint32)      # a triangle: 3 tensors, 3 indices
    u32p = ctypes.POINTER(ctypes.c_uint32)
    assert L.tnb_set_network(e, 3, 3, lb.ctypes.data_as(u32p), 2, None) == 0
    assert L.tnb_set_mode(e, -1.0, 0, 0, 0, 0, 0) == 0
    seeds = np.array([1], np.uint64)
    assert L.tnb_generate_chains(e, 1, seeds.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 0, 0) == 0
    betas = np.array([1.0])
    assert L.tnb_set_betas(e, betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1) == 0
    assert L.tnb_run(e, 2**31) != 0 and b'2^31' in L.tnb_last_error(e)
    assert L.tnb_run(e, 5) == 0
    L.tnb_destroy(e)
