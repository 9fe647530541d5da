#!/usr/bin/env python3
"""Generate tests/golden/host_fuse.json from the UNMODIFIED reference Python layer (/root/reference/tnco), imported
here with the stand-ins of scripts/ref_shims for the packages this image lacks.  Build container only; the
fixture is committed so machines without /root/reference can check tnco_b200's host-side pre-processing
(`fuse`, structure-only `contract`, `load_tn(..., fuse=...)`) against outputs of the real reference.

Index names are strings 'i<k>'.  With several hyper-indices on one merged tensor the reference orders them by
frozenset iteration, i.e. by the per-process string hash; PYTHONHASHSEED is pinned to 0 while generating and
the test compares such index tuples as sets.
"""
import json
import os
import sys
import warnings

if os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'scripts', 'ref_shims'), '/root/reference', os.path.join(ROOT, 'oracle', '_ref'),
                ROOT, os.path.join(ROOT, 'tests')]

import tnco.app.app as ref_app  # noqa: E402
import tnco.utils.tn as ref_tn  # noqa: E402
from helpers import hyper_network, regular_network  # noqa: E402
from tnco.app.tn import Tensor as RefTensor  # noqa: E402
from tnco.app.tn import TensorNetwork as RefTN  # noqa: E402

from tnco_b200 import networks  # noqa: E402


def name(ts):
    return [['i%d' % x for x in xs] for xs in ts]


def cases():
    # name, ts_inds, output_inds (None: dangling), dims (int or per-index list), max_width, seed
    ts, ni = regular_network(24, 1)
    yield 'reg24_w4', name(ts), None, 2, 4, 0
    yield 'reg24_w6', name(ts), None, 2, 6, 5
    ts, ni = regular_network(64, 2)
    yield 'reg64_w4', name(ts), None, 2, 4, 1
    yield 'reg64_w7_d3', name(ts), None, 3, 7, 2
    yield 'reg64_mixed', name(ts), None, {('i%d' % k): (2, 4, 8)[k % 3] for k in range(ni)}, 5, 3
    ts, ni = networks.grid_rqc(4, 4, 8)
    yield 'grid4x4_w4', name(ts), None, 2, 4, 4
    yield 'grid4x4_w8', name(ts), None, 2, 8, 6
    ts, ni, out = hyper_network(32, 3)
    yield 'hyper32_w4', name(ts), ['i%d' % x for x in out], 2, 4, 7
    ts, ni, out = hyper_network(64, 5, n_hyper=10, n_open=5)
    yield 'hyper64_w5', name(ts), ['i%d' % x for x in out], 2, 5, 8
    ts, ni = networks.grid_rqc(6, 6, 12)
    yield 'c2_w4', name(ts), None, 2, 4, 9


def main():
    out = {}
    warnings.simplefilter('ignore')
    for nm, ts, outs, dims, mw, seed in cases():
        path, fused = ref_tn.fuse(ts, dims, max_width=mw, output_inds=outs, seed=seed, return_fused_inds=True)
        c_ts, c_out = ref_tn.contract(path, ts, outs, dims=dims)
        all_inds = list(dict.fromkeys(x for xs in ts for x in xs))
        d = dims if isinstance(dims, dict) else {x: dims for x in all_inds}
        tn = RefTN((RefTensor(xs, [d[x] for x in xs], tags=dict(name=k)) for k, xs in enumerate(ts)),
                   output_inds=outs)
        ltn = ref_app.load_tn(tn, fuse=mw, seed=seed)
        out[nm] = dict(ts_inds=ts, output_inds=outs, dims=dims, max_width=mw, seed=seed,
                       path=[list(p) for p in path], fused_inds=[list(x) for x in fused],
                       contracted_ts_inds=[list(x) for x in c_ts], contracted_output_inds=sorted(c_out),
                       load_tn=dict(ts_inds=[list(x) for x in ltn.ts_inds], ts_dims=[list(t.dims) for t in ltn],
                                    output_inds=sorted(ltn.output_inds), ts_tags=list(ltn.ts_tags),
                                    fuse_path=[list(p) for p in ltn.tags['fuse_path']]))
        print(nm, len(ts), '->', len(c_ts), 'tensors,', len(path), 'merges')
    # the reference's own doctest (tnco/utils/tn.py:636-641)
    assert ref_tn.fuse([['i', 'j'], ['j', 'k'], ['k', 'l']], 2, max_width=2, seed=42) == [(0, 1), (0, 1)]
    with open(os.path.join(ROOT, 'tests', 'golden', 'host_fuse.json'), 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    wire_format()
    ctree_and_paths()


def ctree_and_paths():
    """tests/golden/host_ctree.json: tnco.ctree.ContractionTree (path -> nodes + index sets with the hyper-count rule,
    path(), max_width; tnco/ctree.py:69-251,350-388), get_connected_components and merge_contraction_paths
    (tnco/utils/tn.py:61-106,334-401) of the reference on seeded random inputs."""
    import random

    from helpers import random_tree
    from tnco.ctree import ContractionTree as RefCT

    def linear_path(a, b):
        n = (len(a) + 1) // 2
        pos = list(range(n))
        path = []
        for z in range(n, 2 * n - 1):
            x, y = sorted((pos.index(int(a[z])), pos.index(int(b[z]))))
            path.append([x, y])
            pos.pop(y)
            pos.pop(x)
            pos.append(z)
        return path

    trees = []
    for n, seed, hyper in ((8, 1, False), (20, 2, False), (40, 3, False), (16, 4, True), (36, 5, True)):
        if hyper:
            ts, ni, outs = hyper_network(n, seed)
        else:
            ts, ni = regular_network(n, seed)
            outs = []
        p, a, b, bits = random_tree(ts, ni, seed + 10, outs)
        path = linear_path(a, b)
        dims = 2 if seed % 2 else {i: (2, 3, 4)[i % 3] for i in range(ni)}
        ct = RefCT(path, ts, dims, output_inds=outs if hyper else None)
        trees.append(dict(ts_inds=ts, dims=dims if isinstance(dims, int) else [dims[i] for i in range(ni)],
                          output_inds=outs if hyper else None, path=path,
                          parent=[-1 if nd.parent is None else nd.parent for nd in ct.nodes],
                          children=[[-1 if c is None else c for c in nd.children] for nd in ct.nodes],
                          inds=[sorted(xs) for xs in ct.inds], ref_path=[list(x) for x in ct.path()],
                          max_width=ct.max_width(), n_inds=ct.n_inds, inds_order=list(ct._inds_order)))
    rng = random.Random(7)
    comps, merges = [], []
    for _ in range(6):
        n = rng.randrange(6, 30)
        ts = [[rng.randrange(n) for _ in range(rng.randrange(1, 4))] for _ in range(n)]
        comps.append(dict(ts_inds=ts, components=[list(c) for c in ref_tn.get_connected_components(ts)]))
    for _ in range(6):
        n = rng.randrange(5, 16)
        order = list(range(n))
        rng.shuffle(order)
        k = rng.randrange(1, 4)
        cuts = sorted(rng.sample(range(1, n), k - 1)) if k > 1 else []
        groups = [sorted(order[i:j]) for i, j in zip([0] + cuts, cuts + [n])]
        paths = []
        for g in groups:   # a random linear path over all n tensors that contracts exactly the tensors of g
            pos = list(range(n))
            live = list(g)
            path = []
            while len(live) > 1:
                x, y = rng.sample(live, 2)
                ix, iy = sorted((pos.index(x), pos.index(y)))
                path.append([ix, iy])
                pos.pop(iy)
                pos.pop(ix)
                new = ('m', len(path), tuple(g))
                pos.append(new)
                live = [t for t in live if t not in (x, y)] + [new]
            paths.append(path)
        merges.append(dict(n=n, paths=paths, merged=[list(x) for x in ref_tn.merge_contraction_paths(n, paths)],
                           merged_noauto=[list(x) for x in
                                          ref_tn.merge_contraction_paths(n, paths, autocomplete=False)]))
    with open(os.path.join(ROOT, 'tests', 'golden', 'host_ctree.json'), 'w') as f:
        json.dump(dict(trees=trees, components=comps, merges=merges), f, separators=(',', ':'))


def wire_format():
    """tests/golden/host_wire.json: the reference's JSON wire format (JSONEncoder of tnco/app/app.py:48-61 and
    tnco/app/{infinite_memory,finite_width}/sa.py:36-60, TensorNetwork.to_json, dump_results) for fixed inputs."""
    from decimal import Decimal

    import tnco.app.finite_width.sa as rfw
    import tnco.app.infinite_memory.sa as rim
    kw = dict(cost='1.11685E+9', runtime_s=1.25, path=[[0, 1], [0, 2], [0, 1]], disconnected_costs=['1.11685E+9'],
              disconnected_paths=[[[0, 1], [0, 2], [0, 1]]])
    kw2 = dict(kw, disconnected_slices=[['a']], slices=['a'])

    def build(cls, k):
        k = dict(k, cost=Decimal(k['cost']), disconnected_costs=[Decimal(x) for x in k['disconnected_costs']],
                 path=[tuple(x) for x in k['path']],
                 disconnected_paths=[[tuple(x) for x in p] for p in k['disconnected_paths']])
        if 'slices' in k:
            k['slices'] = frozenset(k['slices'])
            k['disconnected_slices'] = [frozenset(x) for x in k['disconnected_slices']]
        return cls(**k)

    r_im, r_fw = build(rim.ContractionResults, kw), build(rfw.ContractionResults, kw2)
    rows = [[2, 'x', 'y'], [2, 'y', 'z'], [3, 'z', 'x', '*'], [4, 'x', '/']]
    tns = {}
    for name, rws, opts in (('plain', rows, dict(fuse=False)), ('fused', rows[:3], dict(fuse=4, seed=3))):
        tn = ref_app.load_tn(rws, **opts)
        tns[name] = dict(rows=rws, options=opts, tn_json=tn.to_json(), tags=tn.tags,
                         dump_json=ref_app.dump_results(tn, [r_im, r_im], output_format='json'))
    out = dict(rows=rows, im=dict(kwargs=kw, json=r_im.to_json(), repr=repr(r_im)),
               fw=dict(kwargs=kw2, json=r_fw.to_json(), repr=repr(r_fw)), tn=tns)
    with open(os.path.join(ROOT, 'tests', 'golden', 'host_wire.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
