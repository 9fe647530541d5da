"""Randomised lock-step parity: the engine in TNB_RNG_MT19937 mode against the CPU oracle (itself pinned to the
reference by tests/test_oracle.py) on many small random cases -- hyper / open indices, uniform or per-index
dimensions, unconstrained or a random max_width, different re-slicing periods, bitset widths around the word
boundaries.  Everything is compared bit for bit: trees, total / min costs, slices, index sets, draws consumed."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from helpers import hyper_network, leaf_bits, random_tree, regular_network
from oracle import sa_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu_lib():
    subprocess.check_call(['make', '-C', os.path.join(HERE, 'emu')], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    from tnco_b200 import _lib
    return _lib.bind(ctypes.CDLL(os.path.join(HERE, 'emu', 'libtnb_emu.so')))


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    from tnco_b200 import _lib
    if request.param == 'emu':
        monkeypatch.setattr(_lib, '_LIB', request.getfixturevalue('emu_lib'))
    else:
        monkeypatch.setattr(_lib, '_LIB', None)
    return request.param


def _case(seed):
    rng = random.Random(seed)
    n = rng.choice([4, 6, 10, 16, 22, 30, 44])
    if n % 2:
        n += 1
    hyper = rng.random() < 0.4
    if hyper:
        ts, ni, out = hyper_network(n, seed, n_hyper=rng.randint(1, 5), n_open=rng.randint(0, 3))
    else:
        ts, ni = regular_network(n, seed)
        out = []
    kind = rng.choice(['d2', 'd2', 'd3', 'pow2'])
    dim, dims = 2, None
    if kind == 'd3':
        dim = 3
    elif kind == 'pow2':
        dims = np.array([rng.choice([2, 2, 4, 8]) for _ in range(ni)], np.uint64)
    p, a, b, bits = random_tree(ts, ni, seed + 7, out)
    mw = None
    if rng.random() < 0.6:
        l2 = np.log2(dims.astype(float)) if dims is not None else np.full(ni, np.log2(dim))
        w0 = max(sum(l2[i] for i in range(ni) if (int(row[i >> 5]) >> (i & 31)) & 1) for row in bits)
        mw = float(int(w0 * rng.uniform(0.3, 0.9)))
    return dict(n=n, ni=ni, out=out, dim=dim, dims=dims, tree=(p, a, b), bits=bits, mw=mw,
                every=rng.choice([1, 3, 10]), run_seed=rng.randrange(2**32), n_sweeps=rng.choice([40, 120]))


@pytest.mark.parametrize('seed', range(int(os.environ.get('TNB_FUZZ_CASES', '40'))))
def test_random_case_matches_oracle_bit_for_bit(backend, seed):
    from tnco_b200.engine import RNG_MT19937, Engine, pack_index_set
    cs = _case(1000 + seed)
    p, a, b = cs['tree']
    n, ni, mw, every, S = cs['n'], cs['ni'], cs['mw'], cs['every'], cs['n_sweeps']
    betas = [30.0 * s / S for s in range(S)]
    oc = so.Chain(p, a, b, cs['bits'], ni, dim=cs['dim'], dims=cs['dims'], max_width=mw, seed=cs['run_seed'])
    e = Engine()
    e.set_network(cs['bits'][:n], ni, dim=cs['dim'], dims=cs['dims'], output_bits=pack_index_set(cs['out'], ni))
    e.set_mode(max_width=mw, update_slices_every=every, rng=RNG_MT19937)
    e.set_chains(p[None], a[None], b[None], [cs['run_seed']])
    e.set_betas(betas)
    assert (e.bits(0) == cs['bits']).all()
    assert e.costs()[0][0] == oc.total_cost
    if mw is not None:
        assert (e.slices()[0] == oc.slices()).all()
    done = 0
    for upto in (S // 3, S):
        for s in range(done, upto):
            oc.update(betas[s], update_slices=(s % every == 0))
        done = upto
        e.run(upto)
        for x, y in zip(e.trees(), oc.tree()):
            assert (x[0] == y).all()
        for x, y in zip(e.trees(True), oc.tree(True)):
            assert (x[0] == y).all()
        t, m = e.costs()
        assert t[0] == oc.total_cost and m[0] == oc.min_total_cost
        assert (e.bits(0) == oc.bits()).all()
        if mw is not None:
            assert (e.slices()[0] == oc.slices()).all() and (e.slices(True)[0] == oc.slices(True)).all()
    pr, c = e.progress(), oc.counters()
    assert pr['proposals'][0] == c['proposals'] and pr['accepts'][0] == c['accepts']
    assert pr['words'][0] == c['words_drawn']
    e.close()
