"""``Optimizer(ctree, cmodel, seed=...)`` / ``update(prob)`` -- the reference's core object
(tnco/optimize/infinite_memory/optimizer.py:47-245 over infinite_memory/optimizer.hpp:61-310) as ONE chain of the
GPU engine in TNB_RNG_MT19937 mode: for an integer seed every update() consumes the same std::mt19937 draws in
the same order as the reference, so trees, costs and ``prng_state`` are identical to the reference's."""
from __future__ import annotations

import math
import secrets
from decimal import Decimal

import numpy as np

from ..._lib import RNG_MT19937
from ...ctree import ContractionTree
from ...engine import (Engine, mt19937_advance_str, mt19937_state_str, pack_index_set, parse_mt19937_state,
                       unpack_bits)
from ..prob import BaseProbability


def _dec(x):
    return Decimal('%.6g' % float(x))


class Optimizer:
    _finite = False

    def __init__(self, ctree: ContractionTree, cmodel, *, seed=None, disable_shared_inds: bool = False,
                 atol: float = 1e-5, device: int = 0, **kwargs):
        # finite-width core object only (tnco/optimize/finite_width/optimizer.py:59,122-123)
        self._max_new = int(kwargs.pop('max_number_new_slices', 0) or 0)
        if self._max_new < 0 or (self._max_new and not self._finite):
            raise TypeError('Got unexpected keyword arguments.' if not self._finite else
                            "'max_number_new_slices' must be a non-negative number.")
        self._skip_slices = frozenset(kwargs.pop('skip_slices', None) or ())
        kwargs.pop('slice_update', None)
        # what unpickling hands back (optimizer.py:234-247, finite_width/optimizer.py:330-346)
        min_ctree = kwargs.pop('_min_ctree', None)
        slices, min_slices = kwargs.pop('_slices', None), kwargs.pop('_min_slices', None)
        if kwargs:
            raise TypeError('Got unexpected keyword arguments.')
        # seed: an integer, or a std::mt19937 state string as `prng_state` returns it (optimize/optimizer.hpp:68-72)
        self._state0 = parse_mt19937_state(seed) if isinstance(seed, str) else None
        self._seed = 0 if self._state0 is not None else secrets.randbits(32) if seed is None else int(seed) % 2**32
        self._ctree0, self._cmodel = ctree, cmodel
        self._dsi, self._atol = bool(disable_shared_inds), atol
        dims = [int(ctree.dims[x]) for x in ctree._inds_order]
        uniform = len(set(dims)) == 1
        lb, ni = ctree.leaf_bits()
        pos = {x: k for k, x in enumerate(ctree._inds_order)}
        self._e = Engine(device)
        sparse = getattr(cmodel, 'sparse_inds', None)
        if sparse and not frozenset(sparse).issubset(pos):  # tnco/optimize/infinite_memory/cost_model.py:190-194
            raise ValueError("Sparse indices are not a subset of 'inds_order'.")
        self._e.set_network(lb, ni, dim=dims[0], dims=None if uniform else dims,
                            output_bits=pack_index_set([pos[x] for x in ctree.output_inds()], ni),
                            sparse_bits=pack_index_set([pos[x] for x in sparse], ni) if sparse else None,
                            n_projs=getattr(cmodel, 'n_projs', None))
        if self._finite:  # tnco/optimize/finite_width/optimizer.py:96-107
            if not self._skip_slices.issubset(pos):
                raise ValueError("'skip_slices' must be a subset of available indices.")
            if max(cmodel.width(frozenset(xs) & self._skip_slices, ctree.dims) for xs in ctree.inds) > cmodel.max_width:
                raise ValueError("Too many indices in 'skip_slices'.")
            if self._skip_slices:
                self._e.set_skip_slices(pack_index_set([pos[x] for x in self._skip_slices], ni))
        self._e.set_mode(max_width=getattr(cmodel, 'max_width', None) if self._finite else None,
                         update_slices_every=1, disable_shared_inds=self._dsi, rng=RNG_MT19937)
        if self._max_new:
            self._e.set_new_slices(self._max_new)
        p, a, b = ctree.arrays()
        self._e.set_chains(p[None], a[None], b[None], [self._seed])
        if self._state0 is not None or min_ctree is not None or slices is not None or min_slices is not None:
            def bits(names):
                return None if names is None else pack_index_set([pos[x] for x in names], ni)[None]
            best = None if min_ctree is None else tuple(x[None] for x in min_ctree.arrays())
            self._e.set_resume(mt_state=None if self._state0 is None else self._state0[None],
                               slices=bits(slices) if self._finite else None, best_trees=best,
                               best_slices=bits(min_slices) if self._finite else None)
        self._n_updates = 0
        self._e.costs()  # construct caches now: invalid input / "Precision is too low." raise here (ValueError)

    # ---- stepping
    def update(self, prob: BaseProbability, update_slices: bool = True):
        self._e.set_prob(prob.kind)
        # the engine re-slices on sweeps s with s % every == 0: park the sweep index on / off that grid
        self._e.set_betas([getattr(prob, 'beta', 0.0)])
        if self._finite:
            self._set_every(1 if update_slices else 0)
        self._n_updates += 1
        self._e.run(self._n_updates)

    def _set_every(self, every):
        pass

    # ---- state
    def _tree(self, best):
        p, a, b = self._e.trees(best=best, chain0=0, n=1)
        c = self._ctree0
        # (the ORDERED leaf tuples: rebuilding from the frozensets would let the index bit positions -- which feed the
        #  slicer's tie-breaks -- depend on PYTHONHASHSEED for string indices, and a pickled object would resume
        #  differently from the uninterrupted run)
        return ContractionTree.from_arrays(p[0], a[0], b[0], list(c._leaves), c.dims, output_inds=c.output_inds())

    ctree = property(lambda s: s._tree(False))
    min_ctree = property(lambda s: s._tree(True))
    cmodel = property(lambda s: s._cmodel)
    disable_shared_inds = property(lambda s: s._dsi)
    total_cost = property(lambda s: _dec(s._e.costs()[0][0]))
    min_total_cost = property(lambda s: _dec(s._e.costs()[1][0]))
    log2_total_cost = property(lambda s: math.log2(s._e.costs()[0][0]))
    log2_min_total_cost = property(lambda s: math.log2(s._e.costs()[1][0]))

    @property
    def prng_state(self) -> str:
        words = int(self._e.progress()['words'][0])
        if self._state0 is not None:
            return mt19937_advance_str(self._state0, words)
        return mt19937_state_str(self._seed, words)

    # ---- pickling: rebuild through the constructor from the current state, like the reference
    @staticmethod
    def __build__(*args):
        ctree, cmodel, prng_state, disable_shared_inds, min_ctree = args
        return Optimizer(ctree, cmodel, seed=prng_state, disable_shared_inds=disable_shared_inds,
                         _min_ctree=min_ctree)

    def __reduce__(self):
        return self.__build__, (self.ctree, self.cmodel, self.prng_state, self.disable_shared_inds, self.min_ctree)

    def __eq__(self, other):
        return isinstance(other, Optimizer) and self.__reduce__()[1] == other.__reduce__()[1]

    __hash__ = None

    def __repr__(self):
        return 'Optimizer(ctree={}, cmodel={})'.format(self.ctree, self.cmodel)

    def is_valid(self, *, atol: float = 1e-5, return_message: bool = False):
        """Re-derive the cached totals from the trees (infinite_memory/optimizer.hpp:223-252)."""
        t, m = self._e.costs()
        p, a, b = self._e.trees()
        bp, ba, bb = self._e.trees(best=True)
        sl = self._e.slices() if self._finite else None
        bsl = self._e.slices(best=True) if self._finite else None
        _, pc, mw = self._e.eval_cost(p, a, b, slices=sl)
        bseq, _, bmw = self._e.eval_cost(bp, ba, bb, slices=bsl)
        ok, msg = True, ''
        if abs(math.log(pc[0]) - math.log(t[0])) > atol:
            ok, msg = False, 'CostCache is not properly cached.'
        elif abs(math.log(bseq[0]) - math.log(m[0])) > atol:
            ok, msg = False, 'Cost for min ctree is not correct.'
        elif self._finite and (mw[0] > self._cmodel.max_width + atol or bmw[0] > self._cmodel.max_width + atol):
            ok, msg = False, 'Width larger than allowed width after slicing.'
        return (ok, msg) if return_message else ok

    def _slice_names(self, best):
        order = self._ctree0._inds_order
        return frozenset(order[i] for i in unpack_bits(self._e.slices(best=best)[0]))
