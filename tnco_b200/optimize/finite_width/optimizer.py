"""Memory-constrained core object (tnco/optimize/finite_width/optimizer.py:54-346 over
finite_width/greedy/optimizer.hpp:72-518): adds ``update(prob, update_slices=True)``, ``slices``, ``min_slices``."""
from __future__ import annotations

from ..infinite_memory.optimizer import Optimizer as _Base


class Optimizer(_Base):
    _finite = True

    def __init__(self, ctree, cmodel, *, slice_update: str = 'greedy', **kwargs):
        if slice_update != 'greedy':
            raise ValueError("'slice_update' must be 'greedy'.")
        self._every = 1
        super().__init__(ctree, cmodel, **kwargs)

    def _set_every(self, every):
        # engine semantics: re-slice after sweep s iff every > 0 and s % every == 0; with every in {0 -> never,
        # 1 -> always} a per-update flag maps onto it without touching the chains
        if every != self._every:
            self._e._chk(self._e._L.tnb_set_update_slices(self._e._h, int(every)))
            self._every = every

    slices = property(lambda s: s._slice_names(False))
    min_slices = property(lambda s: s._slice_names(True))
    skip_slices = property(lambda s: s._skip_slices)

    @staticmethod
    def __build__(*args):
        ctree, cmodel, prng_state, disable_shared_inds, min_ctree, slices, min_slices, skip_slices = args
        return Optimizer(ctree, cmodel, seed=prng_state, disable_shared_inds=disable_shared_inds,
                         skip_slices=skip_slices, _min_ctree=min_ctree, _slices=slices, _min_slices=min_slices)

    def __reduce__(self):
        return self.__build__, (self.ctree, self.cmodel, self.prng_state, self.disable_shared_inds, self.min_ctree,
                                self.slices, self.min_slices, self.skip_slices)
