"""``SimpleCostModel`` with a maximum width (tnco/optimize/finite_width/cost_model.py,
include/tnco/optimize/finite_width/cost_model/simple.hpp:39-145): width = sum of log2 dims, cost over
``inds_a | inds_b | slices``; with ``sparse_inds`` / ``n_projs`` the sparse-index model
(finite_width/cost_model/simple_sparse_inds.hpp:38-157): the sparse part of a width is capped at log2(n_projs) and
the sparse part of a cost at n_projs."""
from __future__ import annotations

import math

import numpy as np

from ..infinite_memory.cost_model import _prod, check_sparse


class SimpleCostModel:

    def __init__(self, max_width: float, width_type: str = 'float32', cost_type: str = 'float64', sparse_inds=None,
                 n_projs=None):
        if max_width < 0:
            raise ValueError("'max_width' must be a non-negative number.")
        if cost_type != 'float64' or width_type != 'float32':
            raise ValueError("tnco_b200 computes costs in float64 and widths in float32 only.")
        self.sparse_inds, self.n_projs = check_sparse(sparse_inds, n_projs)
        self.max_width = float(np.float32(max_width))
        self.width_type, self.cost_type = width_type, cost_type

    @staticmethod
    def _width(inds, dims):
        try:
            return float(sum(math.log2(dims[x]) for x in inds))
        except TypeError:
            return float(np.float32(math.log2(dims) * len(frozenset(inds))))

    def width(self, inds, dims=2):
        xs = frozenset(inds)
        if not self.sparse_inds:
            return self._width(xs, dims)
        return self._width(xs - self.sparse_inds, dims) + min(self._width(xs & self.sparse_inds, dims),
                                                              math.log2(self.n_projs))

    def contraction_cost(self, inds_a, inds_b, inds_out=None, dims=2, slices=()):
        xs = frozenset(inds_a) | frozenset(inds_b) | frozenset(slices)
        if not self.sparse_inds:
            return _prod(xs, dims)
        return _prod(xs - self.sparse_inds, dims) * min(_prod(xs & self.sparse_inds, dims), float(self.n_projs))

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        r = 'SimpleCostModel(max_width={}, width_type={}, cost_type={}'.format(self.max_width, self.width_type,
                                                                               self.cost_type)
        if self.sparse_inds is not None:
            r += ', sparse_inds={}, n_projs={}'.format(self.sparse_inds, self.n_projs)
        return r + ')'
