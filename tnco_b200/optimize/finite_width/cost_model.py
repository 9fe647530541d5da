"""``SimpleCostModel`` with a maximum width (tnco/optimize/finite_width/cost_model.py,
include/tnco/optimize/finite_width/cost_model/simple.hpp:39-145): width = sum of log2 dims, cost over
``inds_a | inds_b | slices``."""
from __future__ import annotations

import math

import numpy as np


class SimpleCostModel:

    def __init__(self, max_width: float, width_type: str = 'float32', cost_type: str = 'float64', sparse_inds=None,
                 n_projs=None):
        if max_width < 0:
            raise ValueError("'max_width' must be a non-negative number.")
        if cost_type != 'float64' or width_type != 'float32':
            raise ValueError("tnco_b200 computes costs in float64 and widths in float32 only.")
        if sparse_inds or n_projs is not None:
            raise NotImplementedError('tnco_b200: sparse indices are not supported yet.')
        self.max_width = float(np.float32(max_width))
        self.width_type, self.cost_type = width_type, cost_type

    def width(self, inds, dims=2):
        try:
            return float(sum(math.log2(dims[x]) for x in inds))
        except TypeError:
            return float(np.float32(math.log2(dims) * len(frozenset(inds))))

    def contraction_cost(self, inds_a, inds_b, inds_out=None, dims=2, slices=()):
        xs = frozenset(inds_a) | frozenset(inds_b) | frozenset(slices)
        try:
            return float(math.prod(dims[x] for x in xs))
        except TypeError:
            return float(dims)**len(xs)

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        return 'SimpleCostModel(max_width={}, width_type={}, cost_type={})'.format(self.max_width, self.width_type,
                                                                                  self.cost_type)
