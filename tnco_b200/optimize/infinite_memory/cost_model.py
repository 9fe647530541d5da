"""``SimpleCostModel`` without memory constraint (tnco/optimize/infinite_memory/cost_model.py,
include/tnco/optimize/infinite_memory/cost_model/simple.hpp:38-83): cost = prod of dims over ``inds_a | inds_b``."""
from __future__ import annotations

import math


class SimpleCostModel:

    def __init__(self, cost_type: str = 'float64', sparse_inds=None, n_projs=None):
        if cost_type != 'float64':
            raise ValueError("tnco_b200 computes costs in float64 only.")
        if sparse_inds or n_projs is not None:
            raise NotImplementedError('tnco_b200: sparse indices are not supported yet.')
        self.cost_type = cost_type

    def contraction_cost(self, inds_a, inds_b, inds_out=None, dims=2):
        xs = frozenset(inds_a) | frozenset(inds_b)
        try:
            return float(math.prod(dims[x] for x in xs))
        except TypeError:
            return float(dims)**len(xs)

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        return 'SimpleCostModel(cost_type={})'.format(self.cost_type)
