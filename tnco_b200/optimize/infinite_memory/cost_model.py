"""``SimpleCostModel`` without memory constraint (tnco/optimize/infinite_memory/cost_model.py,
include/tnco/optimize/infinite_memory/cost_model/simple.hpp:38-83): cost = prod of dims over ``inds_a | inds_b``;
with ``sparse_inds`` / ``n_projs`` the sparse-index model (cost_model/simple_sparse_inds.hpp:38-49):
cost = prod of the dense dims * min(prod of the sparse dims, n_projs)."""
from __future__ import annotations

import math


def check_sparse(sparse_inds, n_projs):
    """Argument rules of tnco/optimize/infinite_memory/cost_model.py:79-89."""
    if n_projs is not None and (n_projs != int(n_projs) or n_projs < 1):
        raise ValueError("'n_projs' must be a positive number.")
    sparse_inds = None if sparse_inds is None else frozenset(sparse_inds)
    if sparse_inds is None and n_projs:
        raise ValueError("'n_projs' cannot be specified if 'sparse_inds' is not provided.")
    if sparse_inds and not n_projs:
        raise ValueError("'n_projs' must be specified if 'sparse_inds' is provided.")
    return sparse_inds, None if n_projs is None else int(n_projs)


def _prod(xs, dims):
    try:
        return float(math.prod(dims[x] for x in xs))
    except TypeError:
        return float(dims)**len(xs)


class SimpleCostModel:

    def __init__(self, cost_type: str = 'float64', sparse_inds=None, n_projs=None):
        if cost_type != 'float64':
            raise ValueError("tnco_b200 computes costs in float64 only.")
        self.sparse_inds, self.n_projs = check_sparse(sparse_inds, n_projs)
        self.cost_type = cost_type

    def _cost(self, xs, dims):
        if not self.sparse_inds:
            return _prod(xs, dims)
        return _prod(xs - self.sparse_inds, dims) * min(_prod(xs & self.sparse_inds, dims), float(self.n_projs))

    def contraction_cost(self, inds_a, inds_b, inds_out=None, dims=2):
        return self._cost(frozenset(inds_a) | frozenset(inds_b), dims)

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        r = 'SimpleCostModel(cost_type={}'.format(self.cost_type)
        if self.sparse_inds is not None:
            r += ', sparse_inds={}, n_projs={}'.format(self.sparse_inds, self.n_projs)
        return r + ')'
