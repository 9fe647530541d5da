"""Acceptance rules (tnco/optimize/prob.py, include/tnco/optimize/prob/{base,greedy,mh}.hpp).

The objects select the rule the engine applies on the device; ``__call__`` evaluates the same formula on the host
for inspection (the reference exposes it the same way).

>>> MetropolisHastings(beta=1)(-10, 100)
1.0
"""
from __future__ import annotations

from .._lib import PROB_ALWAYS, PROB_GREEDY, PROB_MH

__all__ = ['BaseProbability', 'Greedy', 'MetropolisHastings']


class BaseProbability:
    kind = PROB_ALWAYS

    def __init__(self, cost_type: str = 'float64'):
        self.cost_type = cost_type

    def __call__(self, delta_cost, old_cost):
        return 1.0

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        return '{}(cost_type={})'.format(type(self).__name__, self.cost_type)


class Greedy(BaseProbability):
    kind = PROB_GREEDY

    def __call__(self, delta_cost, old_cost):
        return 1.0 if delta_cost <= 0 else 0.0


class MetropolisHastings(BaseProbability):
    kind = PROB_MH

    def __init__(self, beta: float = 0, cost_type: str = 'float64'):
        super().__init__(cost_type)
        self.beta = float(beta)

    def __call__(self, delta_cost, old_cost):
        if delta_cost <= 0:
            return 1.0
        if old_cost == 0:
            return 0.0
        return float(pow(1 + (delta_cost / old_cost), -self.beta))

    def __repr__(self):
        return 'MetropolisHastings(beta={}, cost_type={})'.format(self.beta, self.cost_type)
