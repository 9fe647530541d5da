"""Simulated annealing without memory constraint -- the plugin ``tnco.app.infinite_memory.sa``
(tnco/app/infinite_memory/sa.py:63-257) on the B200 engine."""
from __future__ import annotations

import json
from typing import Any, Iterable

from .. import _sa
from ..app import BaseContractionResults, BaseOptimizer
from ..app import JSONEncoder as BaseJSONEncoder


class JSONEncoder(BaseJSONEncoder):

    def default(self, obj):
        if isinstance(obj, ContractionResults):
            return dict(**BaseJSONEncoder().default(obj), disconnected_paths=obj.disconnected_paths)
        return super().default(obj)


class ContractionResults(BaseContractionResults):
    """Fields as in tnco/app/infinite_memory/sa.py (dataclass there; lazily materialised record here)."""
    _fields = BaseContractionResults._fields + ('disconnected_costs', 'disconnected_paths')

    def to_json(self):
        return json.dumps(self, cls=JSONEncoder)


class Optimizer(BaseOptimizer):
    """``optimize(tn, betas, n_steps=None, n_runs=1, n_projs=None, timeout=None, **load_tn_options)``."""

    def optimize(self, tn: Any, betas: tuple[float, float] | Iterable[float], n_steps: int | None = None,
                 n_runs: int = 1, n_projs: int | None = None, timeout: float | None = None,
                 **load_tn_options) -> Any:
        return _sa.optimize(self, ContractionResults, tn, betas, n_steps, n_runs, n_projs, 0, timeout, False,
                            load_tn_options)
