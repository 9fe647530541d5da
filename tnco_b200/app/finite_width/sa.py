"""Simulated annealing under a maximum tensor width (sliced indices) -- the plugin
``tnco.app.finite_width.sa`` (tnco/app/finite_width/sa.py:73-289) on the B200 engine."""
from __future__ import annotations

import json
from typing import Any, Iterable

from .. import _sa
from ..app import BaseContractionResults, BaseOptimizer
from ..app import JSONEncoder as BaseJSONEncoder


class JSONEncoder(BaseJSONEncoder):

    def default(self, obj):
        if isinstance(obj, ContractionResults):
            return dict(**BaseJSONEncoder().default(obj), disconnected_paths=obj.disconnected_paths,
                        disconnected_slices=obj.disconnected_slices, slices=obj.slices)
        return super().default(obj)


class ContractionResults(BaseContractionResults):
    """Fields as in tnco/app/finite_width/sa.py (dataclass there; lazily materialised record here)."""
    _fields = BaseContractionResults._fields + ('disconnected_costs', 'disconnected_paths', 'disconnected_slices', 'slices')

    def to_json(self):
        return json.dumps(self, cls=JSONEncoder)


class Optimizer(BaseOptimizer):
    """``optimize(tn, betas, n_steps=None, n_runs=1, n_projs=None, update_slices=10, timeout=None, **opts)``."""

    def optimize(self, tn: Any, betas: tuple[float, float] | Iterable[float], n_steps: int | None = None,
                 n_runs: int = 1, n_projs: int | None = None, update_slices: int = 10,
                 timeout: float | None = None, **load_tn_options) -> Any:
        if int(update_slices) != update_slices or update_slices < 1:
            raise ValueError("'update_slices' must be a positive number.")
        return _sa.optimize(self, ContractionResults, tn, betas, n_steps, n_runs, n_projs, int(update_slices),
                            timeout, True, load_tn_options)
