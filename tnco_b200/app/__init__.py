"""``tnco_b200.app`` -- drop-in for the SA path of ``tnco.app`` (tnco/app/__init__.py, app.py:798-878)."""
from __future__ import annotations

from importlib import import_module
from typing import Any

from ..tn import Tensor, TensorNetwork
from .app import BaseOptimizer, dump_results, load_tn

__all__ = ['Optimizer', 'Tensor', 'TensorNetwork', 'load_tn', 'dump_results']


def Optimizer(method: str = 'sa', max_width: float | None = None, n_jobs: int = -1, width_type: str = 'float32',
              cost_type: str = 'float64', output_format: str | None = None, output_filename: str | None = None,
              output_compression: str = 'auto', overwrite_output_file: bool = False, atol: float = 1e-5,
              dtype: Any | None = None, backend: str | None = None, seed: int | None = None, verbose: int = False,
              **engine_options) -> BaseOptimizer:
    """Factory with the reference's signature (tnco/app/app.py:798-811): picks ``finite_width`` when
    ``max_width`` is finite, else ``infinite_memory``, then the module named ``method``.

    >>> from tnco_b200.app import Optimizer
    >>> opt = Optimizer(method='sa')
    """
    opts = dict(max_width=max_width, n_jobs=n_jobs, width_type=width_type, cost_type=cost_type,
                output_format=output_format, output_filename=output_filename,
                output_compression=output_compression, overwrite_output_file=overwrite_output_file, atol=atol,
                dtype=dtype, backend=backend, seed=seed, verbose=verbose, **engine_options)
    module = 'tnco_b200.app'
    module += '.finite_width' if (max_width is not None and max_width < float('inf')) else '.infinite_memory'
    module += '.' + str(method)
    return import_module(module).Optimizer(**opts)
