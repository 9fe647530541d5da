"""Application layer of the SA path: the plugin surface ``tnco.app.Optimizer`` resolves to.

Mirrors tnco/app/app.py: ``BaseOptimizer`` (:715-795, same dataclass fields, plus engine knobs appended with
defaults), ``BaseContractionResults`` (:64-94), ``dump_results`` (:573-712) and the structure-only subset of
``load_tn`` (:154-570: TensorNetwork objects, lists / strings of indices, index-level ``fuse``).  Circuit front-ends
and hyper-index decomposition (both need tensor data) are out of scope.
"""
from __future__ import annotations

import bz2
import gzip
import io
import json
import pickle
import re
from dataclasses import dataclass
from decimal import Decimal
from pathlib import Path
from random import Random
from typing import Any
from warnings import warn

from ..tn import Tensor, TensorNetwork, read_inds
from ..tn import contract as tn_contract
from ..tn import fuse as tn_fuse

__all__ = ['BaseOptimizer', 'BaseContractionResults', 'load_tn', 'dump_results', 'JSONEncoder']


class JSONEncoder(json.JSONEncoder):

    def default(self, obj):
        if isinstance(obj, Decimal):
            return str(obj)
        if isinstance(obj, frozenset):
            return tuple(obj)
        if isinstance(obj, BaseContractionResults):
            return dict(cost=obj.cost, runtime_s=obj.runtime_s, path=obj.path)
        if hasattr(obj, 'to_json'):
            return obj.to_json()
        return super().default(obj)


class _LazyFields:
    """Result records keep the reference's attribute names (tnco/app/app.py:64-94, sa.py:63-90) but a field may
    be given as a zero-argument callable: it is evaluated and cached on first access.  optimize() hands out
    tens of thousands of results whose paths live in numpy arrays until somebody looks at them."""
    _fields = ()

    def __init__(self, *args, **kwargs):
        names = self._fields
        if len(args) > len(names):
            raise TypeError('too many positional arguments')
        vals = dict(zip(names, args))
        for k, v in kwargs.items():
            if k not in names or k in vals:
                raise TypeError(f'unexpected or duplicate argument {k!r}')
            vals[k] = v
        missing = [k for k in names if k not in vals]
        if missing:
            raise TypeError(f'missing arguments: {missing}')
        for k in names:
            object.__setattr__(self, '_' + k, vals[k])

    @classmethod
    def _from_source(cls, source, key):
        """Record whose every field is ``source(field_name, key)``, evaluated and cached on first access (one
        shared callable instead of per-record closures: optimize() creates one record per run)."""
        o = object.__new__(cls)
        o.__dict__['_src'] = (source, key)   # (not object.__setattr__: half the cost, times tens of thousands of records)
        return o

    def __setattr__(self, k, v):
        raise AttributeError('results are read-only')

    def _get(self, k):
        d = object.__getattribute__(self, '__dict__')
        if '_' + k in d:
            v = d['_' + k]
            if callable(v):
                v = v()
                d['_' + k] = v
            return v
        source, key = d['_src']
        v = d['_' + k] = source(k, key)
        return v

    def __getattr__(self, k):
        if k in type(self)._fields:
            return self._get(k)
        raise AttributeError(k)

    def __getstate__(self):
        return {k: self._get(k) for k in self._fields}

    def __setstate__(self, state):
        for k, v in state.items():
            object.__setattr__(self, '_' + k, v)


class BaseContractionResults(_LazyFields):
    """cost, runtime_s, path (linear einsum format) -- tnco/app/app.py:64-94."""
    _fields = ('cost', 'runtime_s', 'path')

    def __lt__(self, other):
        if not isinstance(other, BaseContractionResults):
            raise ValueError("Cannot compare against '{}'.".format(type(other).__name__))
        return self.cost < other.cost

    def __repr__(self):
        return 'ContractionResults(cost={:1.3g}, runtime={:1.3g}s)'.format(self.cost, self.runtime_s)

    def to_json(self):
        return json.dumps(self, cls=JSONEncoder)


def cost_to_decimal(x: float) -> Decimal:
    """The reference prints costs through ``ostringstream << double`` (6 significant digits, %g) and parses
    the text as Decimal (include/tnco/globals.hpp:48-53, infinite_memory/optimizer.hpp:278-289)."""
    return Decimal('%.6g' % float(x))


def load_tn(obj: Any, *, output_index_token: str = '*', sparse_index_token: str = '/', **options) -> TensorNetwork:
    """Structure-only ``load_tn`` (tnco/app/app.py:154-570): TensorNetwork, list of ``(dim, tensor names...)`` rows
    (one row per index), or the same as text.  As in the reference, tensors are pre-merged (``fuse=4``: while the
    merged tensor has width <= 4) unless ``fuse=False``; ``decompose_hyper_inds`` needs arrays and is a no-op here.

    >>> tn = load_tn([[2, 'i', 'j'], [2, 'j', 'k']], fuse=False)
    >>> len(tn)
    3
    >>> load_tn([[2, 'i', 'j'], [2, 'j', 'k']], fuse=4, decompose_hyper_inds=False, seed=0).tags['fuse_path']
    [(1, 2), (0, 1)]
    """
    if isinstance(obj, TensorNetwork):
        fuse = options.pop('fuse', 4)
        decompose = options.pop('decompose_hyper_inds', True)
        seed = options.pop('seed', None)
        for k in ('simplify_circuit', 'initial_state', 'final_state', 'atol', 'dtype', 'backend', 'verbose'):
            options.pop(k, None)
        if options:
            raise TypeError('Got unexpected keyword arguments: {}'.format(sorted(options)))
        ts_inds, dims, tags, ts_tags = list(obj.ts_inds), obj.dims, dict(obj.tags), list(obj.ts_tags)
        output_inds, sparse_inds = obj.output_inds, obj.sparse_inds
        if sparse_inds:  # app.py:322-326
            warn('The decomposition of hyper-indices and the fusion of indices is not yet supported if there are '
                 'sparse indices')
            decompose = fuse = False
        if decompose:  # app.py:330-335: needs the arrays, which a structure-only network never has
            warn('Cannot decompose hyper-indices if not all arrays are provided.')
        if fuse is not None and fuse > 0:  # app.py:373-414
            path = tn_fuse(ts_inds, dims, max_width=fuse, output_inds=output_inds, seed=seed)
            ts_inds, output_inds = tn_contract(path, ts_inds, output_inds, dims=dims)
            for px, py in map(sorted, path):
                ty, tx = ts_tags.pop(py), ts_tags.pop(px)
                ts_tags.append((tx or None) if not ty else ty if not tx else dict(x=tx, y=ty))
            if 'fuse_path' in tags:
                raise ValueError("'TensorNetwork' has already the tag 'fuse_path'.")
            tags['fuse_path'] = path
        else:
            return obj
        return TensorNetwork((Tensor(xs, [dims[x] for x in xs], tags=t) for xs, t in zip(ts_inds, ts_tags)),
                             output_inds=output_inds, sparse_inds=sparse_inds, tags=tags)
    if isinstance(obj, str):
        rows = []
        for line in obj.splitlines():
            line = re.sub(r'\s+', ' ', line).strip()
            if not line or line.startswith('#'):
                continue
            if not re.match(r'\d+(\s+\S+)*\s*$', line):
                raise TypeError("'obj' is not recognized.")
            d, *xs = line.split()
            rows.append((int(d), *xs))
        return load_tn(rows, output_index_token=output_index_token, sparse_index_token=sparse_index_token, **options)

    def is_int(x):
        try:
            return int(x) == x
        except (TypeError, ValueError):
            return False

    try:
        rows = list(obj)
        ok = all(hasattr(x, '__getitem__') and len(x) > 1 and is_int(x[0]) for x in rows) and len(rows) > 0
    except TypeError:
        ok = False
    if ok:
        tensor_map, dims, output_inds, sparse_inds = read_inds(dict(enumerate(rows)),
                                                               output_index_token=output_index_token,
                                                               sparse_index_token=sparse_index_token)
        return load_tn(TensorNetwork((Tensor(xs, [dims[x] for x in xs], tags=dict(name=name))
                                      for name, xs in tensor_map.items()),
                                     output_inds=output_inds, sparse_inds=sparse_inds), **options)
    raise TypeError("'obj' is not recognized.")


def dump_results(tn, res, *, output_format=None, output_filename=None, output_compression='auto',
                 overwrite_output_file=False, **kwargs):
    """tnco/app/app.py:573-712: returns ``(tn, res)`` (raw) or a JSON string, or writes them to a file."""
    check_only = kwargs.pop('check_only', False)
    if kwargs:
        raise TypeError('Unexpected extra keyword arguments.')
    output_format = 'raw' if output_format is None else str(output_format).lower()
    if output_format not in ['raw', 'json']:
        raise ValueError(f'"{output_format=}" not supported.')
    output_filename = None if output_filename is None else Path(output_filename).expanduser()
    if output_filename and not overwrite_output_file and output_filename.exists():
        raise FileExistsError(
            "'{}' already exists. Please use 'overwrite_output_file=True'.".format(output_filename))
    output_compression = str(output_compression).lower()
    if output_compression not in ['auto', 'none', 'bz2', 'gzip']:
        raise ValueError(f'"{output_compression=}" not supported.')
    if check_only:
        return None
    output = (tn, res)
    if output_format == 'json':
        output = '{{"tn" : {}, "res" : {}}}'.format(tn.to_json(), '[' + ', '.join(r.to_json() for r in res) + ']')
    if output_filename:
        suffix = output_filename.suffix[1:] if output_compression == 'auto' else output_compression
        open_, compress = (gzip.open, True) if suffix == 'gzip' else (bz2.open, True) if suffix == 'bz2' else (io.open, False)
        if isinstance(output, str):
            with open_(output_filename, 'w') as f:
                f.write(output.encode() if compress else output)
            return None
        with open_(output_filename, 'w' if compress else 'bw') as f:
            pickle.dump(output, f)
        return None
    return output


@dataclass(frozen=True)
class BaseOptimizer:
    """Same fields, defaults and order as tnco/app/app.py:755-767; engine knobs follow."""
    max_width: float | None = None
    n_jobs: int = -1          # accepted for compatibility; runs are batched on the GPU instead of processes
    width_type: str = 'float32'
    cost_type: str = 'float64'
    output_format: str | None = None
    output_filename: str | None = None
    output_compression: str = 'auto'
    overwrite_output_file: bool = False
    atol: float = 1e-5
    dtype: Any | None = None
    backend: str | None = None
    seed: int | None = None
    verbose: int = False
    # --- tnco_b200 extras
    rng: str = 'philox'        # 'philox' (production) | 'mt19937' (bit-identical to the reference per seed)
    init_trees: str = 'greedy'  # 'greedy' | 'random' initial contraction trees
    tree_builder: str = 'device'  # 'device' (built by each chain's own lanes) | 'host' (C++ threads)
    device: int | None = None  # CUDA device; default LOCAL_RANK or 0
    distributed: bool = True   # shard runs over torch.distributed ranks when a process group exists
    sync_every: int | None = None  # sweeps between min-reductions of the best cost over ranks (None: only at the end)
    # torch.distributed: best trees stay on the rank that ran them; the k best of all ranks are gathered everywhere
    # (costs of ALL runs always are).  'all' gathers every tree on every rank, like the reference's single process.
    gather_paths: Any = 64

    def optimize(self, *args, **kwargs):
        raise NotImplementedError()

    def _load_tn(self, tn, **load_tn_options):
        return load_tn(tn, atol=self.atol, dtype=self.dtype, backend=self.backend, seed=self.seed,
                       verbose=self.verbose, **load_tn_options)

    def _dump_results(self, tn, res, **opts):
        return dump_results(tn, res, output_format=self.output_format, output_filename=self.output_filename,
                            output_compression=self.output_compression,
                            overwrite_output_file=self.overwrite_output_file, **opts)

    def __post_init__(self):
        object.__setattr__(self, '_rng', Random(self.seed))
        if self.width_type != 'float32' or self.cost_type != 'float64':
            raise ValueError("tnco_b200 computes costs in float64 and widths in float32 (the reference's "
                             "defaults, tnco/app/app.py:757-758); other types are not available.")
        if self.rng not in ('philox', 'mt19937'):
            raise ValueError("'rng' must be 'philox' or 'mt19937'.")
        if self.init_trees not in ('greedy', 'random'):
            raise ValueError("'init_trees' must be 'greedy' or 'random'.")
        if self.tree_builder not in ('device', 'host'):
            raise ValueError("'tree_builder' must be 'device' or 'host'.")
        if self.gather_paths != 'all' and (not isinstance(self.gather_paths, int) or self.gather_paths < 1):
            raise ValueError("'gather_paths' must be a positive number or 'all'.")
        self._dump_results(None, None, check_only=True)

    def __getstate__(self):
        return {k: getattr(self, k) for k in self.__dataclass_fields__}

    def __setstate__(self, state):
        for k, v in state.items():
            object.__setattr__(self, k, v)
        object.__setattr__(self, '_rng', Random(self.seed))
