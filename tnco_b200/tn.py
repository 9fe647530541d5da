"""Minimal tensor-network containers and path utilities on the host side of the SA path.

Mirrors the parts of the reference that ``Optimizer.optimize`` touches:
  Tensor / TensorNetwork          tnco/app/tn.py:77-362   (structure only: inds, dims, tags; no arrays)
  read_inds                       tnco/utils/tn.py:520-569
  get_hyper_count                 tnco/utils/tn.py:572-595
  get_connected_components        tnco/utils/tn.py:61-106
  merge_contraction_paths         tnco/utils/tn.py:334-401
  fuse                            tnco/utils/tn.py:598-824   (index-level pre-merge of small tensors, load_tn's default)
  contract (structure only)       tnco/utils/tn.py:903-1075, tnco/utils/tensor.py:214-243
Numeric pre-processing (arrays, hyper-index decomposition, circuit loading) is out of scope.
"""
from __future__ import annotations

import json
import math
from collections import Counter, defaultdict
from random import Random
from dataclasses import dataclass
from types import MappingProxyType
from typing import Any, Iterable


class JSONEncoder(json.JSONEncoder):

    def default(self, obj):
        if isinstance(obj, frozenset):
            return tuple(obj)
        if isinstance(obj, Tensor):
            return dict(inds=obj.inds, dims=obj.dims, array=None, tags=obj.tags)
        if isinstance(obj, TensorNetwork):
            return dict(tensors=obj.tensors, output_inds=obj.output_inds, sparse_inds=obj.sparse_inds)
        if hasattr(obj, 'to_json'):
            return obj.to_json()
        return super().default(obj)


@dataclass(frozen=True, repr=False, eq=False)
class Tensor:
    """A tensor of the network: its indices and their dimensions (tnco/app/tn.py:77-175, without arrays)."""
    inds: tuple
    dims: Any = None
    array: Any = None
    tags: dict | None = None

    def __post_init__(self):
        if self.array is not None:
            raise ValueError('tnco_b200 handles network structure only: pass dims, not arrays.')
        if self.dims is None:
            raise ValueError("One of 'dims' or 'array' must be provided.")
        object.__setattr__(self, 'inds', tuple(self.inds))
        try:
            d = int(self.dims)
        except (TypeError, ValueError):
            object.__setattr__(self, 'dims', tuple(self.dims))
        else:
            if d != self.dims or d < 1:
                raise ValueError("'dims' must be a positive integer.")
            object.__setattr__(self, 'dims', (d,) * len(self.inds))
        object.__setattr__(self, 'tags', {} if self.tags is None else dict(self.tags))
        if any(int(d) != d or d < 1 for d in self.dims):
            raise ValueError('Every dimension must be a positive integers.')
        if len(self.dims) != len(self.inds):
            raise ValueError("Wrong number of 'inds'.")

    def __eq__(self, other):
        return isinstance(other, Tensor) and self.inds == other.inds and self.dims == other.dims

    def __hash__(self):
        return hash((self.inds, self.dims))

    def __repr__(self):
        return 'Tensor(ndim={}, array=None{})'.format(self.ndim, ', tags={}'.format(self.tags) if self.tags else '')

    @property
    def ndim(self):
        return len(self.dims)

    def to_json(self):
        return json.dumps(self, cls=JSONEncoder)


def get_hyper_count(ts_inds: Iterable[Iterable], output_inds: Iterable | None = None) -> dict:
    """#tensors an index appears in minus one, plus one if it is an output index (tnco/utils/tn.py:572-595)."""
    hc = {x: n - 1 for x, n in Counter(x for xs in ts_inds for x in xs).items()}
    if output_inds is not None:
        for x in output_inds:
            hc[x] = hc.get(x, 0) + 1
    return hc


@dataclass(frozen=True, repr=False)
class TensorNetwork:
    """tnco/app/tn.py:178-362 (structure only)."""
    tensors: tuple
    output_inds: frozenset | None = None
    sparse_inds: frozenset | None = None
    tags: dict | None = None

    def __post_init__(self):
        object.__setattr__(self, 'tensors', tuple(self.tensors))
        if any(not isinstance(t, Tensor) for t in self.tensors):
            raise ValueError("'tensors' must be a list of valid 'Tensor'.")
        object.__setattr__(self, 'sparse_inds', frozenset(() if self.sparse_inds is None else self.sparse_inds))
        object.__setattr__(self, '_inds', frozenset(x for t in self.tensors for x in t.inds))
        dims = {}
        for t in self.tensors:
            for x, d in zip(t.inds, t.dims):
                if dims.setdefault(x, d) != d:
                    raise ValueError("Dimensions of 'tensors' are not consistent.")
        object.__setattr__(self, '_dims', dims)
        hc = get_hyper_count(self.ts_inds)
        if self.output_inds is None:
            if any(v > 1 for v in hc.values()):
                raise ValueError("'output_inds' must be provided if 'ts_inds' has hyper-indices.")
            object.__setattr__(self, 'output_inds', frozenset(x for x, v in hc.items() if v == 0))
        object.__setattr__(self, 'output_inds', frozenset(self.output_inds))
        if not self.output_inds.issubset(self._inds):
            raise ValueError("'output_inds' contains indices not in 'tensors'.")
        if not self.sparse_inds.issubset(self._inds):
            raise ValueError("'sparse_inds' contains indices not in 'tensors'.")
        object.__setattr__(self, 'tags', dict(() if self.tags is None else self.tags))

    def __repr__(self):
        return 'TensorNetwork(n_tensors={}, n_inds={})'.format(self.n_tensors, self.n_inds)

    n_tensors = property(lambda s: len(s.tensors))
    n_inds = property(lambda s: len(s._inds))
    ts_inds = property(lambda s: tuple(t.inds for t in s.tensors))
    arrays = property(lambda s: tuple(None for _ in s.tensors))
    ts_tags = property(lambda s: tuple(t.tags for t in s.tensors))
    inds = property(lambda s: s._inds)
    dims = property(lambda s: MappingProxyType(s._dims))

    def __len__(self):
        return self.n_tensors

    def __getitem__(self, k):
        return self.tensors[k]

    def __iter__(self):
        return iter(self.tensors)

    def to_json(self):
        return json.dumps(self, cls=JSONEncoder)


def read_inds(inds_map: dict, *, output_index_token='*', sparse_index_token='/'):
    """index -> (dimension, tensor names...)  =>  (tensor_map, dims, output_inds, sparse_inds)."""
    if output_index_token == sparse_index_token:
        raise ValueError("'output_index_token' and 'sparse_index_token' must differ.")
    tensor_map, dims = defaultdict(list), {}
    for i, (d, *ts) in inds_map.items():
        dims[i] = int(d)
        for t in ts:
            tensor_map[t].append(i)
    output_inds = frozenset(tensor_map.pop(output_index_token, ()))
    sparse_inds = frozenset(tensor_map.pop(sparse_index_token, ()))
    return {k: tuple(v) for k, v in tensor_map.items()}, dims, output_inds, sparse_inds


def get_connected_components(ts_inds: Iterable[Iterable]) -> list[tuple[int, ...]]:
    """Union-find over shared indices; components in order of their smallest tensor, sorted inside."""
    ts = list(ts_inds)
    parent = list(range(len(ts)))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    owner = {}
    for t, xs in enumerate(ts):
        for x in xs:
            if x in owner:
                a, b = find(t), find(owner[x])
                if a != b:
                    parent[max(a, b)] = min(a, b)
            else:
                owner[x] = t
    comps = {}
    for t in range(len(ts)):
        comps.setdefault(find(t), []).append(t)
    return [tuple(sorted(v)) for v in comps.values()]


def merge_contraction_paths(n_tensors: int, paths, *, autocomplete: bool = True):
    """Merge per-component linear paths (each expressed over all n_tensors tensors) into one path
    (tnco/utils/tn.py:334-401).  Done by the C++ helper ``tnb_merge_paths`` (Fenwick trees instead of list searches).

    >>> merge_contraction_paths(4, [[(0, 1)], [(2, 3)]])
    [(0, 1), (0, 1), (0, 1)]
    """
    import numpy as np

    from .engine import merge_paths
    paths = [[tuple(q) for q in p] for p in paths]
    lens = [len(p) for p in paths]
    cat = np.array([q for p in paths for q in p], np.int32).reshape(1, sum(lens), 2)
    out = merge_paths(n_tensors, lens, cat)[0]
    return [tuple(x) for x in out[:len(out) if autocomplete else sum(lens)].tolist()]


def _unique(xs):
    """Order-preserving de-duplication (the reference's OrderedFrozenSet / unique_everseen)."""
    return tuple(dict.fromkeys(xs))


def _dims_dict(dims, all_inds):
    try:
        d = int(dims)
    except (TypeError, ValueError):
        return dict(dims)
    return {x: d for x in all_inds}


def fuse(ts_inds, dims, max_width, output_inds=None, *, exclude_inds=(), seed=None, return_fused_inds=False,
         verbose=False):
    """Random pre-merge of index-sharing tensors while the merged tensor stays within ``max_width``
    (tnco/utils/tn.py:598-824).  Same draws from ``Random(seed)`` in the same order over the same containers, so
    the returned path equals the reference's for the same seed (tests/golden/host_fuse.json).

    >>> fuse([['i', 'j'], ['j', 'k'], ['k', 'l']], 2, max_width=2, seed=42)
    [(0, 1), (0, 1)]
    """
    rng = Random(seed)
    ts = dict(enumerate(map(tuple, ts_inds)))
    all_inds = _unique(x for xs in ts.values() for x in xs)
    exclude_inds = frozenset(exclude_inds)
    if not exclude_inds.issubset(all_inds):
        raise ValueError("'exclude_inds' contains indices not in 'ts_inds'.")
    dims = _dims_dict(dims, all_inds)
    if not set(all_inds).issubset(dims):
        raise ValueError("'dims' is missing some indices.")
    hyper_count = get_hyper_count(ts.values())
    if output_inds is None:
        if any(v > 1 for v in hyper_count.values()):
            raise ValueError("'output_inds' must be provided if 'ts_inds' has hyper-indices.")
        output_inds = (x for x, v in hyper_count.items() if v == 0)
    output_inds = frozenset(output_inds)
    if not output_inds.issubset(all_inds):
        raise ValueError("'output_inds' is not consistent with 'ts_inds'.")
    # index -> tensors holding it; the draws below iterate these sets, so they are built and updated with the very
    # set operations of the reference (tn.py:690-697, 789-791) -- a set's iteration order depends on its history
    index2tensors = {}
    for t, xs in ts.items():
        for x in xs:
            index2tensors.setdefault(x, []).append(t)
    index2tensors = {x: set(v) for x, v in index2tensors.items()}
    dangling = frozenset(x for x, v in hyper_count.items() if v == 0)
    avail = [x for x in all_inds if x not in exclude_inds and x not in dangling]
    t_idx = len(ts)
    merged = []
    while avail:
        index = avail.pop(rng.randrange(len(avail)))
        if not hyper_count.get(index):
            continue
        px, py = rng.sample(tuple(index2tensors[index]), k=2)
        tx, ty = ts[px], ts[py]
        sx, sy = frozenset(tx), frozenset(ty)
        if (sx | sy) & exclude_inds:
            continue
        shared = sx & sy
        hyper = frozenset(x for x in shared if hyper_count[x] > 1)
        keep = (sx ^ sy) | hyper | (output_inds & (sx | sy))
        tz = _unique([x for x in tx if x in keep] + [y for y in ty if y in keep])
        if sum(map(math.log2, map(dims.get, tz))) > max_width:
            continue
        for x in shared:
            hyper_count[x] -= 1
        for x in tz:
            index2tensors[x] -= {px, py}
            index2tensors[x] |= {t_idx}
        for x in shared - hyper - output_inds:
            del index2tensors[x]
        del ts[px]
        del ts[py]
        ts[t_idx] = tz
        t_idx += 1
        if hyper_count.get(index):
            avail.append(index)
        merged.append((px, py, tz))
    # SSA ids -> linear (einsum-style) positions
    path, fused_inds, positions = [], [], list(range(t_idx))
    for px, py, tz in merged:
        px, py = sorted((px, py))
        py = positions.index(py)
        del positions[py]
        px = positions.index(px)
        del positions[px]
        path.append((px, py))
        fused_inds.append(tz)
    return (path, fused_inds) if return_fused_inds else path


def contract(path, ts_inds, output_inds=None, *, dims=None):
    """Index bookkeeping of a linear contraction path (tnco/utils/tn.py:903-1075 with ``arrays=None``): returns
    ``(ts_inds, output_inds)`` after the path.  Result indices are ordered as ``tensordot`` orders them
    (tnco/utils/tensor.py:229-243): hyper-indices first, then x's own, then y's own.

    >>> ts, out = contract([(0, 1)], [['i', 'j'], ['j', 'k']], dims=2)
    >>> ts, sorted(out)
    ([('i', 'k')], ['i', 'k'])
    """
    if dims is None:
        raise ValueError("Either 'dims' or 'arrays' must be provided.")
    ts = list(map(tuple, ts_inds))
    try:
        int(dims)
    except (TypeError, ValueError):
        if not frozenset(dims).issuperset(x for xs in ts for x in xs):
            raise ValueError("'ts_inds' has indices not in 'dims'.")
    hyper_count = get_hyper_count(ts)
    if output_inds is None:
        if any(v > 1 for v in hyper_count.values()):
            raise ValueError("'output_inds' must be provided if 'ts_inds' has hyper-indices.")
        output_inds = (x for x, v in hyper_count.items() if v == 0)
    output_inds = frozenset(output_inds)
    if not output_inds.issubset(x for xs in ts for x in xs):
        raise ValueError("'output_inds' is not consistent with 'ts_inds'.")
    for x, y in map(sorted, path):
        if x == y:
            raise ValueError("'path' is not valid.")
        ys = ts.pop(y)
        xs = ts.pop(x)
        shared = frozenset(xs) & frozenset(ys)
        hyper = frozenset(i for i in shared if hyper_count[i] > 1) | (output_inds & shared)
        for i in shared:
            hyper_count[i] -= 1
        ts.append(_unique([*hyper, *(i for i in xs if i not in shared), *(i for i in ys if i not in shared)]))
    return ts, output_inds.intersection(x for xs in ts for x in xs)
