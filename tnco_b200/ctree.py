"""Contraction tree container in the reference's tree and result format (tnco/ctree.py).

A tree is the reference's node list: leaves first, root last, ``Node(children=(c0, c1), parent=p)`` with
-1 for "none" (include/tnco/node.hpp:33-43, tree.hpp:73-99), one index set per node.  Construction from a
linear (einsum) path follows tnco/ctree.py:108-226 including the hyper-count rule for intermediate index
sets; ``path()`` follows tnco/ctree.py:350-388 but runs in C++ (Fenwick tree instead of list.index).
"""
from __future__ import annotations

import math
from collections import namedtuple
from types import MappingProxyType
from typing import Iterable

import numpy as np

from .tn import get_hyper_count

__all__ = ['ContractionTree', 'Node']

Node = namedtuple('Node', ['children', 'parent'])


def _unique(seq):
    return tuple(dict.fromkeys(seq))


class ContractionTree:
    """
    >>> from tnco_b200.ctree import ContractionTree
    >>> ctree = ContractionTree([(0, 1)], [['i', 'j'], ['j', 'k']], {'i': 2, 'j': 2, 'k': 2})
    >>> ctree.max_width()
    2.0
    """

    def __init__(self, path, ts_inds, dims, *, output_inds=None, check_shared_inds=False, **kwargs):
        _arrays = kwargs.pop('_arrays', None)
        if kwargs:
            raise TypeError('Got unexpected keyword arguments.')
        ts_inds = [tuple(x) for x in ts_inds]
        n_tensors = len(ts_inds)
        if _arrays is not None:
            # (parent, child0, child1, tensors_pos): a tree over the tensors `tensors_pos` of ts_inds
            parent, c0, c1, tensors_pos = _arrays
            self._tensors_pos = tuple(int(x) for x in tensors_pos)
            contraction = None
        else:
            contraction, pos = [], list(range(n_tensors))
            for i, xs in enumerate(path):
                x, y = sorted(xs)
                py = pos.pop(y)
                px = pos.pop(x)
                pos.append(i + n_tensors)
                contraction.append((px, py, pos[-1]))
            self._tensors_pos = tuple(sorted(x for x in _unique(v for c in contraction for v in c) if x < n_tensors))
            if not contraction:
                raise ValueError("'path' is empty.")
        self._n_tensors = n_tensors
        leaves = [ts_inds[x] for x in self._tensors_pos]
        n = len(leaves)
        all_inds = _unique(x for xs in leaves for x in xs)
        hyper_count = get_hyper_count(leaves)
        if output_inds is None:
            if any(v > 1 for v in hyper_count.values()):
                raise ValueError("'output_inds' must be provided if 'ts_inds' has hyper-indices.")
            output_inds = frozenset(x for x, v in hyper_count.items() if v == 0)
        output_inds = frozenset(output_inds).intersection(all_inds)
        for x in output_inds:
            hyper_count[x] += 1
        if contraction is not None:
            used = sorted(_unique(v for c in contraction for v in c))
            tmap = {p: k for k, p in enumerate(used)}
            tree = [tuple(tmap[v] for v in c) for c in contraction]
            N = max(v for c in tree for v in c) + 1
            if N != 2 * n - 1:
                raise ValueError("'path' does not contract its tensors to a single one.")
            parent = np.full(N, -1, np.int32)
            c0 = np.full(N, -1, np.int32)
            c1 = np.full(N, -1, np.int32)
            for x, y, z in tree:
                parent[x] = parent[y] = z
                c0[z], c1[z] = x, y
        else:
            parent, c0, c1 = (np.asarray(v, np.int32) for v in (parent, c0, c1))
            N = len(parent)
        # index sets of every node, in an order where children precede parents
        sets = [frozenset(xs) for xs in leaves] + [None] * (N - n)
        order = [z for z in range(n, N)] if contraction is not None else self._post_order(c0, c1, n)
        for z in order:
            ix, iy = sets[c0[z]], sets[c1[z]]
            shared = ix & iy
            if check_shared_inds and not shared:
                raise ValueError("'check_shared_inds' failed.")
            iz = set(ix ^ iy)
            for s in shared:
                hyper_count[s] -= 1
                if hyper_count[s] > 0:
                    iz.add(s)
            sets[z] = frozenset(iz)
        self._parent, self._c0, self._c1 = parent, c0, c1
        self._inds = tuple(sets)
        self._leaves = tuple(tuple(v) for v in leaves)   # ordered leaf index tuples: they fix the index bit positions
        self._inds_order = _unique(x for xs in ([tuple(v) for v in leaves] + [tuple(s) for s in sets[n:]]) for x in xs)
        try:
            self._dims = {x: int(dims[x]) for x in self._inds_order}
        except TypeError:
            if int(dims) != dims:
                raise ValueError("'dims' is not valid.")
            self._dims = {x: int(dims) for x in self._inds_order}

    @staticmethod
    def _post_order(c0, c1, n):
        out, stack, seen = [], [len(c0) - 1], set()
        while stack:
            z = stack[-1]
            if z < n or z in seen:
                stack.pop()
                if z >= n:
                    out.append(z)
            else:
                seen.add(z)
                stack.append(int(c1[z]))
                stack.append(int(c0[z]))
        return out

    @classmethod
    def from_arrays(cls, parent, child0, child1, ts_inds, dims, *, tensors_pos=None, output_inds=None):
        """Tree given as the engine returns it (reference node numbering) over tensors ``tensors_pos``."""
        n = (len(parent) + 1) // 2
        tp = tuple(range(n)) if tensors_pos is None else tuple(tensors_pos)
        return cls(None, ts_inds, dims, output_inds=output_inds, _arrays=(parent, child0, child1, tp))

    # ---- reference surface
    def __len__(self):
        return len(self._parent)

    def __repr__(self):
        return f'ContractionTree(n_nodes={len(self)}, n_inds={self.n_inds})'

    def __eq__(self, other):
        return (isinstance(other, ContractionTree) and (self._parent == other._parent).all() and
                (self._c0 == other._c0).all() and (self._c1 == other._c1).all() and self._inds == other._inds and
                self._inds_order == other._inds_order)

    @property
    def n_leaves(self):
        return (len(self._parent) + 1) // 2

    @property
    def n_inds(self):
        return len(self._inds_order)

    @property
    def nodes(self):
        return [Node((int(a), int(b)), int(p)) for a, b, p in zip(self._c0, self._c1, self._parent)]

    @property
    def inds(self):
        return self._inds

    @property
    def dims(self):
        return MappingProxyType(dict(self._dims))

    def all_inds(self):
        return frozenset(self._inds_order)

    def output_inds(self):
        return self._inds[-1]

    def arrays(self):
        """(parent, child0, child1) int32 arrays in reference node numbering."""
        return self._parent, self._c0, self._c1

    def leaf_bits(self):
        """([n_leaves][W32] uint32 bitsets, n_inds) over this tree's own index order."""
        imap = {x: k for k, x in enumerate(self._inds_order)}
        W = (len(imap) + 31) // 32
        out = np.zeros((self.n_leaves, W), np.uint32)
        for t in range(self.n_leaves):
            for x in self._inds[t]:
                k = imap[x]
                out[t, k >> 5] |= np.uint32(1 << (k & 31))
        return out, len(imap)

    def path(self):
        """Contraction path in linear (einsum) format over all tensors of the original network."""
        from .engine import tree_to_path
        p = tree_to_path(self._c0, self._c1, n_tensors=self._n_tensors,
                         tensors_pos=np.asarray(self._tensors_pos, np.int32))
        return [(int(x), int(y)) for x, y in p]

    def max_width(self):
        return max(math.log2(math.prod(self._dims[x] for x in xs)) for xs in self._inds)
