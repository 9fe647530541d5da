"""Synthetic tensor networks for the benchmark configurations of BASELINE.json (structure only, no data).

All circuit networks use the tensorisation of SURVEY.md section 8: one tensor per two-qubit gate (bond
dimension 2), one-qubit gates and the |0> / <x| boundary vectors absorbed, amplitude (no open indices),
hence one index per wire segment between consecutive two-qubit gates of a qubit and no hyper-indices.
Every generator returns ``(ts_inds, n_inds)`` with ``ts_inds[t]`` the list of index ids of tensor ``t``.
"""
from __future__ import annotations


def regular_graph(n, seed=0, degree=3):
    """C1 / C5: random `degree`-regular graph, one tensor per vertex, one index per edge (d=2)."""
    import networkx as nx
    g = nx.random_regular_graph(degree, n, seed=seed)
    while not nx.is_connected(g):
        seed += 1000003
        g = nx.random_regular_graph(degree, n, seed=seed)
    ts = [[] for _ in range(n)]
    for k, (a, b) in enumerate(sorted(tuple(sorted(e)) for e in g.edges())):
        ts[a].append(k)
        ts[b].append(k)
    return ts, g.number_of_edges()


def _circuit_network(n_qubits, layers):
    """layers: list of lists of (qa, qb) two-qubit gates.  Wire segments between consecutive gates on a
    qubit become indices; first / last legs are absorbed boundary vectors."""
    ts, last, n_inds = [], [None] * n_qubits, 0
    for gates in layers:
        for qa, qb in gates:
            t = len(ts)
            ts.append([])
            for q in (qa, qb):
                if last[q] is not None:
                    ts[last[q]].append(n_inds)
                    ts[t].append(n_inds)
                    n_inds += 1
                last[q] = t
    # drop tensors that ended up without any index (isolated gates) -- cannot happen for the configs used
    assert all(ts), 'isolated gate'
    return ts, n_inds


_SEQ = 'ABCDCDAB'


def grid_rqc(rows=6, cols=6, depth=12):
    """C2: rows x cols grid random circuit, `depth` cycles, coupler activation pattern ABCDCDAB..."""
    q = lambda r, c: r * cols + c  # noqa: E731
    pats = {
        'A': [(q(r, c), q(r, c + 1)) for r in range(rows) for c in range(0, cols - 1, 2)],
        'B': [(q(r, c), q(r, c + 1)) for r in range(rows) for c in range(1, cols - 1, 2)],
        'C': [(q(r, c), q(r + 1, c)) for c in range(cols) for r in range(0, rows - 1, 2)],
        'D': [(q(r, c), q(r + 1, c)) for c in range(cols) for r in range(1, rows - 1, 2)],
    }
    return _circuit_network(rows * cols, [pats[_SEQ[m % 8]] for m in range(depth)])


def sycamore(m=20):
    """C3 / C4: Sycamore-style 53-qubit circuit with m cycles.  54-site diagonal lattice (9 rows of 6,
    rows offset alternately), one edge qubit removed -> 53 qubits, 86 couplers; the four coupler classes
    are the two diagonal directions split by row parity (43 + 43 couplers), sequence ABCDCDAB."""
    removed = (0, 1)
    ids = {}
    for j in range(9):
        for i in range(6):
            if (j, i) != removed:
                ids[(j, i)] = len(ids)
    pats = {'A': [], 'B': [], 'C': [], 'D': []}
    for j in range(8):
        for i in range(6):
            if j % 2 == 0:  # x = 2i -> right: (j+1, i), left: (j+1, i-1)
                right, left = (j + 1, i), (j + 1, i - 1)
            else:           # x = 2i+1 -> right: (j+1, i+1), left: (j+1, i)
                right, left = (j + 1, i + 1), (j + 1, i)
            for nb, name in ((right, 'A' if j % 2 == 0 else 'B'), (left, 'C' if j % 2 == 0 else 'D')):
                if (j, i) in ids and nb in ids:
                    pats[name].append((ids[(j, i)], ids[nb]))
    assert len(ids) == 53 and sum(map(len, pats.values())) == 86
    return _circuit_network(53, [pats[_SEQ[k % 8]] for k in range(m)])


CONFIGS = {
    'C1': dict(desc='random 3-regular graph TN, 64 tensors, bond dim 2', make=lambda: regular_graph(64, 0)),
    'C2': dict(desc='2D-grid 6x6 random circuit depth 12 TN', make=lambda: grid_rqc(6, 6, 12)),
    'C3': dict(desc='Sycamore-style 53-qubit m=14 random circuit TN', make=lambda: sycamore(14)),
    'C4': dict(desc='Sycamore-style 53-qubit m=20 TN, max width 32', make=lambda: sycamore(20)),
    'C5': dict(desc='random 3-regular graph TN, 1000 tensors', make=lambda: regular_graph(1000, 0)),
}
