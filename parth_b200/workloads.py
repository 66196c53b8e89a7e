"""Synthetic update streams for bench.py and the full-size GPU tests (host side, numpy only).

Config 2 of BASELINE.json (variant 2a "tree-aware"): 7-point Poisson grid
n^3, dim = 1; update k inserts `edges_per_update` symmetric edges between random DOFs of the two
cousin leaves under one level-(L-1) node S_k of the separator-tree fixture (the construction of
reference tests/example_creator.cpp:61-84), so ~1 % of the DOFs (one level-7 sub-tree at 8 levels)
become dirty per update.  Consecutive updates use different nodes S_k, so every step diffs a real
change against the previous graph.
"""
import numpy as np

from . import synth


def insert_entries(Ap, Ai, cols, rows):
    """insert (col,row) entries (not present yet) into a CSC pattern with ascending columns"""
    cols = np.asarray(cols, np.int64)
    rows = np.asarray(rows, np.int64)
    order = np.lexsort((rows, cols))
    cols, rows = cols[order], rows[order]
    pos = np.empty(len(cols), np.int64)
    for i, (c, r) in enumerate(zip(cols.tolist(), rows.tolist())):
        a, b = int(Ap[c]), int(Ap[c + 1])
        pos[i] = a + np.searchsorted(Ai[a:b], r)
    Ai2 = np.insert(Ai, pos, rows.astype(np.int32))
    add = np.zeros(len(Ap), np.int64)
    np.add.at(add, cols + 1, 1)
    Ap2 = (Ap.astype(np.int64) + np.cumsum(add)).astype(np.int32)
    return Ap2, Ai2


class PoissonUpdateStream:
    """base matrix + tree fixture + a generator of changed matrices"""

    def __init__(self, n=200, levels=8, edges_per_update=1000, seed=1):
        self.n, self.levels, self.edges = n, levels, edges_per_update
        self.N, self.Ap, self.Ai = synth.poisson3d(n)
        self.Mp, self.Mi = synth.grid_graph(n)
        meta, node_ptr, dofs, labels = synth.geometric_tree(n, levels)
        self.meta, self.node_ptr, self.dofs, self.labels = meta, node_ptr, dofs, labels
        self.nodes = len(meta)
        self.rng = np.random.default_rng(seed)
        # level-(levels-1) nodes whose two children are non-empty leaves
        first = 2 ** (levels - 1) - 1
        self.parents = [s for s in range(first, 2 * first + 1)
                        if meta[s, 0] != -1 and meta[s, 1] != -1
                        and node_ptr[meta[s, 0] + 1] > node_ptr[meta[s, 0]]
                        and node_ptr[meta[s, 1] + 1] > node_ptr[meta[s, 1]]]

    def tree(self):
        from types import SimpleNamespace
        return SimpleNamespace(M_n=self.N, levels=self.levels, nodes=self.nodes, meta=self.meta,
                               node_ptr=self.node_ptr, dofs=self.dofs, labels=self.labels)

    def node_dofs(self, i):
        return self.dofs[self.node_ptr[i]:self.node_ptr[i + 1]]

    def dirty_fraction(self, k):
        s = self.parents[k % len(self.parents)]
        sub = synth.subtree_sizes(self.meta, self.node_ptr)
        return float(sub[s]) / self.N

    def edges_of(self, k):
        s = self.parents[k % len(self.parents)]
        a = self.node_dofs(self.meta[s, 0])
        b = self.node_dofs(self.meta[s, 1])
        rng = np.random.default_rng(1000 + k)
        pa = rng.choice(a, self.edges)
        pb = rng.choice(b, self.edges)
        e = np.unique(np.stack([pa, pb], axis=1), axis=0)
        return s, e

    def matrix(self, k):
        """CSC pattern of update k: base + symmetric cousin-leaf edges"""
        s, e = self.edges_of(k)
        cols = np.concatenate([e[:, 0], e[:, 1]])
        rows = np.concatenate([e[:, 1], e[:, 0]])
        Ap, Ai = insert_entries(self.Ap, self.Ai, cols, rows)
        return Ap, Ai

    def mesh(self, k):
        s, e = self.edges_of(k)
        cols = np.concatenate([e[:, 0], e[:, 1]])
        rows = np.concatenate([e[:, 1], e[:, 0]])
        return insert_entries(self.Mp, self.Mi, cols, rows)


def build_bytes(N, nnzA, M, nnzM, dim=1):
    """algorithmic bytes of stage (a), SURVEY.md section 8d"""
    return 4 * ((N + 1) + nnzA // dim + (M + 1) + nnzM)


def detect_bytes(M, nnzM):
    return 2 * 4 * ((M + 1) + nnzM)
