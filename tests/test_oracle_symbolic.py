"""CPU: etree / column counts / nnz(L) restatement (oracle/symbolic_oracle.c) against the known
answers of SURVEY.md Appendix B.3 and against brute-force symbolic elimination."""
import numpy as np
import pytest

from parth_b200 import synth
from tests.golden import fixtures


def brute_colcounts(Mp, Mi, perm):
    n = len(Mp) - 1
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    adj = [set() for _ in range(n)]
    for old in range(n):
        k = inv[old]
        for nb in Mi[Mp[old]:Mp[old + 1]]:
            j = inv[nb]
            if j > k:
                adj[k].add(int(j))
    parent = -np.ones(n, np.int64)
    cc = np.zeros(n, np.int64)
    for k in range(n):
        cc[k] = len(adj[k]) + 1
        if adj[k]:
            p = min(adj[k])
            parent[k] = p
            adj[p] |= adj[k] - {p}
    return parent, cc


KNOWN = [(4, "id", 883), (4, "aff", 793), (10, "id", 91909), (10, "aff", 105344), (20, "id", 3055619),
         (20, "aff", 3782404)]


@pytest.mark.parametrize("n,kind,lnz", KNOWN)
def test_known_nnzL(port, n, kind, lnz):
    Mp, Mi = synth.grid_graph(n)
    N = n ** 3
    perm = np.arange(N) if kind == "id" else (7919 * np.arange(N) + 13) % N
    parent = port.etree(Mp, Mi, perm)
    got, cc = port.colcount(Mp, Mi, perm, parent)
    assert got == lnz and cc.sum() == lnz


def test_known_nnzL_fish(port):
    N, Ap, Ai = fixtures.matrix("original")
    Mp, Mi = synth.mesh_from_pattern(Ap, Ai, 1)
    perm = np.arange(N)
    parent = port.etree(Mp, Mi, perm)
    lnz, _ = port.colcount(Mp, Mi, perm, parent)
    assert lnz == 1332839 and 2 * lnz - N == 2654892


@pytest.mark.parametrize("seed", range(8))
def test_against_brute_force(port, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 60))
    m = int(rng.integers(0, 4 * n))
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = r != c
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
    perm = rng.permutation(n)
    parent = port.etree(Mp, Mi, perm)
    lnz, cc = port.colcount(Mp, Mi, perm, parent)
    bp, bc = brute_colcounts(Mp, Mi, perm)
    assert np.array_equal(parent, bp) and np.array_equal(cc, bc) and lnz == bc.sum()


def brute_structs(Mp, Mi, perm):
    """row structure of every column of L (strictly below the diagonal) by symbolic elimination"""
    n = len(Mp) - 1
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    adj = [set() for _ in range(n)]
    for old in range(n):
        k = inv[old]
        for nb in Mi[Mp[old]:Mp[old + 1]]:
            if inv[nb] > k:
                adj[k].add(int(inv[nb]))
    for k in range(n):
        if adj[k]:
            p = min(adj[k])
            adj[p] |= adj[k] - {p}
    return adj


def check_supernodes(sn, parent, cc, structs):
    """the partition is the fundamental one: inside a supernode the column structures nest exactly, every boundary
    is forced by the definition, the supernodal etree and the sizes follow from it"""
    n = len(parent)
    sup, smap, spar = sn["super"], sn["supermap"], sn["sparent"]
    assert sup[0] == 0 and sup[-1] == n and len(sup) == sn["nsuper"] + 1 and np.all(np.diff(sup) > 0)
    nchild = np.bincount(parent[parent >= 0], minlength=n)
    for s in range(sn["nsuper"]):
        a, b = int(sup[s]), int(sup[s + 1])
        assert np.all(smap[a:b] == s)
        for j in range(a + 1, b):  # continued columns: struct(j-1) minus {j} equals struct(j)
            assert structs[j - 1] - {j} == structs[j] and parent[j - 1] == j and nchild[j] == 1
        if a > 0:  # the boundary is real: one of the three conditions fails
            assert parent[a - 1] != a or cc[a - 1] != cc[a] + 1 or nchild[a] > 1
        p = parent[b - 1]
        assert spar[s] == (-1 if p < 0 else smap[p])
    assert sn["ssize"] == int(sum(cc[sup[:-1]])) and sn["xsize"] == int(sum(cc[sup[:-1]] * np.diff(sup)))


@pytest.mark.parametrize("seed", range(10))
def test_supernodes_against_brute_force(port, seed):
    rng = np.random.default_rng(100 + seed)
    if seed < 4:  # grids under a nested-dissection-like ordering have real multi-column supernodes
        n = [3, 4, 5, 6][seed]
        Mp, Mi = synth.grid_graph(n)
        N = n ** 3
        meta, node_ptr, dofs, labels = synth.geometric_tree(n, 3)
        perm = synth.assemble_mesh_perm(meta, node_ptr, dofs, labels)
    else:
        N = int(rng.integers(2, 70))
        m = int(rng.integers(0, 5 * N))
        r, c = rng.integers(0, N, m), rng.integers(0, N, m)
        keep = r != c
        _, Mp, Mi = synth.coo_to_csc(N, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
        perm = rng.permutation(N)
    parent = port.etree(Mp, Mi, perm)
    lnz, cc = port.colcount(Mp, Mi, perm, parent)
    sn = port.supernodes(parent, cc)
    check_supernodes(sn, parent.astype(np.int64), cc.astype(np.int64), brute_structs(Mp, Mi, perm))
    if seed < 4 and seed > 0:
        assert sn["nsuper"] < N  # the dense trailing separator is one supernode


def test_supernodes_edge_cases(port):
    sn = port.supernodes(np.array([-1], np.int32), np.array([1], np.int32))
    assert sn["nsuper"] == 1 and sn["super"].tolist() == [0, 1] and sn["sparent"].tolist() == [-1]
    # a path 0-1-2-3 under the identity: only the last two columns nest (counts 2, 2, 2, 1)
    parent, cc = np.array([1, 2, 3, -1], np.int32), np.array([2, 2, 2, 1], np.int32)
    sn = port.supernodes(parent, cc)
    assert sn["super"].tolist() == [0, 1, 2, 4] and sn["sparent"].tolist() == [1, 2, -1]
    # a clique: one supernode
    parent, cc = np.array([1, 2, 3, -1], np.int32), np.array([4, 3, 2, 1], np.int32)
    sn = port.supernodes(parent, cc)
    assert sn["super"].tolist() == [0, 4] and sn["sparent"].tolist() == [-1] and sn["xsize"] == 16
