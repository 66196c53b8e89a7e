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
