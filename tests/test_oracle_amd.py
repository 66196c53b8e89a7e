"""CPU: the AMD restatement (oracle/amd_restated.c).  Parity with SuiteSparse v7.11.0 binaries is
UNPINNED (no SuiteSparse in the container, no golden vector in the reference); these tests pin
what can be pinned: valid permutations, documented special cases, and ordering quality against an
exact minimum-degree elimination on small graphs."""
import numpy as np
import pytest

from parth_b200 import synth


def fill_of(Mp, Mi, perm, port):
    parent = port.etree(Mp, Mi, perm)
    return port.colcount(Mp, Mi, perm, parent)[0]


def exact_min_degree_fill(Mp, Mi):
    n = len(Mp) - 1
    adj = [set(Mi[Mp[i]:Mp[i + 1]].tolist()) for i in range(n)]
    alive = set(range(n))
    lnz = 0
    while alive:
        v = min(alive, key=lambda x: (len(adj[x]), x))
        nb = adj[v]
        lnz += len(nb) + 1
        for a in nb:
            adj[a] |= nb - {a}
            adj[a].discard(v)
        alive.discard(v)
    return lnz


@pytest.mark.parametrize("n", [2, 3, 5, 8, 12])
def test_grid_permutation_valid_and_good(port, n):
    Mp, Mi = synth.grid_graph(n)
    P = port.amd(Mp, Mi)
    assert sorted(P.tolist()) == list(range(n ** 3))
    if n <= 8:
        approx, exact = fill_of(Mp, Mi, P, port), exact_min_degree_fill(Mp, Mi)
        natural = fill_of(Mp, Mi, np.arange(n ** 3), port)
        assert approx <= 1.35 * exact + 8
        assert approx <= natural


def test_empty_graph_is_identity(port):
    Mp = np.zeros(11, np.int32)
    assert port.amd(Mp, np.zeros(0, np.int32)).tolist() == list(range(10))


def test_path_and_star(port):
    # path graph: a perfect elimination order exists -> no fill beyond the pattern
    n = 50
    r = np.arange(n - 1)
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r, r + 1]), np.concatenate([r + 1, r]))
    P = port.amd(Mp, Mi)
    assert fill_of(Mp, Mi, P, port) == 2 * n - 1
    # star: the hub is "dense" for n > 16+... either way the leaves must come first
    n = 400
    hub = np.arange(1, n)
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([hub, np.zeros(n - 1, np.int64)]),
                                 np.concatenate([np.zeros(n - 1, np.int64), hub]))
    P = port.amd(Mp, Mi)
    assert sorted(P.tolist()) == list(range(n)) and P[-1] == 0
    assert fill_of(Mp, Mi, P, port) == 2 * n - 1


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs(port, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(5, 300))
    m = int(rng.integers(n, 5 * n))
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = r != c
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
    P = port.amd(Mp, Mi)
    assert sorted(P.tolist()) == list(range(n))
    if n <= 120:
        assert fill_of(Mp, Mi, P, port) <= 1.5 * exact_min_degree_fill(Mp, Mi) + 10


def test_jumbled_input_gives_same_order_as_its_sorted_transpose(port):
    # unsorted columns go through the sorted duplicate-free transpose first (amd_preprocess)
    rng = np.random.default_rng(5)
    Mp, Mi = synth.grid_graph(6)
    Mi2 = Mi.copy()
    for c in range(len(Mp) - 1):
        rng.shuffle(Mi2[Mp[c]:Mp[c + 1]])
    assert np.array_equal(port.amd(Mp, Mi2), port.amd(Mp, Mi))
