"""GPU: the staged copies between the caller's pageable arrays and the device (parth_b200/csrc/hostcopy.cuh).

A transfer moves through a ring of page-locked chunks; the DMA of the last chunk of one array may still be reading its
slot when the upload of the next array (Mp, then Mi; Ap, then Ai) starts to fill the ring.  The sizes below put a
four-byte tail chunk (Mp[n] / Ap[N], the entry count) right before a second multi-chunk upload: at 256^3 that entry was
overwritten by the first bytes of the next array.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _path_graph(n):
    deg = np.full(n, 2, np.int32); deg[0] = deg[-1] = 1
    Mp = np.zeros(n + 1, np.int32); np.cumsum(deg, out=Mp[1:])
    Mi = np.empty(int(Mp[-1]), np.int32)
    left = np.arange(1, n, dtype=np.int64)
    Mi[Mp[left]] = left - 1
    right = np.arange(0, n - 1, dtype=np.int64)
    Mi[Mp[right + 1] - 1] = right + 1
    return Mp, Mi


@pytest.mark.parametrize("n", [(1 << 21), (1 << 21) + 12345, 3 * (1 << 21)])
def test_pageable_mesh_upload_round_trips_bit_exact(gpu_api, n):
    """Mp is k * 8 MB + 4 bytes (or a ragged tail): every entry, the last one included, must arrive"""
    Mp, Mi = _path_graph(n)
    g = gpu_api.ParthAPI(); g.verbose = False
    for rep in range(3):
        g.setMesh(n, Mp, Mi)
        Mp2, Mi2 = g.mesh()
        assert Mp2[-1] == Mp[-1]
        assert np.array_equal(Mp2, Mp) and np.array_equal(Mi2, Mi)


def test_pageable_matrix_upload_keeps_every_entry(gpu_api):
    """setMatrix(N, Ap, Ai) from plain numpy arrays (pageable): the mesh graph built from them is the pattern minus the
    diagonal, down to the last column"""
    n = 1 << 21
    Mp, Mi = _path_graph(n)
    # pattern with the diagonal: row v = sorted(neighbours + [v])
    Ap = Mp + np.arange(n + 1, dtype=np.int32)
    Ai = np.empty(int(Ap[-1]), np.int32)
    v = np.arange(n, dtype=np.int64)
    has_left = v > 0
    Ai[Ap[:-1][has_left]] = v[has_left] - 1
    Ai[Ap[:-1] + has_left] = v
    has_right = v < n - 1
    Ai[(Ap[:-1] + has_left + 1)[has_right]] = v[has_right] + 1
    g = gpu_api.ParthAPI(); g.verbose = False
    for rep in range(3):
        g.setMatrix(n, Ap, Ai)
        Mp2, Mi2 = g.mesh()
        assert np.array_equal(Mp2, Mp) and np.array_equal(Mi2, Mi)
