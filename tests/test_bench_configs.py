"""CPU: the torch generators bench_configs.py uses for the full-size configs (3, 5) against the numpy
generators of parth_b200/synth.py and tests/scenario_defs.py the parity tests are built on."""
import numpy as np
import pytest
import torch

import bench_configs as bc
from parth_b200 import synth, workloads
from tests import scenario_defs as sd

CPU = torch.device("cpu")


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("n", [5, 8])
@pytest.mark.parametrize("dirs,diag", [(bc.GRID_DIRS, True), (bc.KUHN_DIRS, False), (bc.KUHN_DIRS, True)])
def test_lattice_pattern(n, dirs, diag):
    N, Ap, Ai = synth.lattice_pattern(n, dirs, diagonal=diag)
    tAp, tAi = bc.lattice_pattern_dev(n, dirs, diag, CPU, chunk=100)
    assert np.array_equal(Ap, _np(tAp)) and np.array_equal(Ai, _np(tAi))
    assert tAp.dtype == torch.int32 and tAi.dtype == torch.int32


def test_remesh_matrix_kron_chain():
    n, dim = 9, 3
    _, Apv, Aiv = synth.lattice_pattern(n, synth.KUHN_DIRS, diagonal=False)
    P, I, m = sd.remesh_box(n, Apv, Aiv, 3, 6)
    tP, tI, tm = bc.remesh_box_dev(n, torch.from_numpy(Apv), torch.from_numpy(Aiv), 3, 6)
    assert np.array_equal(P, _np(tP)) and np.array_equal(I, _np(tI)) and np.array_equal(m, _np(tm))
    Nm, Apm, Aim = synth.graph_to_matrix(P, I)
    tApm, tAim = bc.graph_to_matrix_dev(tP, tI)
    assert np.array_equal(Apm, _np(tApm)) and np.array_equal(Aim, _np(tAim))
    N3, Ap3, Ai3 = synth.kron_blocks(Apm, Aim, dim)
    tN3, tAp3, tAi3 = bc.kron_blocks_dev(tApm, tAim, dim)
    assert N3 == tN3 and np.array_equal(Ap3, _np(tAp3)) and np.array_equal(Ai3, _np(tAi3))


def test_insert_entries():
    n = 7
    N, Ap, Ai = synth.poisson3d(n)
    rng = np.random.default_rng(3)
    e = rng.integers(0, N, size=(40, 2))
    col = np.repeat(np.arange(N), np.diff(Ap))
    have = set(zip(col.tolist(), Ai.tolist()))
    e = np.array([(a, b) for a, b in e.tolist() if a != b and (a, b) not in have])
    e = np.unique(e, axis=0)
    cols, rows = np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(cols * N + rows)  # the device twin de-duplicates; give numpy the same set
    Ap2, Ai2 = workloads.insert_entries(Ap, Ai, key // N, key % N)
    tAp2, tAi2 = bc.insert_entries_dev(torch.from_numpy(Ap), torch.from_numpy(Ai), torch.from_numpy(cols), torch.from_numpy(rows))
    assert np.array_equal(Ap2, _np(tAp2)) and np.array_equal(Ai2, _np(tAi2))


def test_is_permutation():
    assert bc.is_permutation_dev(torch.tensor([2, 0, 1], dtype=torch.int32), 3)
    assert not bc.is_permutation_dev(torch.tensor([2, 0, 2], dtype=torch.int32), 3)
    assert not bc.is_permutation_dev(torch.tensor([3, 0, 1], dtype=torch.int32), 3)
