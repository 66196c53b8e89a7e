"""CPU: live comparison of the restatement with the compiled reference (oracle/_ref), where the
latter exists (the build container; on the GPU box the prebuilt .so travels with the snapshot)."""
import numpy as np
import pytest

from parth_b200 import synth
from tests import scenario_defs as sd


@pytest.fixture(scope="module")
def both(port, ref):
    if ref is None:
        pytest.skip("oracle/_ref not available")
    return port


def test_kind_strings(both):
    assert both.load("port").cpu_parth_kind() == b"port"
    assert both.load("reference").cpu_parth_kind() == b"reference"


@pytest.mark.parametrize("seed", range(6))
def test_build_random(both, seed):
    rng = np.random.default_rng(seed)
    N = int(rng.integers(5, 400))
    dim = int(rng.choice([1, 1, 2, 3]))
    M = N
    nnz = int(rng.integers(1, 6 * M))
    r, c = rng.integers(0, M, nnz), rng.integers(0, M, nnz)
    if seed % 2:
        r, c = np.concatenate([r, c]), np.concatenate([c, r])
    _, Ap, Ai = synth.coo_to_csc(M, r, c)
    if dim > 1:
        N, Ap, Ai = synth.kron_blocks(Ap, Ai, dim)
    a, b = both.CpuParth("reference"), both.CpuParth("port")
    a.setMatrix(N, Ap, Ai, dim)
    b.setMatrix(N, Ap, Ai, dim)
    for x, y in zip(a.mesh(), b.mesh()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("name", sorted(sd.scenario_list()))
def test_scenarios_live(both, name):
    sc = sd.scenario_list()[name]
    r = sd.run_scenario(both, "reference", sc)
    p = sd.run_scenario(both, "port", sc)
    for k in ("num_changes", "reuse", "edges_hash"):
        assert str(r[k]) == str(p[k]), k
    assert r["dirty_fine"].tolist() == p["dirty_fine"].tolist()
    assert r["dirty_coarse"].tolist() == p["dirty_coarse"].tolist()
    if int(r["deterministic"]):
        for k in r:
            if k.endswith("_hash"):
                assert str(r[k]) == str(p[k]), k


def test_submesh_extraction(both):
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    a, b = both.CpuParth("reference"), both.CpuParth("port")
    for h in (a, b):
        h.set_options(reorder_type=1, nd_levels=5)
        h.import_tree(tree, Mp, Mi)
        h.setMesh(len(Mp) - 1, Mp, Mi)
    for node, coarse in [(40, False), (0, False), (3, True), (12, True), (62, False)]:
        for x, y in zip(a.submesh(node, coarse), b.submesh(node, coarse)):
            assert np.array_equal(x, y)


def test_permute_fine_direct(both):
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    res = []
    for kind in ("reference", "port"):
        h = both.CpuParth(kind)
        h.set_options(reorder_type=1, nd_levels=5)
        h.import_tree(tree, Mp, Mi)
        h.setMesh(len(Mp) - 1, Mp, Mi)
        h.permute_fine([33, 40, 0, 5, 62])
        t = h.export_tree()
        res.append((t.labels.copy(), t.mesh_perm.copy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert sorted(res[0][1].tolist()) == list(range(16 ** 3))


@pytest.mark.parametrize("n", [4, 7, 10])
def test_etree_matches_reference_compute_etree(both, n):
    Mp, Mi = synth.grid_graph(n)
    rng = np.random.default_rng(n)
    for perm in (np.arange(n ** 3), rng.permutation(n ** 3)):
        assert np.array_equal(both.etree(Mp, Mi, perm, "port"), both.etree(Mp, Mi, perm, "reference"))
