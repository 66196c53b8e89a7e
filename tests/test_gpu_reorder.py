"""GPU parity of stage (c): sub-mesh extraction (createMultiSubMesh / getCoarseMesh), batched AMD
(amd_order semantics, bit-exact to oracle/amd_restated.c) and permuteFineHMDNodes."""
import numpy as np
import pytest

from parth_b200 import synth
from tests import scenario_defs as sd

pytestmark = pytest.mark.gpu


def _pair(gpu_api, port, tree_key, Mp, Mi, tree):
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = tree.levels; g.reorder_type = gpu_api.AMD
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    o = port.CpuParth("port"); o.set_options(reorder_type=1, nd_levels=tree.levels)
    g.import_tree(tree, Mp, Mi); o.import_tree(tree, Mp, Mi)
    g.setMesh(len(Mp) - 1, Mp, Mi); o.setMesh(len(Mp) - 1, Mp, Mi)
    return g, o


@pytest.mark.parametrize("key", [("geo", 16, 5), ("fish",), ("g12",)])
def test_submesh_fine_and_coarse(gpu_api, port, key):
    tree, Mp, Mi = sd.tree_of(dict(tree=key))
    g, o = _pair(gpu_api, port, key, Mp, Mi, tree)
    nodes = [x for x in range(tree.nodes) if tree.node_ptr[x + 1] > tree.node_ptr[x]]
    rng = np.random.default_rng(3)
    for nd in [0, 1, 2] + rng.choice(nodes, 6).tolist() + [nodes[-1]]:
        for coarse in (False, True):
            a, b = g.stage_submesh(nd, coarse), o.submesh(nd, coarse)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (nd, coarse)


def _random_sym_graph(rng, n, m):
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = r != c
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
    return Mp, Mi


def test_amd_batch_matches_restatement(gpu_api, port):
    rng = np.random.default_rng(11)
    graphs = [synth.grid_graph(k) for k in (2, 3, 5, 8)]
    graphs += [_random_sym_graph(rng, int(rng.integers(1, 300)), int(rng.integers(0, 900))) for _ in range(24)]
    graphs.append((np.zeros(6, np.int32), np.zeros(0, np.int32)))          # no edges: identity
    graphs.append(_random_sym_graph(rng, 400, 30000))                      # dense rows (> 10*sqrt(n))
    star = np.concatenate([np.zeros(199, np.int64), np.arange(1, 200)]), np.concatenate([np.arange(1, 200), np.zeros(199, np.int64)])
    graphs.append(synth.coo_to_csc(200, star[0], star[1])[1:])              # one dense hub
    g = gpu_api.ParthAPI()
    got = g.amd_batch(graphs)
    for (Mp, Mi), p in zip(graphs, got):
        want = port.amd(Mp, Mi)
        assert sorted(p.tolist()) == list(range(len(Mp) - 1))
        assert np.array_equal(p, want)


@pytest.mark.parametrize("key", [("geo", 16, 5), ("fish",), ("g12",)])
def test_permute_fine_matches_oracle(gpu_api, port, key):
    tree, Mp, Mi = sd.tree_of(dict(tree=key))
    g, o = _pair(gpu_api, port, key, Mp, Mi, tree)
    nodes = [x for x in range(tree.nodes) if tree.node_ptr[x + 1] > tree.node_ptr[x]]
    rng = np.random.default_rng(5)
    pick = sorted(set([0, nodes[-1]] + rng.choice(nodes, 9).tolist()))
    g.stage_permute_fine(pick)
    assert o.permute_fine(pick) == 0
    gt, ot = g.export_tree(), o.export_tree()
    assert np.array_equal(gt["labels"], ot.labels)
    assert np.array_equal(gt["mesh_perm"], ot.mesh_perm)
    assert sorted(gt["mesh_perm"].tolist()) == list(range(tree.M_n))


def test_metis_leaf_ordering_is_refused_not_faked(gpu_api):
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = gpu_api.METIS
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    g.import_tree(tree, Mp, Mi); g.setMesh(len(Mp) - 1, Mp, Mi)
    with pytest.raises(gpu_api.ParthError) as e:
        g.stage_permute_fine([31])
    assert e.value.code == -6


def test_amd_batch_stress_is_deterministic_and_exact(gpu_api, port):
    """compute-sanitizer is not available on the GPU pool: the warp-level synchronisation discipline of
    the cooperative AMD kernel is stressed instead -- 160 graphs of mixed shapes (staged in shared memory
    and slab resident, chunked elements, dense rows, supervariable-rich meshes) in one batch, three
    runs, every permutation equal to the oracle's every time"""
    rng = np.random.default_rng(77)
    graphs = []
    for t in range(150):
        n = int(rng.integers(2, 700))
        m = int(rng.integers(0, 6 * n))
        graphs.append(_random_sym_graph(rng, n, m))
    graphs += [synth.grid_graph(k) for k in (4, 7, 10, 13, 20, 27)]
    graphs += [_random_sym_graph(rng, 3000, 40000), _random_sym_graph(rng, 9000, 30000)]
    graphs += [synth.grid_graph(34), synth.grid_graph(40)]     # 39 304 / 64 000 vertices: slab resident
    want = [port.amd(Mp, Mi) for Mp, Mi in graphs]
    g = gpu_api.ParthAPI()
    for rep in range(3):
        got = g.amd_batch(graphs)
        for k, (p, w) in enumerate(zip(got, want)):
            assert np.array_equal(p, w), (rep, k, len(w))
