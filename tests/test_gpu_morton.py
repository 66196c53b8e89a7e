"""GPU: ReorderingType MORTON_CODE (parth_b200/csrc/morton.cu) against oracle/morton_oracle.py, bit-exact keys and
order; leaf ordering through stage_permute_fine; properties at config-4 size (50 M points).  The reference has no
implementation of this type (HMD.cpp:82-83): parity is against the restated definition only."""
import numpy as np
import pytest

from oracle import morton_oracle as mo
from tests import scenario_defs as sd

pytestmark = pytest.mark.gpu


def _cloud(rng, n, kind):
    if kind == "uniform":
        return rng.random((n, 3))
    if kind == "ties":                      # few distinct points: stability decides
        return rng.integers(0, 4, (n, 3)).astype(np.float64)
    if kind == "flat":                      # a degenerate axis and a negative range
        x = rng.normal(size=(n, 3)) * [1e3, 0, 1e-3] - [5e2, 7.0, 0]
        return x
    if kind == "nan":
        x = rng.random((n, 3))
        x[rng.integers(0, n, max(n // 10, 1)), rng.integers(0, 3, max(n // 10, 1))] = np.nan
        return x
    raise KeyError(kind)


@pytest.mark.parametrize("n", [1, 2, 31, 4095, 4096, 4097, 100003])
@pytest.mark.parametrize("kind,bits", [("uniform", 21), ("uniform", 10), ("ties", 21), ("flat", 21), ("nan", 16)])
def test_keys_and_order_match_oracle(gpu_api, n, kind, bits):
    rng = np.random.default_rng(n * 7 + bits)
    xyz = _cloud(rng, n, kind)
    g = gpu_api.ParthAPI()
    g.setCoordinates(xyz, bits)
    want, k = mo.order(xyz, bits)
    got = g.mortonOrder()
    assert np.array_equal(g.mortonKeys(), k[want])
    assert np.array_equal(got, want)
    # an Eigen-style column-major V gives the same answer
    g.setCoordinates(np.asfortranarray(xyz), bits)
    assert np.array_equal(g.mortonOrder(), want)


def test_hybrid_and_full_sort_agree(gpu_api):
    """clustered points (long runs of equal upper key halves) overflow the hybrid's run limit and take every radix
    pass; spread points take four: both orders equal the oracle's"""
    rng = np.random.default_rng(4)
    n = 300000
    spread = rng.random((n, 3))
    clustered = np.vstack([rng.random((n - 2, 3)) * 1e-3, [[0, 0, 0], [1, 1, 1]]])
    g = gpu_api.ParthAPI()
    for xyz, passes in ((spread, 4), (clustered, None)):
        g.setCoordinates(xyz)
        want, k = mo.order(xyz)
        assert np.array_equal(g.mortonOrder(), want)
        assert np.array_equal(g.mortonKeys(), k[want])
        got = g.stats()["morton_passes"]
        assert got == passes if passes else got > 4


def test_empty_and_bad_arguments(gpu_api):
    g = gpu_api.ParthAPI()
    g.setCoordinates(np.zeros((0, 3)))
    assert len(g.mortonOrder()) == 0
    with pytest.raises(gpu_api.ParthError):
        g.setCoordinates(np.zeros((4, 3)), bits=22)


@pytest.mark.parametrize("key", [("geo", 16, 5), ("fish",), ("g12",)])
def test_permute_fine_morton(gpu_api, key):
    tree, Mp, Mi = sd.tree_of(dict(tree=key))
    M = tree.M_n
    rng = np.random.default_rng(9)
    xyz = rng.random((M, 3))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = tree.levels; g.reorder_type = gpu_api.MORTON_CODE
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    g.import_tree(tree, Mp, Mi); g.setMesh(M, Mp, Mi)
    nodes = [x for x in range(tree.nodes) if tree.node_ptr[x + 1] > tree.node_ptr[x]]
    pick = sorted(set([0, nodes[-1]] + rng.choice(nodes, 9).tolist()))
    with pytest.raises(gpu_api.ParthError):   # no coordinates yet: refused, never faked
        g.stage_permute_fine(pick)
    g.setCoordinates(xyz)
    before = g.export_tree()
    g.stage_permute_fine(pick)
    after = g.export_tree()
    order, _ = mo.order(xyz)
    rank = np.empty(M, np.int32); rank[order] = np.arange(M, dtype=np.int32)
    labels, mesh_perm = before["labels"].copy(), before["mesh_perm"].copy()
    meta = np.asarray(tree.meta).reshape(-1, 7)
    for nd in pick:
        a, b = tree.node_ptr[nd], tree.node_ptr[nd + 1]
        dofs = np.asarray(tree.dofs[a:b])
        lp = mo.node_order(rank, dofs)                   # perm[new] = old local id
        labels[a + lp] = np.arange(b - a, dtype=np.int32)  # HMDNode::setPermutedNewLabel, HMD.h:70-79
        mesh_perm[meta[nd, 4] + np.arange(b - a)] = dofs[lp]  # Integrator.cpp:886-891
    assert np.array_equal(after["labels"], labels)
    assert np.array_equal(after["mesh_perm"], mesh_perm)
    assert sorted(after["mesh_perm"].tolist()) == list(range(M))


def test_config4_size_properties(gpu_api):
    """50 M points (BASELINE config 4): the order is a permutation and the keys along it are monotone, ties in
    ascending point id; keys of a sample agree with the oracle"""
    import torch
    n = 50_000_000
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    xyz = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=gen)
    g = gpu_api.ParthAPI()
    torch.cuda.synchronize()
    g.setCoordinatesDev(xyz.data_ptr(), n)
    g.synchronize()
    dptr, cnt = g.mortonOrderDev()
    assert cnt == n
    perm = torch.from_numpy(g.mortonOrder()).cuda()
    seen = torch.zeros(n, dtype=torch.int32, device="cuda")
    seen.index_add_(0, perm.long(), torch.ones(n, dtype=torch.int32, device="cuda"))
    assert bool((seen == 1).all())
    keys = torch.from_numpy(g.mortonKeys().view(np.int64)).cuda()   # < 2^63: int64 order is the u64 order
    assert bool((keys[1:] >= keys[:-1]).all())
    tie = keys[1:] == keys[:-1]
    assert bool((perm[1:][tie] > perm[:-1][tie]).all())
    # a sample of the keys against the oracle (the box needs the global min / max)
    lo, hi = xyz.min(0).values.cpu().numpy(), xyz.max(0).values.cpu().numpy()
    idx = torch.randint(0, n, (20000,), device="cuda")
    sample = xyz[perm[idx].long()].cpu().numpy()
    want = mo.keys(np.vstack([lo, hi, sample]))[2:]
    assert np.array_equal(keys[idx].cpu().numpy().view(np.uint64), want)
    assert g.stats()["morton_passes"] == 4   # hybrid: upper 32 bits by radix passes, short runs fixed in place


def _slices_follow_rank(t, rank, nodes):
    meta = np.asarray(t["meta"]).reshape(-1, 7)
    for nd in nodes:
        a, b = t["node_ptr"][nd], t["node_ptr"][nd + 1]
        sl = t["mesh_perm"][meta[nd, 4]: meta[nd, 4] + (b - a)]
        assert sorted(sl.tolist()) == t["dofs"][a:b].tolist()
        assert np.all(np.diff(rank[sl]) > 0), nd


def test_compute_permutation_with_morton_cold_and_warm(gpu_api):
    """cold start (built-in dissection) and a warm REPAIR update with ReorderingType MORTON_CODE: every node that is
    (re-)ordered follows the global Z-curve rank; the permutation stays a permutation"""
    from parth_b200 import synth
    n, L = 12, 4
    M = n ** 3
    Mp, Mi = synth.grid_graph(n)
    ind = np.arange(M)
    xyz = np.stack([ind // (n * n), (ind // n) % n, ind % n], 1).astype(np.float64)  # ind = z + y n + x n^2
    order, _ = mo.order(xyz)
    rank = np.empty(M, np.int32); rank[order] = np.arange(M, dtype=np.int32)
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = gpu_api.MORTON_CODE; g.lagging = False
    g.useBuiltinDecomposer(True)
    g.setCoarsePolicy(gpu_api.COARSE_REPAIR)
    g.setMesh(M, Mp, Mi)
    g.setCoordinates(xyz)
    st, perm = g.computePermutation(1)
    assert st == 0 and sorted(perm.tolist()) == list(range(M))
    t = g.export_tree()
    every = [x for x in range(t["nodes"]) if t["node_ptr"][x + 1] > t["node_ptr"][x]]
    _slices_follow_rank(t, rank, every)
    # warm: a few edges across the tree -> relocation into separators, the touched nodes are re-ordered
    rng = np.random.default_rng(3)
    P, I = synth.add_edges(Mp, Mi, rng.integers(0, M, size=(9, 2)))
    g.setMesh(M, P, I)
    st, perm = g.computePermutation(1)
    assert st == 0 and sorted(perm.tolist()) == list(range(M))
    t2 = g.export_tree()
    assert np.array_equal(t2["mesh_perm"], perm)
    every = [x for x in range(t2["nodes"]) if t2["node_ptr"][x + 1] > t2["node_ptr"][x]]
    _slices_follow_rank(t2, rank, every)   # untouched nodes kept their (Morton) order, touched ones were re-sorted


def test_dof_map_invalidates_coordinates(gpu_api):
    """a new-to-old DOF map changes the DOF set: stale coordinates must not be used silently"""
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    M = tree.M_n
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = gpu_api.MORTON_CODE
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    g.import_tree(tree, Mp, Mi)
    xyz = np.random.default_rng(2).random((M, 3))
    g.setCoordinates(xyz)
    g.setMesh(M, Mp, Mi, np.arange(M, dtype=np.int32))     # identity map, same size: still a new DOF set
    with pytest.raises(gpu_api.ParthError):
        g.stage_permute_fine([31])
    g.setCoordinates(xyz)
    g.stage_permute_fine([31])
