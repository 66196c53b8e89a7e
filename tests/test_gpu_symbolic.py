"""GPU parity of stage (d): elimination tree, column counts, nnz(L) through parth_b200_symbolic,
against the known answers of SURVEY.md Appendix B.3 and the oracle restatement
(oracle/symbolic_oracle.c: Liu etree as reference src/utils/etree.cpp:39-115, GNP column counts)."""
import numpy as np
import pytest

from parth_b200 import synth
from tests import scenario_defs as sd
from tests.golden import fixtures

pytestmark = pytest.mark.gpu

KNOWN = [(4, "id", 883), (4, "aff", 793), (10, "id", 91909), (10, "aff", 105344), (20, "id", 3055619),
         (20, "aff", 3782404)]


@pytest.mark.parametrize("n,kind,lnz", KNOWN)
def test_known_answers(gpu_api, port, n, kind, lnz):
    Mp, Mi = synth.grid_graph(n)
    N = n ** 3
    perm = np.arange(N) if kind == "id" else (7919 * np.arange(N) + 13) % N
    g = gpu_api.ParthAPI()
    g.setMesh(N, Mp, Mi)
    parent, cc, got = g.symbolic(perm)
    assert got == lnz and int(cc.sum()) == lnz
    op = port.etree(Mp, Mi, perm)
    olnz, occ = port.colcount(Mp, Mi, perm, op)
    assert np.array_equal(parent, op) and np.array_equal(cc, occ) and olnz == lnz


def test_fish_identity(gpu_api):
    N, Ap, Ai = fixtures.matrix("original")
    g = gpu_api.ParthAPI()
    g.setMatrix(N, Ap, Ai, 1)
    parent, cc, lnz = g.symbolic(np.arange(N))
    assert lnz == 1332839 and 2 * lnz - N == 2654892
    assert int(cc.sum()) == lnz and int((parent == -1).sum()) >= 1


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_vs_oracle(gpu_api, port, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    m = int(rng.integers(0, 5 * n))
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = r != c
    _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
    perm = rng.permutation(n)
    g = gpu_api.ParthAPI()
    g.setMesh(n, Mp, Mi)
    parent, cc, lnz = g.symbolic(perm)
    op = port.etree(Mp, Mi, perm)
    olnz, occ = port.colcount(Mp, Mi, perm, op)
    assert np.array_equal(parent, op) and np.array_equal(cc, occ) and lnz == olnz


@pytest.mark.parametrize("extra_edges", [0, 25])
@pytest.mark.parametrize("key", [("geo", 16, 5), ("fish",), ("g12",)])
def test_tree_aware_path_matches_oracle(gpu_api, port, key, extra_edges):
    """perm = None: the current mesh_perm, scheduled by separator-tree node.  With extra edges the
    separator property is broken (what lagging leaves behind) and the schedule must stay exact."""
    tree, Mp, Mi = sd.tree_of(dict(tree=key))
    M = len(Mp) - 1
    if extra_edges:
        rng = np.random.default_rng(9)
        Mp, Mi = synth.add_edges(Mp, Mi, rng.integers(0, M, size=(extra_edges, 2)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = tree.levels
    g.import_tree(tree, Mp, Mi)
    g.setMesh(M, Mp, Mi)
    parent, cc, lnz = g.symbolic(None)
    perm = g.mesh_perm()
    op = port.etree(Mp, Mi, perm)
    olnz, occ = port.colcount(Mp, Mi, perm, op)
    assert np.array_equal(parent, op) and np.array_equal(cc, occ) and lnz == olnz
    # the same permutation passed explicitly takes the schedule-free path: same result
    p2, c2, l2 = g.symbolic(perm)
    assert np.array_equal(p2, op) and np.array_equal(c2, occ) and l2 == olnz


def test_bad_permutation_is_rejected(gpu_api):
    Mp, Mi = synth.grid_graph(4)
    g = gpu_api.ParthAPI()
    g.setMesh(64, Mp, Mi)
    bad = np.arange(64); bad[3] = 64
    with pytest.raises(gpu_api.ParthError):
        g.symbolic(bad)


def test_device_pointer_variant_hands_off_on_the_device(gpu_api):
    """parth_b200_symbolic_dev: permutation in, etree and column counts out, all on the device (the hand-off
    to a device sparse Cholesky); must equal the host-pointer call"""
    import torch
    n = 12
    Mp, Mi = synth.grid_graph(n)
    N = n ** 3
    perm = ((7919 * np.arange(N) + 13) % N).astype(np.int32)
    g = gpu_api.ParthAPI()
    g.setMesh(N, Mp, Mi)
    parent, cc, lnz = g.symbolic(perm)
    dperm = torch.from_numpy(perm).cuda()
    dparent = torch.empty(N, dtype=torch.int32, device="cuda")
    dcc = torch.empty(N, dtype=torch.int32, device="cuda")
    got = g.symbolicDev(dperm.data_ptr(), dparent.data_ptr(), dcc.data_ptr())
    torch.cuda.synchronize()
    assert got == lnz
    assert np.array_equal(dparent.cpu().numpy(), parent) and np.array_equal(dcc.cpu().numpy(), cc)
    assert g.symbolicDev(dperm.data_ptr(), 0, 0) == lnz     # only nnz(L) wanted


@pytest.mark.parametrize("key", [("geo", 16, 5), ("fish",), ("g12",), ("rand", 7), ("rand", 8), ("clique",)])
def test_supernodes_match_oracle(gpu_api, port, key):
    """(f4) fundamental supernodes + supernodal etree on the device against oracle/symbolic_oracle.c
    (itself checked against a brute-force symbolic factorisation, tests/test_oracle_symbolic.py)"""
    g = gpu_api.ParthAPI()
    if key[0] == "rand":
        rng = np.random.default_rng(key[1])
        n = 3000
        r, c = rng.integers(0, n, 4 * n), rng.integers(0, n, 4 * n)
        keep = r != c
        _, Mp, Mi = synth.coo_to_csc(n, np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))
        perm = rng.permutation(n)
    elif key[0] == "clique":
        n = 5000  # one supernode, longer than a scan tile
        Mp = (np.arange(n + 1, dtype=np.int64) * (n - 1)).astype(np.int32)
        full = np.tile(np.arange(n, dtype=np.int32), n).reshape(n, n)
        Mi = full[~np.eye(n, dtype=bool)]
        perm = np.arange(n)
    else:
        tree, Mp, Mi = sd.tree_of({"tree": key})
        perm = synth.assemble_mesh_perm(tree.meta, tree.node_ptr, tree.dofs, tree.labels)
    n = len(Mp) - 1
    g.setMesh(n, Mp, Mi)
    parent, cc, lnz = g.symbolic(perm)
    got = g.supernodes(parent, cc)
    want = port.supernodes(parent, cc)
    for k in ("super", "supermap", "sparent"):
        assert np.array_equal(got[k], want[k]), k
    assert (got["nsuper"], got["ssize"], got["xsize"]) == (want["nsuper"], want["ssize"], want["xsize"])
    if key[0] == "clique":
        assert got["nsuper"] == 1 and got["xsize"] == n * n


def test_supernodes_device_pointers(gpu_api, port):
    """hand-off without leaving the device: symbolic_dev -> supernodes_dev"""
    import torch
    tree, Mp, Mi = sd.tree_of({"tree": ("geo", 16, 5)})
    n = len(Mp) - 1
    g = gpu_api.ParthAPI()
    g.import_tree(tree, Mp, Mi)
    g.setMesh(n, Mp, Mi)
    dpar = torch.empty(n, dtype=torch.int32, device="cuda")
    dcc = torch.empty(n, dtype=torch.int32, device="cuda")
    dsup = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    dmap = torch.empty(n, dtype=torch.int32, device="cuda")
    dsp = torch.empty(n, dtype=torch.int32, device="cuda")
    lnz = g.symbolicDev(0, dpar.data_ptr(), dcc.data_ptr())
    ns, ssize, xsize = g.supernodesDev(n, dpar.data_ptr(), dcc.data_ptr(), dsup.data_ptr(), dmap.data_ptr(), dsp.data_ptr())
    g.synchronize()
    want = port.supernodes(dpar.cpu().numpy(), dcc.cpu().numpy())
    assert ns == want["nsuper"] and ssize == want["ssize"] and xsize == want["xsize"] and xsize >= lnz
    assert np.array_equal(dsup[:ns + 1].cpu().numpy(), want["super"])
    assert np.array_equal(dmap.cpu().numpy(), want["supermap"]) and np.array_equal(dsp[:ns].cpu().numpy(), want["sparent"])
    # outputs are optional
    assert g.supernodesDev(n, dpar.data_ptr(), dcc.data_ptr()) == (ns, ssize, xsize)


@pytest.mark.parametrize("key,hub", [(("geo", 16, 5), 0), (("geo", 16, 5), 45), (("g12",), 20), (("fish",), 60)])
def test_batch_etree_kernel_equals_sequential_walk_and_oracle(gpu_api, port, key, hub, monkeypatch):
    """stage (d) etree, valid separator tree: the batch-parallel node kernel (128 columns at a time, climb +
    atomicMin rounds) against the sequential per-column walk (PARTH_B200_ETREE_SEQ) and the oracle.  `hub` > 0
    gives one vertex `hub` extra neighbours inside its own tree node, so a row longer than the 32-entry prefetch
    slot makes its batch fall back to the sequential walk on the same state."""
    tree, Mp, Mi = sd.tree_of(dict(tree=key))
    M = len(Mp) - 1
    if hub:
        sizes = np.diff(tree.node_ptr)
        node = int(np.argmax(sizes))
        d = tree.node_dofs(node)
        assert len(d) > hub + 1
        Mp, Mi = synth.add_edges(Mp, Mi, [(int(d[0]), int(x)) for x in d[1:hub + 1]])
    outs = []
    for mode in ("batch", "batch_global_links", "sequential"):
        monkeypatch.delenv("PARTH_B200_ETREE_SMEM_KB", raising=False)
        if mode == "batch_global_links":  # slices beyond the shared-memory budget keep their links in global memory
            monkeypatch.setenv("PARTH_B200_ETREE_SMEM_KB", "1")
        if mode == "sequential":
            monkeypatch.setenv("PARTH_B200_ETREE_SEQ", "1")
        g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = tree.levels
        g.import_tree(tree, Mp, Mi)
        g.setMesh(M, Mp, Mi)
        outs.append(g.symbolic(None))
        assert g.stats()["symbolic_path"] == 2
        perm = g.mesh_perm()
    op = port.etree(Mp, Mi, perm)
    olnz, occ = port.colcount(Mp, Mi, perm, op)
    for parent, cc, lnz in outs:
        assert np.array_equal(parent, op) and np.array_equal(cc, occ) and lnz == olnz
