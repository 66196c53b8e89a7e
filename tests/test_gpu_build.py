"""GPU parity of stage (a), the CSC -> mesh graph build, through the C ABI
(parth_b200_set_matrix) against the oracle restatement of ParthAPI::computeMeshFromMatrix
(reference src/Parth/ParthAPI.cpp:366-418).  Bit-exact: Mp and Mi must be identical."""
import numpy as np
import pytest

from parth_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_mesh(port, N, Ap, Ai, dim):
    o = port.CpuParth("port")
    o.setMatrix(N, Ap, Ai, dim)
    return o.mesh()


def check(gpu_api, port, N, Ap, Ai, dim, expect_path=None, assume=False):
    g = gpu_api.ParthAPI()
    g.setAssumeSymmetric(assume)
    g.setMatrix(N, Ap, Ai, dim)
    Mp, Mi = g.mesh()
    oMp, oMi = oracle_mesh(port, N, Ap, Ai, dim)
    assert np.array_equal(Mp, oMp)
    assert np.array_equal(Mi, oMi)
    if expect_path is not None:
        assert g.stats()["build_path"] == expect_path
    return g


@pytest.mark.parametrize("n", [1, 2, 3, 7, 20, 33])
def test_grid_full_pattern_takes_filter_path(gpu_api, port, n):
    N, Ap, Ai = synth.poisson3d(n)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)


@pytest.mark.parametrize("n", [2, 7, 20])
def test_grid_lower_triangle_takes_general_path(gpu_api, port, n):
    N, Ap, Ai = synth.poisson3d(n, lower_only=True)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=2)


@pytest.mark.parametrize("dim", [2, 3, 4])
def test_kron_blocks(gpu_api, port, dim):
    N, Ap, Ai = synth.poisson3d(9)
    Nb, Bp, Bi = synth.kron_blocks(Ap, Ai, dim)
    check(gpu_api, port, Nb, Bp, Bi, dim, expect_path=1)


def test_kron_blocks_lower_only_dim3(gpu_api, port):
    N, Ap, Ai = synth.poisson3d(8, lower_only=True)
    Nb, Bp, Bi = synth.kron_blocks(Ap, Ai, 3)
    check(gpu_api, port, Nb, Bp, Bi, 3, expect_path=2)


def test_no_diagonal(gpu_api, port):
    N, Ap, Ai = synth.poisson3d(11, diagonal=False)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)


def test_unsorted_columns_and_duplicates(gpu_api, port):
    rng = np.random.default_rng(3)
    N, Ap, Ai = synth.poisson3d(10)
    Ai = Ai.copy()
    for c in range(N):  # shuffle every column: the ordering proof must fail -> general path
        rng.shuffle(Ai[Ap[c]:Ap[c + 1]])
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=2)
    # duplicates inside columns are removed like the reference's sort+unique does
    N, Ap, Ai = synth.poisson3d(6)
    Ai2 = np.repeat(Ai, 2)
    Ap2 = (Ap * 2).astype(np.int32)
    check(gpu_api, port, N, Ap2, Ai2, 1, expect_path=2)


def test_one_missing_mirror_entry_is_detected(gpu_api, port):
    N, Ap, Ai = synth.poisson3d(12)
    # drop a single strictly-lower entry: pattern no longer symmetric, counts differ by one
    col = 700
    seg = Ai[Ap[col]:Ap[col + 1]]
    k = int(np.nonzero(seg > col)[0][0]) + Ap[col]
    Ai2 = np.delete(Ai, k)
    Ap2 = Ap.copy()
    Ap2[col + 1:] -= 1
    check(gpu_api, port, N, Ap2, Ai2, 1, expect_path=2)
    # move an entry so that counts still match but the mirror is wrong
    Ai3 = Ai.copy()
    col2 = 900
    seg = Ai3[Ap[col2]:Ap[col2 + 1]]
    j = int(np.nonzero(seg > col2)[0][-1])
    seg[j] = min(N - 1, seg[j] + 3) if seg[j] + 3 not in seg else seg[j]
    check(gpu_api, port, N, Ap, Ai3, 1)


def test_random_sparse_with_empty_columns(gpu_api, port):
    rng = np.random.default_rng(11)
    for N, nnz in [(50, 60), (1000, 3000), (5000, 200000)]:
        r = rng.integers(0, N, nnz)
        c = rng.integers(0, N // 2, nnz) * 2  # odd columns stay empty
        _, Ap, Ai = synth.coo_to_csc(N, r, c)
        check(gpu_api, port, N, Ap, Ai, 1)
        # symmetrised version takes the filter path
        _, Ap, Ai = synth.coo_to_csc(N, np.concatenate([r, c]), np.concatenate([c, r]))
        check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)


def test_long_runs_of_empty_columns(gpu_api, port):
    # more empty columns than one tile's column window holds, in front, in the middle, at the end
    N = 30000
    blk = np.arange(40)
    r0, c0 = np.meshgrid(blk, blk)
    r = np.concatenate([r0.ravel() + 12000, r0.ravel() + 21000])
    c = np.concatenate([c0.ravel() + 12000, c0.ravel() + 21000])
    _, Ap, Ai = synth.coo_to_csc(N, r, c)
    check(gpu_api, port, N, Ap, Ai, 1)
    # moderate runs stay on the filter path
    N = 3000
    r = np.concatenate([r0.ravel() + 500, r0.ravel() + 2900])
    c = np.concatenate([c0.ravel() + 500, c0.ravel() + 2900])
    _, Ap, Ai = synth.coo_to_csc(N, r, c)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)


def test_dense_column_and_skewed_degrees(gpu_api, port):
    N = 6000
    hub = np.arange(1, N)
    r = np.concatenate([hub, np.zeros(N - 1, np.int64), np.arange(N)])
    c = np.concatenate([np.zeros(N - 1, np.int64), hub, np.arange(N)])
    _, Ap, Ai = synth.coo_to_csc(N, r, c)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)
    _, Ap, Ai = synth.coo_to_csc(N, hub, np.zeros(N - 1, np.int64))  # one dense column only
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=2)


def test_assume_symmetric_hint_skips_the_proof_but_not_the_result(gpu_api, port):
    N, Ap, Ai = synth.poisson3d(16)
    check(gpu_api, port, N, Ap, Ai, 1, expect_path=1, assume=True)


def test_out_of_range_index_is_an_error(gpu_api):
    N, Ap, Ai = synth.poisson3d(5)
    Ai = Ai.copy()
    Ai[17] = N + 5
    g = gpu_api.ParthAPI()
    with pytest.raises(gpu_api.ParthError):
        g.setMatrix(N, Ap, Ai, 1)


def test_golden_matrices(gpu_api, port):
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "matrices.npz")
    if not os.path.exists(path):
        pytest.skip("golden matrices not generated")
    from tests.golden import fixtures
    for name in ["original", "0", "1", "2", "3", "4"]:
        N, Ap, Ai = fixtures.matrix(name)
        g = check(gpu_api, port, N, Ap, Ai, 1, expect_path=1)
        Mp, Mi = g.mesh()
        gold = fixtures.build_golden(name)
        assert synth.hexhash(Mp) == gold["hash_Mp"] and synth.hexhash(Mi) == gold["hash_Mi"]


def test_large_grid_properties(gpu_api):
    """full-size-style check without the oracle: 96^3 grid, size-independent properties"""
    n = 96
    N, Ap, Ai = synth.poisson3d(n)
    g = gpu_api.ParthAPI()
    g.setMatrix(N, Ap, Ai, 1)
    Mp, Mi = g.mesh()
    assert g.stats()["build_path"] == 1
    assert Mp[-1] == len(Ai) - N == len(Mi)
    # strictly ascending rows, no diagonal, symmetric edge count
    row = np.repeat(np.arange(N), np.diff(Mp))
    assert np.all(Mi != row)
    inner = np.ones(len(Mi), bool)
    inner[Mp[1:-1][np.diff(Mp)[1:] > 0]] = False
    inner[0] = False
    assert np.all((np.diff(Mi.astype(np.int64), prepend=-1) > 0) | ~inner)
    assert np.array_equal(np.bincount(Mi, minlength=N), np.diff(Mp))


def _update_pair(gpu_api, port):
    from tests import scenario_defs as sd
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = 1
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    o = port.CpuParth("port"); o.set_options(reorder_type=1, nd_levels=5)
    g.import_tree(tree, Mp, Mi); o.import_tree(tree, Mp, Mi)
    return g, o, Mp, Mi


def test_incremental_symmetry_proof_over_consecutive_updates(gpu_api, port):
    """after one proven build, symmetry of the next graph is proven from the changed rows only
    (and the row diff is shared with change detection): symmetric updates stay on the filter path,
    a missing or a misplaced mirror entry is still caught and symmetrised like the reference does"""
    g, o, Mp, Mi = _update_pair(gpu_api, port)
    M = len(Mp) - 1
    rng = np.random.default_rng(5)

    def step(Ap, Ai, expect_path):
        g.setMatrix(M, Ap, Ai, 1); o.setMatrix(M, Ap, Ai, 1)
        gm, om = g.mesh(), o.mesh()
        assert np.array_equal(gm[0], om[0]) and np.array_equal(gm[1], om[1])
        assert g.stats()["build_path"] == expect_path
        st, perm = g.computePermutation(1)
        ost, operm = o.computePermutation(1)
        assert g.getNumChanges() == o.getNumChanges()
        assert np.array_equal(g.changed_edges(), o.changed_edges())
        assert g.getReuse() == o.getReuse()

    cur = (Mp, Mi)
    N, Ap, Ai = synth.graph_to_matrix(*cur)
    step(Ap, Ai, 1)                                   # full proof (snapshot came from import_tree)
    for k in range(3):                                # incremental proofs
        cur = synth.add_edges(cur[0], cur[1], rng.integers(0, M, size=(4, 2)))
        N, Ap, Ai = synth.graph_to_matrix(*cur)
        step(Ap, Ai, 1)
    step(Ap, Ai, 1)                                   # identical matrix again
    # drop one strictly-lower entry of the CSC pattern: asymmetric, counts differ
    col = 700
    seg = Ai[Ap[col]:Ap[col + 1]]
    kk = int(np.nonzero(seg > col)[0][0]) + Ap[col]
    Ai2 = np.delete(Ai, kk); Ap2 = Ap.copy(); Ap2[col + 1:] -= 1
    step(Ap2, Ai2, 2)
    step(Ap, Ai, 1)                                   # back to a symmetric pattern
    # move one entry so that the counts still match but a mirror is missing
    Ai3 = Ai.copy()
    col2 = 900
    seg = Ai3[Ap[col2]:Ap[col2 + 1]]
    j = int(np.nonzero(seg > col2)[0][-1])
    new = seg[j] + 3
    assert new < M and new not in seg
    seg[j] = new
    step(Ap, Ai3, 2)
    # removed edges only (symmetric): stays on the filter path
    src = np.repeat(np.arange(M), np.diff(cur[0]))
    pick = rng.choice(len(cur[1]), 10, replace=False)
    cur2 = synth.remove_edges(cur[0], cur[1], np.stack([src[pick], cur[1][pick]], axis=1))
    N, Ap4, Ai4 = synth.graph_to_matrix(*cur2)
    step(Ap4, Ai4, 1)
