"""GPU: ONE mesh over several ranks (row blocks of the graph build and of change detection, stage_sharded.cu).

The ranks are handles of this process driven by one thread each (parth_b200_comm_init_local: host barriers and
device-to-device copies instead of NCCL, the same call sequence) so the sharded logic is exercised on a single GPU;
scripts/rowblock_check.py runs the same comparison over NCCL with one process per GPU.  Every rank must return exactly
what a single handle holding the whole graph returns: permutation, changed edges (in the reference's order), dirty
sets, reuse ratio and the separator tree.
"""
import threading
from types import SimpleNamespace

import numpy as np
import pytest

from parth_b200 import api, synth

pytestmark = pytest.mark.gpu


def _tree(n, L):
    meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
    return SimpleNamespace(M_n=n ** 3, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)


def _configure(g, L, variant):
    g.verbose = False
    g.ND_levels = L
    g.reorder_type = api.AMD
    g.setCoarsePolicy(api.COARSE_REPAIR)
    if variant == "nolag":
        g.lagging = False
    elif variant == "aggr":
        g.lagging = True
        g.activate_aggressive_reuse = True
        g.relocate_lvl = L
    else:
        g.lagging = True


def _matrices(n, steps, edges, seed, remove_every=0):
    Mp, Mi = synth.grid_graph(n)
    M = n ** 3
    rng = np.random.default_rng(seed)
    cur = (Mp, Mi)
    out = [synth.graph_to_matrix(*cur)]
    for k in range(steps):
        if remove_every and k % remove_every == remove_every - 1:
            cur = (Mp, Mi)  # back to the base graph: every inserted edge is removed again
        else:
            cur = synth.add_edges(cur[0], cur[1], rng.integers(0, M, size=(edges, 2)))
        out.append(synth.graph_to_matrix(*cur))
    return (Mp, Mi), out


def _run_single(tree, base, mats, L, variant):
    ref = api.ParthAPI()
    _configure(ref, L, variant)
    ref.import_tree(tree, base[0], base[1])
    res = []
    for N, Ap, Ai in mats[1:]:
        ref.setMatrix(N, Ap, Ai, 1)
        rc, perm = ref.computePermutation(1)
        f, c = ref.dirty()
        res.append(dict(rc=rc, perm=perm.copy(), edges=ref.changed_edges().copy(), fine=f, coarse=c, reuse=ref.getReuse(),
                        tree=ref.export_tree()))
    return res


def _run_ranks(world, tree, mats, L, variant, dev_arrays=False):
    apis = [api.ParthAPI() for _ in range(world)]
    for g in apis:
        _configure(g, L, variant)
    api.ParthAPI.commInitLocal(apis)
    results = [None] * world
    errors = [None] * world

    def work(r):
        try:
            g = apis[r]
            assert g.commInfo() == (r, world, "local")
            g.import_tree(tree)
            N, Ap, Ai = mats[0]
            g.setMatrixBlock(N, int(Ap[N]), Ap, Ai, 1)
            g.acceptSnapshot()
            res = []
            for N, Ap, Ai in mats[1:]:
                if dev_arrays:
                    import torch
                    c0, c1 = g.blockRange(N, int(Ap[N]))
                    dAp = torch.from_numpy(Ap).cuda()
                    dAi = torch.from_numpy(Ai).cuda()
                    torch.cuda.synchronize()
                    g.setMatrixBlockDev(N, int(Ap[N]), dAp.data_ptr(), dAi.data_ptr(), int(Ap[c0]), int(Ap[c1]), 1)
                else:
                    g.setMatrix(N, Ap, Ai, 1)  # with a communicator: only the rank's block is uploaded
                rc, perm = g.computePermutation(1)
                f, c = g.dirty()
                res.append(dict(rc=rc, perm=perm.copy(), edges=g.changed_edges().copy(), fine=f, coarse=c,
                                reuse=g.getReuse(), tree=g.export_tree(), h2d=g.stats()["h2d_bytes"]))
            results[r] = res
        except Exception as e:  # noqa: BLE001
            errors[r] = e

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective"
    for e in errors:
        if e is not None:
            raise e
    return results


def _same(a, b):
    assert a["rc"] == b["rc"]
    assert np.array_equal(a["perm"], b["perm"])
    assert np.array_equal(a["edges"], b["edges"])
    assert np.array_equal(a["fine"], b["fine"]) and np.array_equal(a["coarse"], b["coarse"])
    assert a["reuse"] == b["reuse"]
    for k in ("labels", "dofs", "mesh_perm", "dof2node", "node_ptr", "meta"):
        assert np.array_equal(a["tree"][k], b["tree"][k]), k


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("variant", ["nolag", "aggr", "lag"])
def test_row_blocks_equal_single_handle(world, variant):
    n, L = 24, 6
    tree = _tree(n, L)
    base, mats = _matrices(n, steps=5, edges=40, seed=3 + world, remove_every=3)
    want = _run_single(tree, base, mats, L, variant)
    got = _run_ranks(world, tree, mats, L, variant)
    M = n ** 3
    for r in range(world):
        for a, b in zip(got[r], want):
            _same(a, b)
            assert sorted(a["perm"].tolist()) == list(range(M))
    assert any(len(w["edges"]) > 0 or len(w["fine"]) > 0 for w in want)  # the updates really dirty something
    if world > 1:
        # a rank uploads its block only: about 1 / world of the matrix
        N, Ap, Ai = mats[1]
        whole = 4 * (N + 1 + int(Ap[N]))
        assert max(res[0]["h2d"] for res in got) < whole * (1.0 / world + 0.15)


def test_row_blocks_device_arrays_and_no_change_update():
    n, L = 20, 5
    tree = _tree(n, L)
    base, mats = _matrices(n, steps=2, edges=30, seed=11)
    mats = mats + [mats[-1]]  # the last update changes nothing
    want = _run_single(tree, base, mats, L, "nolag")
    got = _run_ranks(2, tree, mats, L, "nolag", dev_arrays=True)
    for r in range(2):
        for a, b in zip(got[r], want):
            _same(a, b)
    assert len(want[-1]["edges"]) == 0 and want[-1]["reuse"] == 1.0


def test_block_result_is_the_ranks_slice_of_the_permutation():
    """computePermutationBlock: every rank's host receives entries [c0 * dim, c1 * dim) of the expanded permutation;
    the blocks tile it"""
    n, L, world, dim = 20, 5, 3, 3
    tree = _tree(n, L)
    base, mats = _matrices(n, steps=2, edges=30, seed=5)
    want = _run_single(tree, base, mats, L, "nolag")
    ref_full = np.repeat(want[-1]["perm"].astype(np.int64) * dim, dim) + np.tile(np.arange(dim), n ** 3)
    apis = [api.ParthAPI() for _ in range(world)]
    for g in apis:
        _configure(g, L, "nolag")
    api.ParthAPI.commInitLocal(apis)
    out, errors = [None] * world, [None] * world

    def work(r):
        try:
            g = apis[r]
            g.import_tree(tree)
            N, Ap, Ai = mats[0]
            g.setMatrixBlock(N, int(Ap[N]), Ap, Ai, 1)
            g.acceptSnapshot()
            for N, Ap, Ai in mats[1:]:
                g.setMatrix(N, Ap, Ai, 1)
                rc, first, blk = g.computePermutationBlock(dim)
            c0, c1 = g.blockRange(N, int(Ap[N]))
            assert rc == 0 and first == c0 * dim and len(blk) == (c1 - c0) * dim
            # the block plus the update's own small read-backs, not the whole permutation
            assert 4 * (c1 - c0) * dim <= g.stats()["d2h_bytes"] < 4 * (c1 - c0) * dim + 4 * n ** 3
            out[r] = (first, blk.copy())
        except Exception as e:  # noqa: BLE001
            errors[r] = e

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=300) for t in ts]
    for e in errors:
        if e is not None:
            raise e
    got = np.concatenate([b for _, b in sorted(out, key=lambda x: x[0])])
    assert np.array_equal(got, ref_full)


def test_row_blocks_reject_asymmetric_update():
    n, L = 16, 4
    tree = _tree(n, L)
    Mp, Mi = synth.grid_graph(n)
    N, Ap, Ai = synth.graph_to_matrix(Mp, Mi)
    # one directed entry (column 5 gains row 4000, column 4000 does not gain row 5): rows on different ranks
    from parth_b200.workloads import insert_entries
    Ap2, Ai2 = insert_entries(Ap, Ai, [5], [4000])
    apis = [api.ParthAPI() for _ in range(2)]
    for g in apis:
        _configure(g, L, "nolag")
    api.ParthAPI.commInitLocal(apis)
    codes = [None, None]

    def work(r):
        g = apis[r]
        g.import_tree(tree)
        g.setMatrixBlock(N, int(Ap[N]), Ap, Ai, 1)
        g.acceptSnapshot()
        try:
            g.setMatrix(N, Ap2, Ai2, 1)
            codes[r] = 0
        except api.ParthError as e:
            codes[r] = e.code

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(2)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert codes == [-4, -4]  # PARTH_B200_ERR_INPUT on every rank


def test_first_matrix_symmetry_is_proven_across_blocks():
    n, L = 16, 4
    tree = _tree(n, L)
    Mp, Mi = synth.grid_graph(n)
    N, Ap, Ai = synth.graph_to_matrix(Mp, Mi)
    from parth_b200.workloads import insert_entries
    Ap2, Ai2 = insert_entries(Ap, Ai, [7], [3900])
    for bad in (False, True):
        apis = [api.ParthAPI() for _ in range(3)]
        for g in apis:
            _configure(g, L, "nolag")
        api.ParthAPI.commInitLocal(apis)
        codes = [None] * 3

        def work(r):
            try:
                apis[r].import_tree(tree)
                a, b = (Ap2, Ai2) if bad else (Ap, Ai)
                apis[r].setMatrixBlock(N, int(a[N]), a, b, 1)
                codes[r] = 0
            except api.ParthError as e:
                codes[r] = e.code

        ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(3)]
        [t.start() for t in ts]
        [t.join(timeout=120) for t in ts]
        assert codes == ([-4] * 3 if bad else [0] * 3)
