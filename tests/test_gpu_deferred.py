"""GPU: the stream-ordered device-pointer entry points (parth_b200_set_matrix_dev /
parth_b200_compute_permutation_dev) against the host-pointer entry points on the same update
sequence.  In the steady state set_matrix_dev returns with its work in flight and
compute_permutation_dev expands the permutation before the flags are read; every observable result
(mesh graph, changed edges, dirty sets, reuse, permutation, tree) must be identical to the
synchronous path, which tests/test_gpu_scenarios.py / test_gpu_fullsize.py pin against the oracle
(ParthAPI::setMatrix + computePermutation, reference src/Parth/ParthAPI.cpp:100-124,191-279)."""
import numpy as np
import pytest

from parth_b200 import workloads

pytestmark = pytest.mark.gpu


def _handles(gpu_api, st, **opts):
    out = []
    for _ in range(2):
        g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = st.levels
        for k, v in opts.items():
            setattr(g, k, v)
        g.import_tree(st.tree(), st.Mp, st.Mi)
        out.append(g)
    return out


def _update_dev(g, torch, N, Ap, Ai, dim=1):
    dAp = torch.from_numpy(Ap).cuda(); dAi = torch.from_numpy(Ai).cuda()
    dperm = torch.empty(N * dim, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)
    rc = g.computePermutationDev(dperm.data_ptr(), dim)
    g.synchronize()
    return rc, dperm.cpu().numpy(), (dAp, dAi)


def _same_state(a, b):
    assert a.getNumChanges() == b.getNumChanges() and a.getReuse() == b.getReuse()
    assert np.array_equal(a.changed_edges(), b.changed_edges())
    fa, ca = a.dirty(); fb, cb = b.dirty()
    assert fa.tolist() == fb.tolist() and ca.tolist() == cb.tolist()
    ta, tb = a.export_tree(), b.export_tree()
    for key in ("node_ptr", "dofs", "labels", "dof2node", "mesh_perm"):
        assert np.array_equal(ta[key], tb[key]), key


@pytest.mark.parametrize("variant", ["lag", "nolag", "aggr"])
def test_deferred_updates_equal_synchronous_updates(gpu_api, variant):
    import torch
    st = workloads.PoissonUpdateStream(n=40, levels=7, edges_per_update=60)  # level-6 parents: ids >= 62, lag-dropped
    N = st.N
    opts = {"lag": {}, "nolag": {"lagging": False},
            "aggr": {"activate_aggressive_reuse": True, "relocate_lvl": 7, "reorder_type": gpu_api.AMD}}[variant]
    a, b = _handles(gpu_api, st, **opts)
    deferred = speculative = 0
    seq = [0, 1, 1, 2, 3, 3, 4]  # repeated indices: an update that changes nothing
    for dim, k in zip([1, 1, 3, 1, 2, 1, 1], seq):
        Ap, Ai = st.matrix(k)
        a.setMatrix(N, Ap, Ai, 1)
        rc_a, perm_a = a.computePermutation(dim)
        rc_b, perm_b, keep = _update_dev(b, torch, N, Ap, Ai, dim)
        s = b.stats()
        deferred += s["build_deferred"]; speculative += s["expand_speculative"]
        assert rc_a == rc_b == 0
        assert np.array_equal(perm_a, perm_b), (variant, k)
        assert np.array_equal(np.sort(perm_b[::dim] // dim), np.arange(N, dtype=np.int32))
        _same_state(a, b)
        Mpa, Mia = a.mesh(); Mpb, Mib = b.mesh()
        assert np.array_equal(Mpa, Mpb) and np.array_equal(Mia, Mib)
    # the first update proves the graph from scratch; the steady state behind it is deferred
    assert deferred >= len(seq) - 2
    if variant == "lag":
        assert speculative >= len(seq) - 2  # nothing is re-ordered: the early expansion always stands
    else:
        assert 0 < speculative < deferred    # only the no-change updates keep it


def test_deferred_build_falls_back_when_the_device_raises_a_flag(gpu_api):
    """steady state, then inputs the fast path cannot accept: a one-sided (asymmetric) entry, many
    changed rows, an unsorted column, an out-of-range row index (reported by the settling call)"""
    import torch
    st = workloads.PoissonUpdateStream(n=24, levels=4, edges_per_update=20)
    N = st.N
    a, b = _handles(gpu_api, st)
    keep = []

    def both(Ap, Ai):
        a.setMatrix(N, Ap, Ai, 1)
        rc_a, perm_a = a.computePermutation(1)
        rc_b, perm_b, k = _update_dev(b, torch, N, Ap, Ai, 1)
        keep.append(k)
        assert rc_a == rc_b and np.array_equal(perm_a, perm_b)
        _same_state(a, b)
        Mpa, Mia = a.mesh(); Mpb, Mib = b.mesh()
        assert np.array_equal(Mpa, Mpb) and np.array_equal(Mia, Mib)

    both(*st.matrix(0))
    both(*st.matrix(1))
    assert b.stats()["build_deferred"] == 1
    # one-sided entry (5, 3000): the general path symmetrises it (ParthAPI.cpp:381-395)
    Ap, Ai = workloads.insert_entries(st.Ap, st.Ai, [5], [3000])
    both(Ap, Ai)
    assert b.stats()["build_path"] == 2 and b.stats()["build_deferred"] == 0
    both(*st.matrix(2))
    both(*st.matrix(3))
    assert b.stats()["build_deferred"] == 1
    # every row changes: far more than the changed-row capacity of the incremental proof
    rng = np.random.default_rng(3)
    u = rng.permutation(N)[: (N // 2) * 2].reshape(-1, 2)
    u = u[np.abs(u[:, 0] - u[:, 1]) > 2 * 24 * 24]
    Ap, Ai = workloads.insert_entries(st.Ap, st.Ai, np.concatenate([u[:, 0], u[:, 1]]), np.concatenate([u[:, 1], u[:, 0]]))
    both(Ap, Ai)
    both(*st.matrix(4))
    both(*st.matrix(5))
    assert b.stats()["build_deferred"] == 1
    # unsorted column: swap two entries of one column
    Ap, Ai = st.matrix(6)
    Ai = Ai.copy(); p = Ap[100]; Ai[p], Ai[p + 1] = Ai[p + 1], Ai[p]
    both(Ap, Ai)
    both(*st.matrix(7))
    both(*st.matrix(8))
    # out-of-range row index: the host-pointer call reports it at once, the deferred one when it settles
    Ap, Ai = st.matrix(9)
    Ai = Ai.copy(); Ai[Ap[N] - 1] = N + 7
    with pytest.raises(gpu_api.ParthError):
        a.setMatrix(N, Ap, Ai, 1)
    dAp = torch.from_numpy(Ap).cuda(); dAi = torch.from_numpy(Ai).cuda()
    dperm = torch.empty(N, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with pytest.raises(gpu_api.ParthError):
        b.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)
        b.computePermutationDev(dperm.data_ptr(), 1)


def test_calls_between_the_deferred_build_and_the_permutation_settle_it(gpu_api):
    import torch
    st = workloads.PoissonUpdateStream(n=24, levels=4, edges_per_update=20)
    N = st.N
    (g, ref) = _handles(gpu_api, st)
    bufs = []
    for k in range(2):
        Ap, Ai = st.matrix(k)
        rc, perm, kp = _update_dev(g, torch, N, Ap, Ai)
        bufs.append(kp)
        ref.setMatrix(N, Ap, Ai, 1); ref.computePermutation(1)
    Ap, Ai = st.matrix(2)
    dAp = torch.from_numpy(Ap).cuda(); dAi = torch.from_numpy(Ai).cuda()
    torch.cuda.synchronize()
    g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)
    # reading the mesh in between must see the new graph
    assert g.mesh_dims() == (N, len(Ai) - N)
    Mp, Mi = g.mesh()
    ref.setMatrix(N, Ap, Ai, 1)
    rMp, rMi = ref.mesh()
    assert np.array_equal(Mp, rMp) and np.array_equal(Mi, rMi)
    assert g.stats()["build_deferred"] == 1
    dperm = torch.empty(N, dtype=torch.int32, device="cuda")
    assert g.computePermutationDev(dperm.data_ptr(), 1) == 0
    g.synchronize()
    _, rperm = ref.computePermutation(1)
    assert np.array_equal(dperm.cpu().numpy(), rperm)
    _same_state(g, ref)
    # two builds in a row: the second one replaces the first
    A3 = st.matrix(3); A4 = st.matrix(4)
    d3 = [torch.from_numpy(x).cuda() for x in A3]; d4 = [torch.from_numpy(x).cuda() for x in A4]
    torch.cuda.synchronize()
    g.setMatrixDev(N, d3[0].data_ptr(), d3[1].data_ptr(), len(A3[1]), 1)
    g.setMatrixDev(N, d4[0].data_ptr(), d4[1].data_ptr(), len(A4[1]), 1)
    assert g.computePermutationDev(dperm.data_ptr(), 1) == 0
    g.synchronize()
    ref.setMatrix(N, A3[0], A3[1], 1)
    ref.setMatrix(N, A4[0], A4[1], 1)
    _, rperm = ref.computePermutation(1)
    assert np.array_equal(dperm.cpu().numpy(), rperm)
    _same_state(g, ref)
