"""GPU: REPAIR policy for dirty coarse nodes (the device's replacement for the reference's METIS
re-dissection, Integrator.cpp:812-849 -- not bit-comparable, so checked through invariants):
the permutation stays a permutation, the reported changed edges / dirty sets / reuse equal the
oracle's, and afterwards no problematic edge remains (the reference's own debug invariant,
Integrator.cpp:671-682), so an identical follow-up update reports full reuse."""
import numpy as np
import pytest

from parth_b200 import synth
from tests import scenario_defs as sd

pytestmark = pytest.mark.gpu


def problematic_edges(Mp, Mi, dof2node):
    src = np.repeat(np.arange(len(Mp) - 1), np.diff(Mp))
    a, b = dof2node[src], dof2node[Mi]
    small, big = np.minimum(a, b), np.maximum(a, b)
    x = big.copy()
    anc = np.zeros(len(x), bool)
    for _ in range(16):
        x = np.where(x > small, (x - 1) // 2, x)
        anc |= x == small
    return int(np.count_nonzero((small != big) & ~anc))


@pytest.mark.parametrize("k", [1, 7, 60, 400])
@pytest.mark.parametrize("rt", [0, 1])
@pytest.mark.parametrize("policy", ["REPAIR", "RELOCATE"])
def test_repair_restores_the_separator_property(gpu_api, port, k, rt, policy):
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    M = 16 ** 3
    rng = np.random.default_rng(k)
    P, I = synth.add_edges(Mp, Mi, rng.integers(0, M, size=(k, 2)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = rt; g.lagging = False
    g.setCoarsePolicy(getattr(gpu_api, "COARSE_" + policy))
    o = port.CpuParth("port"); o.set_options(reorder_type=rt, nd_levels=5, lagging=False)
    g.import_tree(tree, Mp, Mi); o.import_tree(tree, Mp, Mi)
    g.setMesh(M, P, I); o.setMesh(M, P, I)
    st, perm = g.computePermutation(1)
    o.computePermutation(1)
    assert st == 0
    assert np.array_equal(g.changed_edges(), o.changed_edges())
    gf, gc = g.dirty(); of, oc = o.dirty()
    assert gf.tolist() == of.tolist() and gc.tolist() == oc.tolist()
    if o.getReuse() > 0:
        assert g.getReuse() == o.getReuse()
    assert sorted(perm.tolist()) == list(range(M))
    t = g.export_tree()
    assert problematic_edges(P, I, t["dof2node"]) == 0
    assert np.array_equal(t["mesh_perm"], perm)
    # per node: labels are a permutation of 0..size-1, dofs ascending
    for x in range(t["nodes"]):
        a, b = t["node_ptr"][x], t["node_ptr"][x + 1]
        assert sorted(t["labels"][a:b].tolist()) == list(range(b - a))
        assert np.all(np.diff(t["dofs"][a:b]) > 0)
    # same graph again: nothing to do
    g.setMesh(M, P, I)
    st2, perm2 = g.computePermutation(1)
    assert st2 == 0 and g.getNumChanges() == 0 and g.getReuse() == 1.0 and np.array_equal(perm, perm2)


def test_report_status_keeps_the_snapshot_and_returns_a_valid_permutation(gpu_api, port):
    """COARSE_REPORT: PARTH_B200_NEEDS_REDISSECT hands back the tree's current (valid) permutation and does not
    advance the snapshot -- the same edges are reported again until the caller accepts the snapshot
    (the reference snapshots only after a finished update, ParthAPI.cpp:244-278)."""
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    M = 16 ** 3
    rng = np.random.default_rng(3)
    P, I = synth.add_edges(Mp, Mi, rng.integers(0, M, size=(60, 2)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = 1; g.lagging = False
    g.setCoarsePolicy(gpu_api.COARSE_REPORT)
    g.report_advances_snapshot = False
    g.import_tree(tree, Mp, Mi)
    g.setMesh(M, P, I)
    st, perm = g.computePermutation(1)
    assert st == gpu_api.NEEDS_REDISSECT
    assert sorted(perm.tolist()) == list(range(M))
    n1, e1 = g.getNumChanges(), g.changed_edges().copy()
    assert n1 > 0
    sp, si = g.snapshot()
    assert np.array_equal(sp, Mp) and np.array_equal(si, Mi)
    # second update on the same graph: the unresolved edges come back
    g.setMesh(M, P, I)
    st2, perm2 = g.computePermutation(1)
    assert st2 == gpu_api.NEEDS_REDISSECT and g.getNumChanges() == n1
    assert np.array_equal(g.changed_edges(), e1) and np.array_equal(perm, perm2)
    # the caller accepts: snapshot advances, the next identical update is clean
    g.acceptSnapshot()
    sp, si = g.snapshot()
    assert np.array_equal(sp, P) and np.array_equal(si, I)
    g.setMesh(M, P, I)
    st3, perm3 = g.computePermutation(1)
    assert st3 == 0 and g.getNumChanges() == 0 and np.array_equal(perm3, perm)


def test_dof_map_getters_match_the_oracle(gpu_api, port):
    """ParthAPI::old_to_new_map / added_dofs / removed_dofs (ParthAPI.h:79-84) after computeOldToNewDOFMap"""
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 12, 4)))
    M = 12 ** 3
    rng = np.random.default_rng(5)
    keep = np.sort(rng.choice(M, M - 40, replace=False)).astype(np.int32)
    n2o = np.concatenate([keep, -np.ones(25, np.int32)]).astype(np.int32)
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 4; g.reorder_type = 1
    g.import_tree(tree, Mp, Mi)
    Mn = len(n2o)
    # any graph over Mn vertices will do: a path
    P = np.zeros(Mn + 1, np.int32); deg = np.full(Mn, 2, np.int32); deg[0] = deg[-1] = 1
    P[1:] = np.cumsum(deg)
    I = np.empty(P[-1], np.int32)
    for v in range(Mn):
        nb = [u for u in (v - 1, v + 1) if 0 <= u < Mn]
        I[P[v]:P[v + 1]] = nb
    g.setMesh(Mn, P, I, n2o)
    o2n, added, removed = g.dof_maps()
    want = -np.ones(M, np.int32); want[keep] = np.arange(len(keep), dtype=np.int32)
    assert np.array_equal(o2n, want)
    assert added.tolist() == list(range(len(keep), Mn))
    assert removed.tolist() == sorted(set(range(M)) - set(keep.tolist()))


@pytest.mark.parametrize("variant", ["aggr", "nolag", "relocate"])
def test_sparse_relocation_equals_the_dense_formulation(gpu_api, variant):
    """stage_maps.cu has two formulations of relocateDOFsForAggressiveReuse (Integrator.cpp:377-544): passes over all
    M DOFs, and a sparse one over the moved DOFs and the nodes they touch (taken for <= 8192 changed edges).  They must
    leave the same tree, update after update (PARTH_B200_DENSE_RELOCATE forces the dense one)."""
    import os
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 20, 5)))
    M = 20 ** 3
    hs = []
    for _ in range(3):  # dense | sparse with the cluster regroup kernel | sparse with one CTA per node
        g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = gpu_api.AMD
        g.lagging = variant == "aggr"
        g.activate_aggressive_reuse = variant == "aggr"
        g.relocate_lvl = 5
        g.setCoarsePolicy(gpu_api.COARSE_RELOCATE if variant == "relocate" else gpu_api.COARSE_REPAIR)
        g.import_tree(tree, Mp, Mi)
        hs.append(g)
    rng = np.random.default_rng(5)
    cur = (Mp, Mi)
    for step in range(6):
        k = [3, 40, 400, 1, 2000, 25][step]
        cur = synth.add_edges(cur[0], cur[1], rng.integers(0, M, size=(k, 2)))
        outs = []
        for which, g in enumerate(hs):
            os.environ.pop("PARTH_B200_DENSE_RELOCATE", None)
            os.environ.pop("PARTH_B200_REGROUP_ONE_CTA", None)
            if which == 0:
                os.environ["PARTH_B200_DENSE_RELOCATE"] = "1"
            elif which == 2:
                os.environ["PARTH_B200_REGROUP_ONE_CTA"] = "1"
            g.setMesh(M, cur[0], cur[1])
            st, perm = g.computePermutation(1)
            assert st == 0
            outs.append((perm.copy(), g.export_tree(), g.changed_edges().copy(), g.getReuse()))
        os.environ.pop("PARTH_B200_DENSE_RELOCATE", None)
        os.environ.pop("PARTH_B200_REGROUP_ONE_CTA", None)
        for other in (1, 2):
            assert np.array_equal(outs[0][0], outs[other][0])
            assert np.array_equal(outs[0][2], outs[other][2]) and outs[0][3] == outs[other][3]
            for key in ("labels", "dofs", "mesh_perm", "dof2node", "node_ptr", "meta"):
                assert np.array_equal(outs[0][1][key], outs[other][1][key]), (step, other, key)


def test_regroup_error_flags_immediate_and_deferred(gpu_api):
    """A node that receives DOFs must hold its survivors in ascending order (the reference's merge relies on it).  The
    host-pointer update reports the violation itself; the stream-ordered update returns at once and the sticky flag
    is reported by the next call that synchronises."""
    import torch
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    M = 16 ** 3
    # node 0 (top separator) with two of its DOFs swapped: still a valid tree, no longer ascending
    dofs, labels = tree.dofs.copy(), tree.labels.copy()
    a, b = int(tree.node_ptr[0]), int(tree.node_ptr[0]) + 5
    dofs[[a, b]] = dofs[[b, a]]
    labels[[a, b]] = labels[[b, a]]
    from oracle.cpu_parth import Tree
    bad = Tree(tree.M_n, tree.levels, tree.nodes, tree.meta, tree.node_ptr, dofs, labels)
    # an edge between the two halves below the root: relocation moves an end point into node 0
    left, right = int(tree.node_dofs(3)[0]), int(tree.node_dofs(5)[0])
    P, I = synth.add_edges(Mp, Mi, [(left, right)])

    def handle():
        g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.lagging = False
        g.setCoarsePolicy(gpu_api.COARSE_RELOCATE)
        g.import_tree(bad, Mp, Mi)
        return g

    g = handle()
    g.setMesh(M, P, I)
    with pytest.raises(gpu_api.ParthError) as e:
        g.computePermutation(1)
    assert e.value.code == -6  # PARTH_B200_ERR_UNSUPPORTED
    g = handle()
    dP, dI = torch.from_numpy(P).cuda(), torch.from_numpy(I).cuda()
    dperm = torch.empty(M, dtype=torch.int32, device="cuda")
    g.setMeshDev(M, dP.data_ptr(), dI.data_ptr(), len(I))
    assert g.computePermutationDev(dperm.data_ptr(), 1) == 0  # stream-ordered: the flag is still on the device
    with pytest.raises(gpu_api.ParthError) as e:
        g.synchronize()
    assert e.value.code == -6
    g.synchronize()  # reported once
