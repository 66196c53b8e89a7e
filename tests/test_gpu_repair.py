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
def test_repair_restores_the_separator_property(gpu_api, port, k, rt):
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", 16, 5)))
    M = 16 ** 3
    rng = np.random.default_rng(k)
    P, I = synth.add_edges(Mp, Mi, rng.integers(0, M, size=(k, 2)))
    g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = 5; g.reorder_type = rt; g.lagging = False
    g.setCoarsePolicy(gpu_api.COARSE_REPAIR)
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
