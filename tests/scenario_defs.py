"""Shared warm-update scenarios: the same seeded inputs are fed to the reference
(tests/golden/make_golden.py -> committed expectations), to the oracle restatement (CPU tests) and
to the CUDA path (GPU tests).

A scenario = separator-tree fixture + previous mesh graph + new graph (as a mesh or as a CSC
matrix, optionally with a new-to-old DOF map) + options.  Results are compared as small arrays
(changed edges, dirty sets) and FNV-1a-64 hashes of the big ones (permutation, tree arrays).
"""
import os

import numpy as np

from parth_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMD, METIS = 1, 0
_cache = {}


def _geo(n, L):
    key = ("geo", n, L)
    if key not in _cache:
        from oracle.cpu_parth import Tree
        meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
        Mp, Mi = synth.grid_graph(n)
        _cache[key] = (Tree(n ** 3, L, len(meta), meta, node_ptr, dofs, labels), Mp, Mi)
    return _cache[key]


def _fish():
    if "fish" not in _cache:
        from oracle.cpu_parth import Tree
        from tests.golden import fixtures
        t = Tree.load(os.path.join(GOLD, "tree_fish.npz"))
        N, Ap, Ai = fixtures.matrix("original")
        Mp, Mi = synth.mesh_from_pattern(Ap, Ai, 1)
        _cache["fish"] = (t, Mp, Mi)
    return _cache["fish"]


def _g12():
    if "g12" not in _cache:
        from oracle.cpu_parth import Tree
        t = Tree.load(os.path.join(GOLD, "tree_g12.npz"))
        Mp, Mi = synth.grid_graph(12)
        _cache["g12"] = (t, Mp, Mi)
    return _cache["g12"]


def build_cases():
    """extra stage-(a) inputs: name -> (N, Ap, Ai, dim)"""
    N, Ap, Ai = synth.poisson3d(20)
    out = {"grid20_full": (N, Ap, Ai, 1)}
    N2, Ap2, Ai2 = synth.poisson3d(20, lower_only=True)
    out["grid20_lower"] = (N2, Ap2, Ai2, 1)
    N3, Ap3, Ai3 = synth.kron_blocks(Ap, Ai, 3)
    out["grid20_kron3"] = (N3, Ap3, Ai3, 3)
    return out


def remesh_box(n, Mp, Mi, lo, hi, rewire_seed=None):
    """delete the vertices of the box [lo,hi)^3 of an n^3 grid and re-create them with ids appended
    at the end (shape of reference src/remesher/remeshPatch.cpp:160-176).  Optionally rewire a few
    edges inside the patch so the update also carries real topology changes."""
    M = n ** 3
    idx = np.arange(M)
    z, y, x = idx % n, (idx // n) % n, idx // (n * n)
    inside = (x >= lo) & (x < hi) & (y >= lo) & (y < hi) & (z >= lo) & (z < hi)
    surv, newv = idx[~inside], idx[inside]
    new_to_old = np.concatenate([surv, -np.ones(len(newv), np.int64)]).astype(np.int32)
    relabel = np.empty(M, np.int64)
    relabel[surv] = np.arange(len(surv))
    relabel[newv] = len(surv) + np.arange(len(newv))
    src = np.repeat(np.arange(M), np.diff(Mp))
    r, c = relabel[Mi], relabel[src]
    if rewire_seed is not None:
        rng = np.random.default_rng(rewire_seed)
        k = max(2, len(newv) // 8)
        a = np.concatenate([rng.choice(newv, k), rng.integers(0, M, 6)])
        b = rng.integers(0, M, k + 6)
        r = np.concatenate([r, relabel[a], relabel[b]])
        c = np.concatenate([c, relabel[b], relabel[a]])
        keep = r != c
        r, c = r[keep], c[keep]
    _, P, I = synth.coo_to_csc(M, r, c)
    return P, I, new_to_old


def delete_dofs(Mp, Mi, victims):
    """remove DOFs (and their edges); survivors keep order"""
    M = len(Mp) - 1
    alive = np.ones(M, bool)
    alive[victims] = False
    new_id = np.cumsum(alive) - 1
    src = np.repeat(np.arange(M), np.diff(Mp))
    keep = alive[src] & alive[Mi]
    M2 = int(alive.sum())
    _, P, I = synth.coo_to_csc(M2, new_id[Mi[keep]], new_id[src[keep]])
    new_to_old = np.nonzero(alive)[0].astype(np.int32)
    return P, I, new_to_old


def scenario_list():
    S = {}
    rng = np.random.default_rng(20260101)

    def opts(**kw):
        o = dict(reorder_type=AMD, nd_levels=8, lagging=True, aggressive=False, relocate_lvl=2, lag_lvl=5)
        o.update(kw)
        return o

    # ---- behavioural KATs of SURVEY.md Appendix B.4 (48^3 grid, 8 levels)
    t48, Mp48, Mi48 = _geo(48, 8)
    for (a, b) in [(1, 3), (3, 3), (123, 124), (125, 126), (255, 256)]:
        e = [(int(t48.node_dofs(a)[0]), int(t48.node_dofs(b)[-1]))]
        P, I = synth.add_edges(Mp48, Mi48, e)
        for lag in (True, False):
            S[f"kat48_{a}_{b}_lag{int(lag)}"] = dict(tree=("geo", 48, 8), new=("mesh", P, I, None),
                                                    opts=opts(lagging=lag, nd_levels=8), dim=1)

    # ---- 16^3 grid, 5 levels
    t16, Mp16, Mi16 = _geo(16, 5)
    M16 = 16 ** 3
    S["g16_nochange"] = dict(tree=("geo", 16, 5), new=("mesh", Mp16, Mi16, None), opts=opts(nd_levels=5), dim=3)
    for k in (1, 7, 60):
        e = rng.integers(0, M16, size=(k, 2))
        P, I = synth.add_edges(Mp16, Mi16, e)
        S[f"g16_add{k}_default"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, None), opts=opts(nd_levels=5), dim=1)
        S[f"g16_add{k}_nolag"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, None),
                                      opts=opts(nd_levels=5, lagging=False), dim=1)
        S[f"g16_add{k}_aggr_all"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, None),
                                         opts=opts(nd_levels=5, lagging=False, aggressive=True, relocate_lvl=5), dim=2)
        S[f"g16_add{k}_aggr_lvl2"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, None),
                                          opts=opts(nd_levels=5, lagging=True, aggressive=True, relocate_lvl=2), dim=1)
    # removed edges only (never reported; rows still "changed")
    src = np.repeat(np.arange(M16), np.diff(Mp16))
    pick = rng.choice(len(Mi16), 25, replace=False)
    P, I = synth.remove_edges(Mp16, Mi16, np.stack([src[pick], Mi16[pick]], axis=1))
    S["g16_remove25"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, None), opts=opts(nd_levels=5, lagging=False), dim=1)
    # through setMatrix with dim = 3 blocks
    e = rng.integers(0, M16, size=(5, 2))
    P, I = synth.add_edges(Mp16, Mi16, e)
    N1, Ap1, Ai1 = synth.graph_to_matrix(P, I)
    N3, Ap3, Ai3 = synth.kron_blocks(Ap1, Ai1, 3)
    S["g16_matrix_dim3_aggr"] = dict(tree=("geo", 16, 5), new=("matrix", N3, Ap3, Ai3, 3, None),
                                     opts=opts(nd_levels=5, lagging=False, aggressive=True, relocate_lvl=5), dim=3)

    # ---- DOF-map updates on the 16^3 grid
    for (lo, hi) in [(2, 6), (0, 4), (5, 12)]:
        P, I, m = remesh_box(16, Mp16, Mi16, lo, hi)
        S[f"g16_remesh_{lo}_{hi}"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, m),
                                          opts=opts(nd_levels=5, lagging=False), dim=1)
        P, I, m = remesh_box(16, Mp16, Mi16, lo, hi, rewire_seed=lo + 1)
        S[f"g16_remesh_{lo}_{hi}_rewired_aggr"] = dict(
            tree=("geo", 16, 5), new=("mesh", P, I, m),
            opts=opts(nd_levels=5, lagging=False, aggressive=True, relocate_lvl=5), dim=1)
    victims = rng.choice(M16, 40, replace=False)
    P, I, m = delete_dofs(Mp16, Mi16, victims)
    S["g16_delete40"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, m), opts=opts(nd_levels=5, lagging=False), dim=1)
    # delete whole leaf + neighbours: exercises label compaction of several nodes
    victims = np.concatenate([t16.node_dofs(40), t16.node_dofs(9)[:5], t16.node_dofs(0)[::7]])
    P, I, m = delete_dofs(Mp16, Mi16, victims)
    S["g16_delete_leaf40"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, m), opts=opts(nd_levels=5, lagging=False), dim=1)

    # ---- METIS-built trees (fixtures): the reference's own quick_start flow (config 1)
    if os.path.exists(os.path.join(GOLD, "tree_fish.npz")) and os.path.exists(os.path.join(GOLD, "matrices.npz")):
        from tests.golden import fixtures
        for k in "01234":
            N, Ap, Ai = fixtures.matrix(k)
            S[f"fish_{k}_default"] = dict(tree=("fish",), new=("matrix", N, Ap, Ai, 1, None),
                                          opts=opts(reorder_type=METIS, nd_levels=8), dim=3)
            S[f"fish_{k}_aggr_amd"] = dict(tree=("fish",), new=("matrix", N, Ap, Ai, 1, None),
                                           opts=opts(nd_levels=8, lagging=False, aggressive=True, relocate_lvl=8), dim=1)
    if os.path.exists(os.path.join(GOLD, "tree_g12.npz")):
        t12, Mp12, Mi12 = _g12()
        e = rng.integers(0, 12 ** 3, size=(9, 2))
        P, I = synth.add_edges(Mp12, Mi12, e)
        S["g12metis_add9_aggr"] = dict(tree=("g12",), new=("mesh", P, I, None),
                                       opts=opts(nd_levels=4, lagging=False, aggressive=True, relocate_lvl=4), dim=1)
        S["g12metis_add9_nolag"] = dict(tree=("g12",), new=("mesh", P, I, None),
                                        opts=opts(nd_levels=4, lagging=False), dim=1)
    # ---- added DOFs with rows wider than a warp (placement: the one-thread walk of the pull kernel); own generator,
    # appended last so that the scenarios above keep their random draws
    rng2 = np.random.default_rng(77)
    P, I, m = remesh_box(16, Mp16, Mi16, 5, 12)
    n_surv = int((m != -1).sum())
    added_ids = np.arange(n_surv, M16)
    hub_hi, hub_lo = M16 - 1, n_surv  # waits for every added neighbour / releases every added neighbour
    e = [(hub_hi, int(x)) for x in rng2.choice(added_ids[:-1], 30, replace=False)]
    e += [(hub_hi, int(x)) for x in rng2.choice(n_surv, 15, replace=False)]
    e += [(hub_lo, int(x)) for x in rng2.choice(added_ids[1:-1], 40, replace=False)]
    P, I = synth.add_edges(P, I, e)
    S["g16_remesh_5_12_hubs"] = dict(tree=("geo", 16, 5), new=("mesh", P, I, m), opts=opts(nd_levels=5, lagging=False), dim=1)
    return S


def tree_of(sc):
    kind = sc["tree"][0]
    if kind == "geo":
        return _geo(sc["tree"][1], sc["tree"][2])
    if kind == "fish":
        return _fish()
    if kind == "g12":
        return _g12()
    raise KeyError(kind)


def _pack(status, h, perm, tree, metis_used):
    f, c = h.dirty()
    edges = h.changed_edges()
    res = {
        "status": np.int32(status), "num_changes": np.int32(h.getNumChanges()), "reuse": np.float64(h.getReuse()),
        "dirty_fine": np.asarray(f, np.int32), "dirty_coarse": np.asarray(c, np.int32),
        "edges_hash": np.array(synth.hexhash(edges.reshape(-1))), "edges_head": edges[:64].copy(),
        "deterministic": np.int32(0 if metis_used else 1),
    }
    if not metis_used and perm is not None:
        res["perm_hash"] = np.array(synth.hexhash(perm))
        res["perm_len"] = np.int32(len(perm))
        for k in ("meta", "node_ptr", "dofs", "labels", "dof2node", "mesh_perm"):
            res["tree_" + k + "_hash"] = np.array(synth.hexhash(np.asarray(tree[k]).reshape(-1)))
    return res


def run_scenario(cp, kind, sc):
    """run on a CPU arm (cp = oracle.cpu_parth module, kind = "reference" | "port")"""
    tree, Mp, Mi = tree_of(sc)
    h = cp.CpuParth(kind)
    h.set_options(**sc["opts"])
    h.import_tree(tree, Mp, Mi)
    new = sc["new"]
    if new[0] == "mesh":
        h.setMesh(len(new[1]) - 1, new[1], new[2], new[3])
    else:
        h.setMatrix(new[1], new[2], new[3], new[4], new[5])
    st, perm = h.computePermutation(sc["dim"])
    f, c = h.dirty()
    metis_used = len(c) > 0 or (h.getNumChanges() > 0 and h.getReuse() == 0.0) or st != 0
    t = h.export_tree()
    td = {k: getattr(t, k) for k in ("meta", "node_ptr", "dofs", "labels", "dof2node", "mesh_perm")}
    return _pack(st, h, perm, td, metis_used)


def run_scenario_gpu(api, sc, coarse_policy=None):
    """run on the CUDA path through the C ABI"""
    tree, Mp, Mi = tree_of(sc)
    g = api.ParthAPI()
    o = sc["opts"]
    g.reorder_type = o["reorder_type"]; g.ND_levels = o["nd_levels"]; g.lagging = o["lagging"]
    g.activate_aggressive_reuse = o["aggressive"]; g.relocate_lvl = o["relocate_lvl"]; g.lag_lvl = o["lag_lvl"]
    g.verbose = False
    g.setCoarsePolicy(api.COARSE_REPORT if coarse_policy is None else coarse_policy)
    g.import_tree(tree, Mp, Mi)
    new = sc["new"]
    if new[0] == "mesh":
        g.setMesh(len(new[1]) - 1, new[1], new[2], new[3])
    else:
        g.setMatrix(new[1], new[2], new[3], new[4], new[5])
    st, perm = g.computePermutation(sc["dim"])
    f, c = g.dirty()
    metis_used = len(c) > 0 or (g.getNumChanges() > 0 and g.getReuse() == 0.0) or st != 0
    td = g.export_tree()
    return _pack(st, g, perm, td, metis_used), g
