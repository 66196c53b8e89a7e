"""bench_configs.py -- the other BASELINE.json configs (1, 3, 4, 5) inside the default `python bench.py` run.

bench.py's headline stays config 2 (200^3, the configuration the metric is quoted on).  This module adds, under the
key "configs" of the same JSON line, one bounded measurement of every other config, each with
  * the device time of our path (CUDA events inside the library / around the calls),
  * a parity verdict of the same run (against the CPU reference where that finishes in seconds, through
    size-independent properties at 512^3 and for the Morton order),
  * the reference's CPU path timed on the same inputs where it finishes in < 60 s,
  * the roofline of the config's dominant kernel.
Synthetic inputs of configs 3 and 5 are generated ON THE DEVICE with torch (plumbing: the numpy generators of
parth_b200/synth.py need ~1-3 minutes at these sizes); tests/test_bench_configs.py checks the torch generators against
the numpy ones.  Nothing here is on the product path.
"""
import os
import time
from types import SimpleNamespace

import numpy as np

GRID_DIRS = ((1, 0, 0), (0, 1, 0), (0, 0, 1))
KUHN_DIRS = ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1))


# ------------------------------------------------------------------ torch generators (device side)
def lattice_pattern_dev(n, dirs, diagonal, dev, chunk=1 << 24):
    """symmetric pattern on the n^3 lattice (index z + y n + x n^2), columns ascending: the torch twin of
    synth.lattice_pattern.  Returns (Ap int32 [N+1] -- int64 when nnz >= 2^31 --, Ai int32)"""
    import torch
    N = n ** 3
    offs = []
    for (dx, dy, dz) in dirs:
        for s in (-1, 1):
            offs.append((s * (dz + dy * n + dx * n * n), s * dx, s * dy, s * dz))
    if diagonal:
        offs.append((0, 0, 0, 0))
    offs.sort()
    o_t = torch.tensor([o[0] for o in offs], dtype=torch.int32, device=dev)
    cnts, parts = [], []
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        ind = torch.arange(a, b, dtype=torch.int32, device=dev)
        z, y, x = ind % n, (ind // n) % n, ind // (n * n)
        mask = torch.ones((b - a, len(offs)), dtype=torch.bool, device=dev)
        for k, (_, dx, dy, dz) in enumerate(offs):
            for c, d in ((x, dx), (y, dy), (z, dz)):
                if d > 0:
                    mask[:, k] &= c < n - d
                elif d < 0:
                    mask[:, k] &= c >= -d
        vals = ind[:, None] + o_t[None, :]
        parts.append(vals[mask])
        cnts.append(mask.sum(1, dtype=torch.int32))
        del mask, vals
    Ai = torch.cat(parts)
    del parts
    Ap = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.cat(cnts), 0, out=Ap[1:])
    if int(Ap[-1]) < 2 ** 31:
        Ap = Ap.to(torch.int32)
    return Ap, Ai


def coo_to_csc_dev(n, rows, cols):
    """sorted-unique CSC of (row, col) pairs (torch twin of synth.coo_to_csc)"""
    import torch
    key = torch.unique(cols.to(torch.int64) * n + rows.to(torch.int64))
    c = key // n
    Ap = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    torch.cumsum(torch.bincount(c, minlength=n), 0, out=Ap[1:])
    return Ap.to(torch.int32), (key % n).to(torch.int32)


def remesh_box_dev(n, Mp, Mi, lo, hi):
    """torch twin of tests/scenario_defs.remesh_box (no rewiring): the vertices of the box [lo,hi)^3 are deleted and
    re-created with ids appended at the end (shape of reference src/remesher/remeshPatch.cpp:160-176)"""
    import torch
    dev = Mp.device
    M = n ** 3
    idx = torch.arange(M, device=dev)
    z, y, x = idx % n, (idx // n) % n, idx // (n * n)
    inside = (x >= lo) & (x < hi) & (y >= lo) & (y < hi) & (z >= lo) & (z < hi)
    surv, newv = idx[~inside], idx[inside]
    new_to_old = torch.cat([surv, torch.full((len(newv),), -1, dtype=torch.int64, device=dev)]).to(torch.int32)
    relabel = torch.empty(M, dtype=torch.int64, device=dev)
    relabel[surv] = torch.arange(len(surv), device=dev)
    relabel[newv] = len(surv) + torch.arange(len(newv), device=dev)
    src = torch.repeat_interleave(torch.arange(M, device=dev), (Mp[1:] - Mp[:-1]).to(torch.int64))
    P, I = coo_to_csc_dev(M, relabel[Mi.to(torch.int64)], relabel[src])
    return P, I, new_to_old


def graph_to_matrix_dev(Mp, Mi):
    """graph + full diagonal -> CSC pattern (torch twin of synth.graph_to_matrix)"""
    import torch
    dev = Mp.device
    M = len(Mp) - 1
    src = torch.repeat_interleave(torch.arange(M, device=dev), (Mp[1:] - Mp[:-1]).to(torch.int64))
    d = torch.arange(M, device=dev)
    return coo_to_csc_dev(M, torch.cat([Mi.to(torch.int64), d]), torch.cat([src, d]))


def kron_blocks_dev(Ap, Ai, dim):
    """pattern (x) full dim x dim blocks (torch twin of synth.kron_blocks): column c*dim+d holds rows r*dim+e for
    every r of column c and e in 0..dim-1, ascending"""
    import torch
    dev = Ap.device
    n = len(Ap) - 1
    Ap64 = Ap.to(torch.int64)
    deg = Ap64[1:] - Ap64[:-1]
    Bp = torch.zeros(n * dim + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.repeat_interleave(deg * dim, dim), 0, out=Bp[1:])
    total = int(Bp[-1])
    out = torch.empty(total, dtype=torch.int32, device=dev)
    step = 1 << 26
    for a in range(0, total, step):
        p = torch.arange(a, min(total, a + step), device=dev)
        col = torch.searchsorted(Bp, p, right=True) - 1      # expanded column c*dim+d
        c = col // dim
        l = p - Bp[col]                                       # position inside the expanded column
        out[a:a + len(p)] = (Ai[Ap64[c] + l // dim].to(torch.int64) * dim + l % dim).to(torch.int32)
    return n * dim, Bp.to(torch.int32), out


def insert_entries_dev(Ap, Ai, cols, rows):
    """torch twin of workloads.insert_entries for short columns (grids): (col,row) entries, not present yet, inserted
    into a CSC pattern with ascending columns"""
    import torch
    dev = Ai.device
    N = len(Ap) - 1
    key = torch.unique(cols.to(torch.int64) * N + rows.to(torch.int64))
    c, r = key // N, key % N
    a, b = Ap[c].to(torch.int64), Ap[c + 1].to(torch.int64)
    below = torch.zeros_like(a)
    for t in range(int((b - a).max())):
        q = a + t
        ok = q < b
        below += (ok & (Ai[torch.where(ok, q, a)].to(torch.int64) < r)).to(torch.int64)
    pos = (a + below).tolist()                 # ascending (keys ascend, columns ascend)
    r32 = r.to(torch.int32)
    pieces, prev = [], 0
    for j, p in enumerate(pos):
        if p > prev:
            pieces.append(Ai[prev:p])
        pieces.append(r32[j:j + 1])
        prev = p
    pieces.append(Ai[prev:])
    Ai2 = torch.cat(pieces)
    add = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    add.index_add_(0, c + 1, torch.ones_like(c))
    Ap2 = Ap.to(torch.int64) + torch.cumsum(add, 0)
    return (Ap2.to(torch.int32) if int(Ap2[-1]) < 2 ** 31 else Ap2), Ai2


def is_permutation_dev(p, n):
    import torch
    if p.numel() != n:
        return False
    seen = torch.zeros(n, dtype=torch.int32, device=p.device)
    ok = bool(((p >= 0) & (p < n)).all())
    if ok:
        seen.index_add_(0, p.to(torch.int64), torch.ones(n, dtype=torch.int32, device=p.device))
        ok = bool((seen == 1).all())
    return ok


# ------------------------------------------------------------------ helpers
def _tree_ns(n, L, synth):
    meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
    return SimpleNamespace(M_n=n ** 3, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)


def _cpu(kind):
    from oracle import cpu_parth
    if kind == "port":
        cpu_parth.build("port")
    return cpu_parth


def _same_update(g, o, st, perm, ost, operm, ordering_is_pinned=True):
    """every METIS-independent result of one update, ours against the CPU implementation: list of what differs.
    The permutation is compared when no node was re-ordered by a third-party orderer the oracle cannot pin
    (METIS_NodeND on dirty fine nodes under ReorderingType METIS, METIS re-dissection of dirty coarse nodes)."""
    bad = []
    if g.getNumChanges() != o.getNumChanges() or not np.array_equal(g.changed_edges(), o.changed_edges()):
        bad.append("changed_edges")
    gf, gc = g.dirty()
    of, oc = o.dirty()
    if gf.tolist() != of.tolist() or gc.tolist() != oc.tolist():
        bad.append("dirty_sets")
    if g.getReuse() != o.getReuse():
        bad.append("reuse")
    if ost == 0 and len(oc) == 0 and (ordering_is_pinned or len(of) == 0):
        if st != 0 or not np.array_equal(perm, operm):
            bad.append("permutation")
    return bad


# ------------------------------------------------------------------ config 1: the reference's own test matrices
def config1(api, kind, peak):
    """tests/matrices/{original,0..4}_matrix.mtx (10 786 DOF, dim = 1; committed as tests/golden/matrices.npz with the
    METIS tree the unmodified reference built for `original`): the quick_start flow -- computePermutation(perm, 1) and the
    default-dim form (3 * M_n output)"""
    from parth_b200 import synth
    from tests.golden import fixtures
    cpu = _cpu(kind)
    tree = fixtures.tree("tree_fish.npz")
    N, Ap0, Ai0 = fixtures.matrix("original")
    Mp0, Mi0 = synth.mesh_from_pattern(Ap0, Ai0, 1)
    mats = [fixtures.matrix(k) for k in "01234"]
    out = {"workload": "reference tests/matrices: original -> 0..4 (10 786 DOF, 75 490 nnz, dim=1), warm update on the "
                       "reference's METIS tree, reference defaults", "updates": len(mats)}
    for dim in (1, 3):
        g = api.ParthAPI()
        g.verbose = False
        o = cpu.CpuParth(kind)
        o.set_options(reorder_type=0, nd_levels=8, lagging=True)
        tg, tc, bad = [], [], []
        for rep in range(2):  # the second pass over the five matrices is the timed one
            for (n_, Ap, Ai) in mats:
                # every update starts from the tree the reference built for `original` (untimed): an update that
                # reaches METIS changes the tree in a way no other METIS build reproduces, so chained updates would
                # stop being comparable after the first dirty coarse node
                g.import_tree(tree, Mp0, Mi0)
                o.import_tree(tree, Mp0, Mi0)
                t0 = time.perf_counter()
                g.setMatrix(n_, Ap, Ai, 1)
                st, perm = g.computePermutation(dim)
                t1 = time.perf_counter()
                o.setMatrix(n_, Ap, Ai, 1)
                ost, operm = o.computePermutation(dim)
                t2 = time.perf_counter()
                bad += _same_update(g, o, st, perm, ost, operm, ordering_is_pinned=False)
                if len(perm) != dim * n_:
                    bad.append("length")
                if rep:
                    tg.append(t1 - t0)
                    tc.append(t2 - t1)
        out[f"dim{dim}"] = {"ms_per_update": 1e3 * float(np.mean(tg)), "updates_per_s": 1.0 / float(np.mean(tg)),
                            "cpu_ms_per_update": 1e3 * float(np.mean(tc)), "parity": not bad, "differs": sorted(set(bad)),
                            "last_update": {"changes": g.getNumChanges(), "reuse": g.getReuse()}}
    out["cpu_kind"] = kind
    out["note"] = "host pointers in, host permutation out (wall clock); a 10 786-DOF update is launch-latency bound"
    return out


# ------------------------------------------------------------------ config 3: tet mesh, dim = 3, remesh with a DOF map
def config3(api, kind, peak, n=126, frac=0.05):
    import torch
    from parth_b200 import synth, workloads
    cpu = _cpu(kind)
    dev = torch.device("cuda", torch.cuda.current_device())
    L, dim = 8, 3
    t0 = time.time()
    Mp_d, Mi_d = lattice_pattern_dev(n, KUHN_DIRS, False, dev)
    Nv = n ** 3
    lo = 40 * n // 126
    hi = lo + int(round((frac * Nv) ** (1 / 3)))
    P, I, n2o = remesh_box_dev(n, Mp_d, Mi_d, lo, hi)
    Apm, Aim = graph_to_matrix_dev(P, I)
    N3, Ap3, Ai3 = kron_blocks_dev(Apm, Aim, dim)
    del P, I, Apm, Aim
    torch.cuda.synchronize()
    Mp, Mi = Mp_d.cpu().numpy(), Mi_d.cpu().numpy()
    new_to_old = n2o.cpu().numpy()
    tree = _tree_ns(n, L, synth)
    gen_s = time.time() - t0
    nnzA = int(Ai3.numel())

    g = api.ParthAPI()
    g.verbose = False
    g.ND_levels = L
    g.reorder_type = api.AMD
    g.lagging = False
    g.setCoarsePolicy(api.COARSE_REPORT)
    dperm = torch.empty(N3, dtype=torch.int32, device=dev)
    reps = []
    for rep in range(3):  # the same update three times from a freshly imported tree; medians are reported
        g.import_tree(tree, Mp, Mi)
        g.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(torch.cuda.ExternalStream(g.stream(), device=dev)):
            e0.record()
            g.setMatrixDev(N3, Ap3.data_ptr(), Ai3.data_ptr(), nnzA, dim)
            sb = g.stats()
            g.setNewToOldDOFMapDev(n2o.data_ptr(), int(n2o.numel()))  # the map is device-resident like the matrix
            rc = g.computePermutationDev(dperm.data_ptr(), dim)
            e1.record()
        torch.cuda.synchronize()
        s = g.stats()
        reps.append(dict(update_ms=e0.elapsed_time(e1), build_ms=sb["build_ms"], build_kernel_ms=sb["build_kernel_ms"],
                         proof_kernel_ms=sb["proof_kernel_ms"], integrate_ms=s["integrate_ms"], detect_ms=s["detect_ms"],
                         reorder_ms=s["reorder_ms"], expand_ms=s["expand_ms"], sweeps=s["integrate_sweeps"], rc=int(rc)))
    med = {k: float(np.median([r[k] for r in reps])) for k in reps[0] if k != "rc"}
    st, perm = reps[-1]["rc"], dperm.cpu().numpy()
    gMp, gMi = g.mesh()

    # the reference's CPU path on the same inputs (timed: it is this config's cpu_baseline)
    Ap3h, Ai3h = Ap3.cpu().numpy(), Ai3.cpu().numpy()
    o = cpu.CpuParth(kind)
    o.set_options(reorder_type=1, nd_levels=L, lagging=False)
    o.import_tree(cpu.Tree(Nv, L, tree.nodes, tree.meta, tree.node_ptr, tree.dofs, tree.labels), Mp, Mi)
    t1 = time.perf_counter()
    o.setMatrix(N3, Ap3h, Ai3h, dim, new_to_old)
    t2 = time.perf_counter()
    ost, operm = o.computePermutation(dim)
    t3 = time.perf_counter()
    oMp, oMi = o.mesh()
    ok = bool(np.array_equal(gMp, oMp) and np.array_equal(gMi, oMi))
    ok &= not _same_update(g, o, st, perm, ost, operm)
    gt, ot = g.export_tree(), o.export_tree()
    for k in ("node_ptr", "dofs", "labels", "dof2node", "meta"):
        ok &= bool(np.array_equal(np.asarray(gt[k]).reshape(-1), np.asarray(getattr(ot, k)).reshape(-1)))
    M, nnzM = len(gMp) - 1, len(gMi)
    alg = workloads.build_bytes(N3, nnzA, M, nnzM, dim)
    kms = med["build_kernel_ms"] + med["proof_kernel_ms"]
    return {
        "workload": f"Kuhn tet lattice {n}^3 ({Nv} vertices, 14-neighbour graph) (x) 3x3 blocks: N = {N3}, nnz(A) = {nnzA}, "
                    f"dim = 3; all vertices of a box holding 5 % of them deleted and re-created with new ids appended "
                    f"(new_to_old map), lagging off, AMD",
        "update_ms": med["update_ms"], "updates_per_s": 1e3 / med["update_ms"], "gnnz_per_s": nnzA / med["update_ms"] / 1e6,
        "stage_ms": {k: med[k] for k in ("build_ms", "integrate_ms", "detect_ms", "reorder_ms", "expand_ms")},
        "integrate_sweeps": int(med["sweeps"]), "added_dofs": int((new_to_old == -1).sum()),
        "changes": g.getNumChanges(), "reuse": g.getReuse(), "status": int(st), "parity": bool(ok),
        "parity_against": f"oracle ({kind}): mesh graph, changed edges, dirty sets, reuse, separator-tree arrays"
                          + (", permutation" if ost == 0 else " (permutation not compared: METIS reached, status %d)" % ost),
        "roofline": {"bound": "hbm", "kernel": "build_cols_spec_kernel<3> + symmetry proof (stage a, dim 3)",
                     "algorithmic_bytes": alg, "kernel_ms": kms, "build_kernel_ms": med["build_kernel_ms"],
                     "proof_kernel_ms": med["proof_kernel_ms"], "achieved": alg / kms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg / kms / 1e6 / peak, "frac_build_kernel_alone": alg / med["build_kernel_ms"] / 1e6 / peak},
        "cpu_baseline": {"kind": kind, "cores": os.cpu_count(), "set_matrix_s": t2 - t1, "compute_permutation_s": t3 - t2,
                         "updates_per_s": 1.0 / (t3 - t1), "status": int(ost)},
        "gen_s": gen_s,
    }


# ------------------------------------------------------------------ config 4: 50 M points, Morton order + dim-3 expansion
def config4(api, kind, peak, M=50_000_000):
    import torch
    cpu = _cpu(kind)
    dev = torch.device("cuda", torch.cuda.current_device())
    dim = 3
    g = api.ParthAPI()
    mesh_perm = torch.randperm(M, device=dev, generator=torch.Generator(device=dev).manual_seed(7)).to(torch.int32)
    d_out = torch.empty(M * dim, dtype=torch.int32, device=dev)
    for _ in range(3):
        g.mapMeshPermToMatrixPermDev(mesh_perm.data_ptr(), M, d_out.data_ptr(), dim)
    ms = []
    for _ in range(10):
        g.mapMeshPermToMatrixPermDev(mesh_perm.data_ptr(), M, d_out.data_ptr(), dim)
        ms.append(g.stats()["expand_ms"])
    t = float(np.median(ms))
    # parity: against the CPU implementation on the first 4 M entries (bounded sample), by definition on all of them
    exp_ok = bool((d_out.view(M, dim) == (mesh_perm[:, None] * dim + torch.arange(dim, dtype=torch.int32, device=dev))).all())
    sample = mesh_perm[:4_000_000].cpu().numpy()
    o = cpu.CpuParth(kind)
    t0 = time.perf_counter()
    o_out = o.mapMeshPermToMatrixPerm(sample, dim)
    cpu_s = time.perf_counter() - t0
    exp_ok &= bool(np.array_equal(o_out, d_out[:4_000_000 * dim].cpu().numpy()))
    alg = 4 * M * (1 + dim)
    out = {"workload": f"{M} points uniform in [0,1)^3: MORTON_CODE ordering (extension: the reference has no implementation, "
                       f"HMD.cpp:82-83) + mapMeshPermToMatrixPerm(dim = 3)",
           "expand": {"ms": t, "parity": exp_ok,
                      "roofline": {"bound": "hbm", "kernel": "expand_vec_kernel<3>", "algorithmic_bytes": alg, "kernel_ms": t,
                                   "achieved": alg / t / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e6 / peak},
                      "cpu_baseline": {"kind": kind, "sample": "first 4 M entries", "ms": 1e3 * cpu_s,
                                       "ms_scaled_to_all": 1e3 * cpu_s * M / 4e6}}}
    del d_out
    gen = torch.Generator(device=dev).manual_seed(7)
    xyz = torch.rand((M, 3), dtype=torch.float64, device=dev, generator=gen)
    torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        g.setCoordinatesDev(xyz.data_ptr(), M)
        ms.append(g.stats()["morton_ms"])
    passes = g.stats()["morton_passes"]
    perm = torch.from_numpy(g.mortonOrder()).to(dev)
    keys = torch.from_numpy(g.mortonKeys().view(np.int64)).to(dev)
    tie = keys[1:] == keys[:-1]
    ok = is_permutation_dev(perm, M) and bool((keys[1:] >= keys[:-1]).all()) and bool((perm[1:][tie] > perm[:-1][tie]).all())
    alg = M * (24 + 36 + 8 + passes * 24)
    t = float(np.median(ms[1:]))
    out["morton"] = {"ms": t, "gpoints_per_s": M / t / 1e6, "radix_passes": int(passes),
                     "parity": "unpinned by construction (no reference output exists); checked: permutation, keys ascending, "
                               "ties in point order = " + str(bool(ok)), "valid": bool(ok),
                     "roofline": {"bound": "hbm", "kernel": "radix_scatter_kernel x passes + keys + histogram",
                                  "algorithmic_bytes": alg, "kernel_ms": t, "achieved": alg / t / 1e6, "peak": peak,
                                  "unit": "GB/s", "frac": alg / t / 1e6 / peak}}
    return out


# ------------------------------------------------------------------ config 5: 512^3, largest single-GPU mesh
def config5(api, kind, peak, n=512, steps=5):
    import torch
    from parth_b200 import synth, workloads
    dev = torch.device("cuda", torch.cuda.current_device())
    L = 8
    t0 = time.time()
    Ap, Ai = lattice_pattern_dev(n, GRID_DIRS, True, dev)
    N, nnzA = n ** 3, int(Ai.numel())
    tree = _tree_ns(n, L, synth)
    sub = synth.subtree_sizes(tree.meta, tree.node_ptr)
    first = 2 ** (L - 1) - 1
    parents = list(range(first, 2 * first + 1))
    # 1 % dirty: whole level-7 sub-trees, left to right, until their DOFs reach 1 % of N;
    # three different updates (different sub-trees), 1000 cousin-leaf edges per dirty sub-tree
    dofs_d = torch.from_numpy(tree.dofs).to(dev)
    updates, at = [], 0
    for u in range(3):
        chosen, acc = [], 0
        while acc < 0.01 * N:
            s = parents[at % len(parents)]
            at += 1
            chosen.append(s)
            acc += int(sub[s])
        gen = torch.Generator(device=dev).manual_seed(100 + u)
        cols, rows = [], []
        for s in chosen:
            la, ra = int(tree.meta[s, 0]), int(tree.meta[s, 1])
            a = dofs_d[int(tree.node_ptr[la]):int(tree.node_ptr[la + 1])]
            b = dofs_d[int(tree.node_ptr[ra]):int(tree.node_ptr[ra + 1])]
            pa = a[torch.randint(len(a), (1000,), device=dev, generator=gen)].to(torch.int64)
            pb = b[torch.randint(len(b), (1000,), device=dev, generator=gen)].to(torch.int64)
            cols += [pa, pb]
            rows += [pb, pa]
        Ap2, Ai2 = insert_entries_dev(Ap, Ai, torch.cat(cols), torch.cat(rows))
        updates.append((Ap2, Ai2, chosen, acc))
    torch.cuda.synchronize()
    gen_s = time.time() - t0

    g = api.ParthAPI()
    g.verbose = False
    g.ND_levels = L
    g.lagging = False
    g.reorder_type = api.METIS
    g.setCoarsePolicy(api.COARSE_RELOCATE)
    g.import_tree(tree)
    g.setMatrixDev(N, Ap.data_ptr(), Ai.data_ptr(), nnzA, 1)
    g.acceptSnapshot()  # the tree fixture belongs to the base matrix
    dperm = torch.empty(N, dtype=torch.int32, device=dev)
    ext = torch.cuda.ExternalStream(g.stream(), device=dev)

    def timed(mats):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
            for (a, b) in mats:
                g.setMatrixDev(N, a.data_ptr(), b.data_ptr(), int(b.numel()), 1)
                g.computePermutationDev(dperm.data_ptr(), 1)
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / len(mats)

    # no-change update (the snapshot equals the input)
    for _ in range(3):
        timed([(Ap, Ai)])
    nochange_ms = timed([(Ap, Ai)] * steps)
    kms = []
    for _ in range(3):
        g.setMatrixDev(N, Ap.data_ptr(), Ai.data_ptr(), nnzA, 1)
        kms.append(g.stats()["build_kernel_ms"])
        g.computePermutationDev(dperm.data_ptr(), 1)
    ok = g.getNumChanges() == 0 and g.getReuse() == 1.0 and is_permutation_dev(dperm, N)
    M, nnzM = g.mesh_dims()
    fused = bool(g.stats().get("build_fused_detect", 0))
    alg = workloads.build_bytes(N, nnzA, M, nnzM, 1) + (workloads.detect_bytes(M, nnzM) // 2 if fused else 0)
    k = float(np.mean(kms))
    # 1 % dirty updates: each one against the previous graph (base -> u0 -> u1 -> u2), every one re-orders
    dirty = []
    for (a, b, chosen, acc) in updates:
        ms = timed([(a, b)])
        s = g.stats()
        dirty.append({"ms": ms, "dirty_subtrees": len(chosen), "dirty_dofs": int(acc), "dirty_pct": 100.0 * acc / N,
                      "changes": g.getNumChanges(), "reuse": g.getReuse(), "dirty_coarse": g.dirty()[1].tolist(),
                      "valid_permutation": is_permutation_dev(dperm, N), "reorder_ms": s["reorder_ms"]})
    # ---- parity of the first dirty update against the CPU oracle (plain-C port, one core): everything that does not
    # depend on the third-party partitioner -- the mesh graph, the changed-edge list, the dirty sets, the reuse ratio
    oracle = None
    try:
        cpu = _cpu("port")
        t0 = time.perf_counter()
        Mp_d, Mi_d = lattice_pattern_dev(n, GRID_DIRS, False, dev)
        Mp_h, Mi_h = Mp_d.cpu().numpy(), Mi_d.cpu().numpy()
        del Mp_d, Mi_d
        Ap2, Ai2 = updates[0][0], updates[0][1]
        Ap2h, Ai2h = Ap2.cpu().numpy(), Ai2.cpu().numpy()
        o = cpu.CpuParth("port")
        o.set_options(reorder_type=0, nd_levels=L, lagging=False)
        o.import_tree(cpu.Tree(N, L, tree.nodes, tree.meta, tree.node_ptr, tree.dofs, tree.labels), Mp_h, Mi_h)
        t1 = time.perf_counter()
        o.setMatrix(N, Ap2h, Ai2h, 1)
        t2 = time.perf_counter()
        ost, _ = o.computePermutation(1)
        t3 = time.perf_counter()
        g2 = api.ParthAPI()
        g2.verbose = False
        g2.ND_levels = L
        g2.lagging = False
        g2.reorder_type = api.METIS
        g2.setCoarsePolicy(api.COARSE_RELOCATE)
        g2.import_tree(tree)
        g2.setMatrixDev(N, Ap.data_ptr(), Ai.data_ptr(), nnzA, 1)
        g2.acceptSnapshot()
        g2.setMatrixDev(N, Ap2.data_ptr(), Ai2.data_ptr(), int(Ai2.numel()), 1)
        gMp, gMi = g2.mesh()
        oMp, oMi = o.mesh()
        same_mesh = bool(np.array_equal(gMp, oMp) and np.array_equal(gMi, oMi))
        del gMp, gMi, oMp, oMi
        g2.computePermutationDev(dperm.data_ptr(), 1)
        gf, gc = g2.dirty()
        of, oc = o.dirty()
        oracle = {"kind": "port", "cores": 1, "mesh_graph_equal": same_mesh,
                  "changed_edges_equal": bool(g2.getNumChanges() == o.getNumChanges()
                                              and np.array_equal(g2.changed_edges(), o.changed_edges())),
                  "dirty_sets_equal": bool(gf.tolist() == of.tolist() and gc.tolist() == oc.tolist()),
                  "reuse_equal": bool(g2.getReuse() == o.getReuse()), "oracle_status": int(ost),
                  "set_matrix_s": t2 - t1, "compute_permutation_s": t3 - t2, "prepare_s": t1 - t0,
                  "note": "the oracle stops at the partitioner (dirty coarse nodes): the permutation itself is checked by "
                          "its properties above and bit for bit at 200^3"}
        oracle["parity"] = bool(oracle["mesh_graph_equal"] and oracle["changed_edges_equal"] and oracle["dirty_sets_equal"]
                                and oracle["reuse_equal"])
        del g2, o
    except Exception as e:  # noqa: BLE001 (host memory of the box is the usual reason)
        oracle = {"error": repr(e)}
    return {
        "oracle_first_dirty_update": oracle,
        "workload": f"poisson3d_{n}^3 dim=1 ({N / 1e6:.1f}M DOF, {nnzA / 1e6:.1f}M nnz): largest single-GPU mesh; no-change update "
                    f"and 1 % dirty (whole level-7 sub-trees, 1000 cousin-leaf edges each, lagging off, RELOCATE policy)",
        "nochange": {"ms_per_update": nochange_ms, "updates_per_s": 1e3 / nochange_ms, "gnnz_per_s": nnzA / nochange_ms / 1e6,
                     "steps": steps, "parity": "size-independent properties: 0 changes, reuse 1, valid permutation = " + str(bool(ok)),
                     "valid": bool(ok)},
        "dirty_1pct": {"ms_per_update": float(np.mean([d["ms"] for d in dirty[1:]])), "first_update_ms": dirty[0]["ms"],
                       "note": "the first re-ordering update of a handle grows its scratch buffers (memory-pool growth); "
                               "ms_per_update is the mean of the updates after it",
                       "updates": dirty,
                       "parity": "mesh graph, changed edges, dirty sets and reuse ratio of the first update equal the CPU oracle "
                                 "(oracle_first_dirty_update); valid permutation after every update; the relocation path is "
                                 "bit-compared with the oracle at 200^3 (tests/test_gpu_fullsize.py)"},
        "roofline": {"bound": "hbm", "kernel": "build_pipe_kernel" + (" (stage a + fused stage b)" if fused else ""),
                     "algorithmic_bytes": alg, "kernel_ms": k, "achieved": alg / k / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg / k / 1e6 / peak},
        "cpu_baseline": ({"kind": "port", "cores": 1, "updates_per_s": 1.0 / (oracle["set_matrix_s"] + oracle["compute_permutation_s"]),
                          "sample": "the first 1.5 % dirty update (setMatrix + computePermutation up to the partitioner)"}
                         if oracle and "set_matrix_s" in oracle else None),
        "cpu_baseline_note": "the plain-C port on one core (the compiled reference needs ~2 minutes per update at 512^3 and "
                             "tens of GB for its tuple sort: not run inside the default bench)",
        "gen_s": gen_s,
    }


def run_all(api, kind, peak, which=("1", "3", "4", "5")):
    """every config in its own try block: a failure is reported in the line, never takes the headline down"""
    import torch
    out = {}
    fns = {"1": config1, "3": config3, "4": config4, "5": config5}
    for k in which:
        t0 = time.time()
        try:
            out["config" + k] = fns[k](api, kind, peak)
        except Exception as e:  # noqa: BLE001
            out["config" + k] = {"error": repr(e)}
        out["config" + k]["wall_s"] = time.time() - t0
        torch.cuda.empty_cache()
    return out
