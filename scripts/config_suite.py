"""Full-size runs of BASELINE.json configs 3, 4 and 5 on the GPU box (development / evidence
script; bench.py stays on config 2).  Every config checks its result -- against the CPU oracle
where that finishes in seconds, through size-independent properties otherwise -- and prints one
JSON line with device stage times.

  python scripts/config_suite.py 3            tet-elasticity mesh, dim = 3, 5 % remesh with DOF map
  python scripts/config_suite.py 4            mapMeshPermToMatrixPerm, 50 M points, dim = 3
  python scripts/config_suite.py 5 [n]        dirty-fraction sweep on the n^3 grid (default 512)
Under torchrun, config 5 shards the dirty nodes across the ranks (parth_b200/sharded.py).
"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parth_b200 import api, synth, workloads  # noqa: E402

PEAK = 6545.6


def tree_ns(n, L):
    meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
    return SimpleNamespace(M_n=n ** 3, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)


def graph_of(Ap, Ai):
    """drop the diagonal of a full symmetric pattern -> mesh graph (numpy, generator side)"""
    col = np.repeat(np.arange(len(Ap) - 1, dtype=np.int64), np.diff(Ap))
    keep = Ai != col
    Mp = np.zeros(len(Ap), np.int64)
    np.cumsum(np.bincount(col[keep], minlength=len(Ap) - 1), out=Mp[1:])
    return Mp.astype(np.int32), Ai[keep]


def config3():
    from oracle import cpu_parth
    from tests import scenario_defs as sd
    cpu_parth.build("port")
    n, L, dim = 126, 8, 3
    t0 = time.time()
    Nv, Apv, Aiv = synth.lattice_pattern(n, synth.KUHN_DIRS)
    Mp, Mi = graph_of(Apv, Aiv)
    tree = tree_ns(n, L)
    lo = 40
    hi = lo + int(round((0.05 * Nv) ** (1 / 3)))
    P, I, new_to_old = sd.remesh_box(n, Mp, Mi, lo, hi)
    Nm, Apm, Aim = synth.graph_to_matrix(P, I)
    N3, Ap3, Ai3 = synth.kron_blocks(Apm, Aim, dim)
    gen_s = time.time() - t0
    g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = api.AMD; g.lagging = False
    g.setCoarsePolicy(api.COARSE_REPORT)
    dAp, dAi = torch.from_numpy(Ap3).cuda(), torch.from_numpy(Ai3).cuda()
    for rep in range(3):  # same update three times from a freshly imported tree; the last one is reported
        g.import_tree(tree, Mp, Mi)
        g.setMatrixDev(N3, dAp.data_ptr(), dAi.data_ptr(), len(Ai3), dim)
        sb = g.stats()
        g.setNewToOldDOFMap(new_to_old)
        gMp, gMi = g.mesh()
        st, perm = g.computePermutation(dim)
        s = g.stats()
    # oracle on the same inputs (the port's build takes a while at 2.7e8 nnz; it is the checker)
    t1 = time.time()
    o = cpu_parth.CpuParth("port"); o.set_options(reorder_type=1, nd_levels=L, lagging=False)
    o.import_tree(cpu_parth.Tree(Nv, L, tree.nodes, tree.meta, tree.node_ptr, tree.dofs, tree.labels), Mp, Mi)
    o.setMatrix(N3, Ap3, Ai3, dim, new_to_old)
    oMp, oMi = o.mesh()
    ost, operm = o.computePermutation(dim)
    cpu_s = time.time() - t1
    ok = bool(np.array_equal(gMp, oMp) and np.array_equal(gMi, oMi))
    ok &= g.getNumChanges() == o.getNumChanges() and bool(np.array_equal(g.changed_edges(), o.changed_edges()))
    gf, gc = g.dirty(); of, oc = o.dirty()
    ok &= gf.tolist() == of.tolist() and gc.tolist() == oc.tolist() and g.getReuse() == o.getReuse()
    gt, ot = g.export_tree(), o.export_tree()
    for k in ("node_ptr", "dofs", "labels", "dof2node", "meta"):
        ok &= bool(np.array_equal(np.asarray(gt[k]).reshape(-1), np.asarray(getattr(ot, k)).reshape(-1)))
    if ost == 0:
        ok &= st == 0 and bool(np.array_equal(perm, operm))
    alg = workloads.build_bytes(N3, len(Ai3), len(gMp) - 1, len(gMi), dim)
    print(json.dumps({"config": 3, "ok": ok, "N": N3, "nnzA": int(len(Ai3)), "M": int(len(gMp) - 1), "nnzM": int(len(gMi)),
                      "removed_added": int((new_to_old == -1).sum()), "build_ms": sb["build_ms"],
                      "build_kernel_ms": sb["build_kernel_ms"], "proof_kernel_ms": sb["proof_kernel_ms"], "build_path": sb["build_path"],
                      "build_GBs": alg / (sb["build_kernel_ms"] + sb["proof_kernel_ms"]) / 1e6,
                      "build_frac": alg / (sb["build_kernel_ms"] + sb["proof_kernel_ms"]) / 1e6 / PEAK,
                      "integrate_ms": s["integrate_ms"], "integrate_passes": s["integrate_passes"], "integrate_sweeps": s["integrate_sweeps"], "detect_ms": s["detect_ms"], "expand_ms": s["expand_ms"],
                      "changes": g.getNumChanges(), "reuse": g.getReuse(), "status": int(st), "oracle_status": int(ost),
                      "cpu_port_s": cpu_s, "gen_s": gen_s}), flush=True)


def config4():
    M, dim = 50_000_000, 3
    rng = np.random.default_rng(7)
    mesh_perm = rng.permutation(M).astype(np.int32)
    g = api.ParthAPI()
    d_in = torch.from_numpy(mesh_perm).cuda()
    d_out = torch.empty(M * dim, dtype=torch.int32, device="cuda")
    ext = torch.cuda.ExternalStream(g.stream())
    for _ in range(3):
        g.mapMeshPermToMatrixPermDev(d_in.data_ptr(), M, d_out.data_ptr(), dim)
    ms = []
    for _ in range(10):
        g.mapMeshPermToMatrixPermDev(d_in.data_ptr(), M, d_out.data_ptr(), dim)
        ms.append(g.stats()["expand_ms"])
    out = d_out.cpu().numpy()
    exp = (mesh_perm[:, None].astype(np.int64) * dim + np.arange(dim)).reshape(-1)
    ok = bool(np.array_equal(out, exp))
    alg = 4 * M * (1 + dim)
    t = float(np.median(ms))
    print(json.dumps({"config": 4, "ok": ok, "M": M, "dim": dim, "expand_ms": t, "GBs": alg / t / 1e6,
                      "frac": alg / t / 1e6 / PEAK, "note": "MORTON_CODE has no reference implementation "
                      "(HMD.cpp:82-83); only the dim expansion of config 4 has an oracle"}), flush=True)
    # the Morton ordering itself (extension, parity against oracle/morton_oracle.py only): 50 M points uniform in
    # [0,1)^3 (SURVEY 8d asks mt19937_64(seed 7); the device generator is used here, the distribution is the same)
    del d_out
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    xyz = torch.rand((M, 3), dtype=torch.float64, device="cuda", generator=gen)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        g.setCoordinatesDev(xyz.data_ptr(), M)
        ms.append(g.stats()["morton_ms"])
    passes = g.stats()["morton_passes"]
    perm = torch.from_numpy(g.mortonOrder()).cuda()
    keys = torch.from_numpy(g.mortonKeys().view(np.int64)).cuda()
    seen = torch.zeros(M, dtype=torch.int32, device="cuda")
    seen.index_add_(0, perm.long(), torch.ones(M, dtype=torch.int32, device="cuda"))
    tie = keys[1:] == keys[:-1]
    ok = bool((seen == 1).all()) and bool((keys[1:] >= keys[:-1]).all()) and bool((perm[1:][tie] > perm[:-1][tie]).all())
    # bytes: box 24n, keys 24n + 12n, histogram 8n, scatter 24n per pass
    alg = M * (24 + 36 + 8 + passes * 24)
    t = float(np.median(ms[1:]))
    print(json.dumps({"config": 4, "stage": "morton_order", "ok": ok, "M": M, "bits": 21, "radix_passes": passes,
                      "morton_ms": t, "all_ms": ms, "Mpoints_per_s": M / t / 1e3, "GBs": alg / t / 1e6,
                      "frac": alg / t / 1e6 / PEAK}), flush=True)


def config5(n):
    import torch.distributed as dist
    from parth_b200 import sharded
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = 8
    t0 = time.time()
    N, Ap, Ai = synth.lattice_pattern(n, synth.GRID_DIRS)
    Mp, Mi = graph_of(Ap, Ai)
    tree = tree_ns(n, L)
    sub = synth.subtree_sizes(tree.meta, tree.node_ptr)
    gen_s = time.time() - t0
    if world > 1:
        sh = sharded.ShardedParth(local)
        g = sh.g
    else:
        sh, g = None, api.ParthAPI(local)
    g.verbose = False; g.ND_levels = L; g.lagging = False; g.reorder_type = api.AMD
    g.setCoarsePolicy(api.COARSE_REPAIR)
    g.import_tree(tree, Mp, Mi)
    dperm = torch.empty(N, dtype=torch.int32, device="cuda")
    dAp, dAi = torch.from_numpy(Ap).cuda(), torch.from_numpy(Ai).cuda()
    first = 2 ** (L - 1) - 1
    parents = list(range(first, 2 * first + 1))
    rows = []
    for pct in (0.0, 0.1, 0.3, 1, 3, 10, 30, 50):
        # dirty whole level-(L-1) sub-trees, left to right, until their DOFs reach pct % of N
        chosen, acc = [], 0
        for s in parents:
            if acc >= pct / 100 * N:
                break
            chosen.append(s); acc += int(sub[s])
        rng = np.random.default_rng(int(pct * 10) + 1)
        cols, rws = [], []
        for s in chosen:
            a = tree.dofs[tree.node_ptr[tree.meta[s, 0]]:tree.node_ptr[tree.meta[s, 0] + 1]]
            b = tree.dofs[tree.node_ptr[tree.meta[s, 1]]:tree.node_ptr[tree.meta[s, 1] + 1]]
            e = np.unique(np.stack([rng.choice(a, 1000), rng.choice(b, 1000)], axis=1), axis=0)
            cols += [e[:, 0], e[:, 1]]; rws += [e[:, 1], e[:, 0]]
        if chosen:
            Ap2, Ai2 = workloads.insert_entries(Ap, Ai, np.concatenate(cols), np.concatenate(rws))
        else:
            Ap2, Ai2 = Ap, Ai
        dAp2, dAi2 = torch.from_numpy(Ap2).cuda(), torch.from_numpy(Ai2).cuda()
        # every sweep point starts from the pristine tree and snapshot
        g.import_tree(tree, Mp, Mi)
        g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)   # proven snapshot
        g.computePermutationDev(dperm.data_ptr(), 1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        g.setMatrixDev(N, dAp2.data_ptr(), dAi2.data_ptr(), len(Ai2), 1)
        sb = g.stats()
        rc = (sh.computePermutationDev(dperm.data_ptr(), 1) if sh else g.computePermutationDev(dperm.data_ptr(), 1))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t1
        s2 = g.stats()
        perm = dperm.cpu().numpy()
        ok = rc == 0 and bool(np.array_equal(np.sort(perm), np.arange(N, dtype=np.int32)))
        alg = workloads.build_bytes(N, len(Ai2), *g.mesh_dims(), 1)
        rows.append({"dirty_pct": pct, "dirty_nodes": len(chosen), "dirty_dofs": acc, "ok": ok, "changes": g.getNumChanges(),
                     "reuse": g.getReuse(), "update_wall_ms": 1e3 * wall, "build_ms": sb["build_ms"],
                     "build_kernel_ms": sb["build_kernel_ms"], "build_frac": alg / sb["build_kernel_ms"] / 1e6 / PEAK,
                     "detect_ms": s2["detect_ms"], "reorder_ms": s2["reorder_ms"], "expand_ms": s2["expand_ms"],
                     "exchanged_bytes": (sh.exchanged_bytes if sh else 0)})
        del dAp2, dAi2
    if rank == 0:
        print(json.dumps({"config": 5, "n": n, "N": N, "nnzA": int(len(Ai)), "world": world, "gen_s": gen_s,
                          "sweep": rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "4"
    if which == "3":
        config3()
    elif which == "4":
        config4()
    else:
        config5(int(sys.argv[2]) if len(sys.argv) > 2 else 512)
