"""torchrun entry: ONE mesh over the ranks (row blocks, the library's own NCCL communicator) vs the same updates on a
private single-GPU handle holding the whole graph; permutation, changed edges, dirty sets, reuse and the separator tree
must be identical on every rank.  Usage: torchrun --nproc-per-node N scripts/rowblock_check.py [n] [levels] [variant]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch, torch.distributed as dist
from parth_b200 import api, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 6
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["nolag", "aggr", "lag"]
meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
tree = SimpleNamespace(M_n=n ** 3, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)
Mp, Mi = synth.grid_graph(n)
M = n ** 3
ok, detail = True, []
for variant in variants:
    sh, ref = api.ParthAPI(local), api.ParthAPI(local)
    for g in (sh, ref):
        g.verbose = False; g.ND_levels = L; g.reorder_type = api.AMD
        g.setCoarsePolicy(api.COARSE_REPAIR)
        g.lagging = variant != "nolag"
        g.activate_aggressive_reuse = variant == "aggr"
        if variant == "aggr":
            g.relocate_lvl = L
    uid = [api.ParthAPI.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sh.commInit(uid[0], rank, world)
    assert sh.commInfo() == (rank, world, "nccl")
    sh.import_tree(tree)
    ref.import_tree(tree, Mp, Mi)
    N0, Ap0, Ai0 = synth.graph_to_matrix(Mp, Mi)
    sh.setMatrix(N0, Ap0, Ai0, 1)
    sh.acceptSnapshot()
    rng = np.random.default_rng(7)  # the same stream on every rank
    cur = (Mp, Mi)
    for step in range(5):
        cur = (Mp, Mi) if step == 3 else synth.add_edges(cur[0], cur[1], rng.integers(0, M, size=(60, 2)))
        N, Ap, Ai = synth.graph_to_matrix(*cur)
        sh.setMatrix(N, Ap, Ai, 1); ref.setMatrix(N, Ap, Ai, 1)
        h2d = sh.stats()["h2d_bytes"]
        rc, perm = sh.computePermutation(1)
        rc2, perm2 = ref.computePermutation(1)
        same = rc == rc2 and bool(np.array_equal(perm, perm2)) and bool(np.array_equal(sh.changed_edges(), ref.changed_edges()))
        same &= all(bool(np.array_equal(a, b)) for a, b in zip(sh.dirty(), ref.dirty())) and sh.getReuse() == ref.getReuse()
        a, b = sh.export_tree(), ref.export_tree()
        same &= all(bool(np.array_equal(a[k], b[k])) for k in ("labels", "dofs", "mesh_perm", "dof2node", "node_ptr", "meta"))
        same &= sorted(perm.tolist()) == list(range(M))
        same &= world == 1 or h2d < 4 * (N + 1 + int(Ap[N])) * (1.0 / world + 0.15)
        ok &= same
        detail.append((variant, step, bool(same), int(ref.getNumChanges()), float(ref.getReuse())))
    del sh, ref
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), "world": world, "n": n, "levels": L, "rank0": detail}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
