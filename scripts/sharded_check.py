"""torchrun entry: sharded reuse updates (REPAIR policy, several dirty nodes per update) on every
rank vs the same updates on a private single-GPU handle; everything must be identical."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from parth_b200 import api, sharded, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, L = 24, 6
meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
from types import SimpleNamespace
tree = SimpleNamespace(M_n=n ** 3, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)
Mp, Mi = synth.grid_graph(n)
M = n ** 3
sh = sharded.ShardedParth(local)
ref = api.ParthAPI(local)
for g in (sh.g, ref):
    g.verbose = False; g.ND_levels = L; g.lagging = False; g.reorder_type = api.AMD
    g.setCoarsePolicy(api.COARSE_REPAIR)
    g.import_tree(tree, Mp, Mi)
rng = np.random.default_rng(3)  # same stream on every rank
cur = (Mp, Mi)
ok, total_bytes = True, 0
for step in range(4):
    cur = synth.add_edges(cur[0], cur[1], rng.integers(0, M, size=(40, 2)))
    N, Ap, Ai = synth.graph_to_matrix(*cur)
    sh.g.setMatrix(N, Ap, Ai, 1); ref.setMatrix(N, Ap, Ai, 1)
    rc, perm = sh.computePermutation(1)
    rc2, perm2 = ref.computePermutation(1)
    nodes, owner, sizes = sh.g.exchange_plan()
    ok &= rc == 0 and rc2 == 0 and bool(np.array_equal(perm, perm2))
    ok &= bool(np.array_equal(owner, sharded.plan_lpt(nodes, sizes, world))) and len(nodes) > 1
    a, b = sh.g.export_tree(), ref.export_tree()
    for k in ("labels", "dofs", "mesh_perm", "dof2node", "node_ptr", "meta"):
        ok &= bool(np.array_equal(a[k], b[k]))
    ok &= sorted(perm.tolist()) == list(range(M))
    total_bytes += sh.exchanged_bytes
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), "world": world, "exchanged_bytes": int(total_bytes)}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
