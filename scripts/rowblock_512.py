"""torchrun entry: config 5 (512^3 grid, 134 M DOF, 938 M nnz) as ONE mesh over the ranks -- row blocks of the graph
build + change detection, the library's own NCCL communicator.  Every rank generates ONLY its block of the
matrix on its device, uploads nothing else, and the no-change update is timed with CUDA events (max over ranks).
Checks: 0 changes, reuse 1, the returned permutation is a permutation and equal on every rank.
  torchrun --nproc-per-node N scripts/rowblock_512.py [n] [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch, torch.distributed as dist
import bench_configs as bc
from parth_b200 import api, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L, N = 8, n ** 3
t0 = time.time()
meta, node_ptr, dofs, labels = synth.geometric_tree(n, L)
tree = SimpleNamespace(M_n=N, levels=L, nodes=len(meta), meta=meta, node_ptr=node_ptr, dofs=dofs, labels=labels)

g = api.ParthAPI(local)
g.verbose = False; g.ND_levels = L; g.lagging = False; g.reorder_type = api.METIS
g.setCoarsePolicy(api.COARSE_RELOCATE)
uid = [api.ParthAPI.commUniqueId() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
g.commInit(uid[0], rank, world)
g.import_tree(tree)

# column counts of the 7-point pattern in closed form -> global Ap (every rank, 0.5 GB), Ai of this rank's block only
ind = torch.arange(N, dtype=torch.int32, device=dev)
z, y, x = ind % n, (ind // n) % n, ind // (n * n)
cnt = (1 + (z > 0).int() + (z < n - 1).int() + (y > 0).int() + (y < n - 1).int() + (x > 0).int() + (x < n - 1).int())
Ap = torch.zeros(N + 1, dtype=torch.int64, device=dev)
torch.cumsum(cnt, 0, out=Ap[1:])
nnz = int(Ap[-1])
del ind, z, y, x, cnt
c0, c1 = g.blockRange(N, nnz)
q0, q1 = int(Ap[c0]), int(Ap[c1])
offs = torch.tensor([-n * n, -n, -1, 0, 1, n, n * n], dtype=torch.int32, device=dev)
cols = torch.arange(c0, c1, dtype=torch.int32, device=dev)
zz, yy, xx = cols % n, (cols // n) % n, cols // (n * n)
mask = torch.stack([xx > 0, yy > 0, zz > 0, torch.ones_like(zz, dtype=torch.bool), zz < n - 1, yy < n - 1, xx < n - 1], 1)
pad = q0 & 3
dAi = torch.zeros(pad + (q1 - q0) + 8, dtype=torch.int32, device=dev)
dAi[pad:pad + q1 - q0] = (cols[:, None] + offs[None, :])[mask]
dAp = Ap[c0:c1 + 1].to(torch.int32).contiguous()
del Ap, cols, zz, yy, xx, mask
torch.cuda.synchronize()
gen_s = time.time() - t0
pAp, pAi = dAp.data_ptr() - 4 * c0, dAi.data_ptr() + 4 * pad - 4 * q0

g.setMatrixBlockDev(N, nnz, pAp, pAi, q0, q1, 1)
g.acceptSnapshot()  # the tree fixture belongs to this matrix
dperm = torch.empty(N, dtype=torch.int32, device=dev)


def update():
    g.setMatrixBlockDev(N, nnz, pAp, pAi, q0, q1, 1)
    g.computePermutationDev(dperm.data_ptr(), 1)


for _ in range(3):
    update()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(torch.cuda.ExternalStream(g.stream(), device=dev)):
    e0.record()
    for _ in range(steps):
        update()
    e1.record()
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
kms = []
for _ in range(3):
    g.setMatrixBlockDev(N, nnz, pAp, pAi, q0, q1, 1)
    kms.append(g.stats()["build_kernel_ms"])
    g.computePermutationDev(dperm.data_ptr(), 1)
ok = g.getNumChanges() == 0 and g.getReuse() == 1.0 and bc.is_permutation_dev(dperm, N)
h = torch.tensor([int(dperm.to(torch.int64).mul_(torch.arange(1, N + 1, device=dev) % 1000003).sum() % (2 ** 61))], device=dev)
hs = [torch.zeros_like(h) for _ in range(world)]
dist.all_gather(hs, h)
ok &= all(int(v) == int(hs[0]) for v in hs)
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
cols_, qa = c1 - c0, q1 - q0
alg = 4 * ((cols_ + 1) + qa + 2 * ((cols_ + 1) + (qa - cols_)))
if rank == 0:
    k = float(np.mean(kms))
    print(json.dumps({"config": 5, "n": n, "N": N, "nnzA": nnz, "world": world, "ok": bool(flag.item()),
                      "nochange_ms_per_update": float(t.item()), "updates_per_s": 1e3 / float(t.item()),
                      "gnnz_per_s": nnz / float(t.item()) / 1e6, "steps": steps,
                      "rank0_block": {"columns": [c0, c1], "nnz": qa, "build_kernel_ms": k, "algorithmic_bytes": alg,
                                      "GBs": alg / k / 1e6, "frac_of_6545.6": alg / k / 1e6 / 6545.6},
                      "comm": g.commInfo()[2], "gen_s": gen_s}), flush=True)
del g
dist.destroy_process_group()
