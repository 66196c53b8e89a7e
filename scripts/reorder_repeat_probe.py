"""developer probe: the same many-dirty-nodes update repeated in one process (is a slow ordering phase a first-use cost?)"""
import os, sys, json, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parth_b200 import api, synth, workloads
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import config_suite as cs
n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 8
N, Ap, Ai = synth.lattice_pattern(n, synth.GRID_DIRS)
Mp, Mi = cs.graph_of(Ap, Ai)
tree = cs.tree_ns(n, L)
first = 2 ** (L - 1) - 1
rng = np.random.default_rng(3)
cols, rws = [], []
for s in range(first, first + 68):
    a = tree.dofs[tree.node_ptr[tree.meta[s, 0]]:tree.node_ptr[tree.meta[s, 0] + 1]]
    b = tree.dofs[tree.node_ptr[tree.meta[s, 1]]:tree.node_ptr[tree.meta[s, 1] + 1]]
    e = np.unique(np.stack([rng.choice(a, 1000), rng.choice(b, 1000)], axis=1), axis=0)
    cols += [e[:, 0], e[:, 1]]; rws += [e[:, 1], e[:, 0]]
Ap2, Ai2 = workloads.insert_entries(Ap, Ai, np.concatenate(cols), np.concatenate(rws))
g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.lagging = False; g.reorder_type = api.AMD
g.setCoarsePolicy(api.COARSE_REPAIR)
dperm = torch.empty(N, dtype=torch.int32, device="cuda")
dAp, dAi = torch.from_numpy(Ap).cuda(), torch.from_numpy(Ai).cuda()
dAp2, dAi2 = torch.from_numpy(Ap2).cuda(), torch.from_numpy(Ai2).cuda()
perms = []
for rep in range(6):
    g.import_tree(tree, Mp, Mi)
    g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)
    g.computePermutationDev(dperm.data_ptr(), 1)
    torch.cuda.synchronize()
    g.setMatrixDev(N, dAp2.data_ptr(), dAi2.data_ptr(), len(Ai2), 1)
    t0 = time.perf_counter()
    g.computePermutationDev(dperm.data_ptr(), 1)
    torch.cuda.synchronize()
    print("rep", rep, "computePermutation wall ms", round(1e3 * (time.perf_counter() - t0), 2), "reorder_ms", round(g.stats()["reorder_ms"], 2), flush=True)
    perms.append(dperm.cpu().numpy().copy())
print("all identical:", all(np.array_equal(perms[0], p) for p in perms[1:]))
