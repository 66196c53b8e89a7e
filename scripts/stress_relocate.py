"""developer stress: the sparse relocation (side stream, cluster regroup, device-side mirrors) against the dense
formulation at 1 M DOF / 8 levels, update after update: same permutation, same tree, separator property intact
(stage (d) takes its tree-aware path, which it only does when no edge violates the dissection), nnz(L) equal.
usage: python scripts/stress_relocate.py [n] [steps]"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parth_b200 import api, synth
from tests import scenario_defs as sd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = 8
tree, Mp, Mi = sd.tree_of(dict(tree=("geo", n, L)))
M = n ** 3
out = {"n": n, "levels": L, "steps": []}
for policy, pname in ((api.COARSE_RELOCATE, "relocate"), (api.COARSE_REPAIR, "repair")):
    hs = []
    for _ in range(2):
        g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = api.AMD; g.lagging = False
        g.setCoarsePolicy(policy)
        g.import_tree(tree, Mp, Mi)
        hs.append(g)
    rng = np.random.default_rng(11)
    cur = (Mp, Mi)
    for step in range(steps):
        k = [3, 60, 700, 1, 3000, 25, 1000, 1000, 5, 200][step % 10]
        a = rng.integers(0, M, size=k)
        far = rng.integers(0, M, size=k)
        near = np.minimum(M - 1, a + rng.integers(1, 2 * n, size=k))
        b = np.where(rng.random(k) < 0.7, near, far)  # mostly local edges: leaves and low separators get dirty
        cur = synth.add_edges(cur[0], cur[1], np.stack([a, b], 1))
        res = []
        for which, g in enumerate(hs):
            os.environ.pop("PARTH_B200_DENSE_RELOCATE", None)
            if which == 0:
                os.environ["PARTH_B200_DENSE_RELOCATE"] = "1"
            g.setMesh(M, cur[0], cur[1])
            st, perm = g.computePermutation(1)
            t = g.export_tree()
            _, _, lnz = g.symbolic(None)
            res.append((st, perm.copy(), t, int(lnz), g.stats()["symbolic_path"], g.getReuse(), g.getNumChanges()))
        os.environ.pop("PARTH_B200_DENSE_RELOCATE", None)
        same = res[0][0] == res[1][0] == 0 and np.array_equal(res[0][1], res[1][1]) and res[0][3] == res[1][3] and \
            all(np.array_equal(res[0][2][key], res[1][2][key]) for key in ("labels", "dofs", "mesh_perm", "dof2node", "node_ptr", "meta"))
        valid = bool(np.array_equal(np.sort(res[1][1]), np.arange(M)))
        rec = dict(policy=pname, step=step, edges=k, equal_to_dense=bool(same), valid_perm=valid,
                   tree_aware_symbolic=[r[4] for r in res], lnz=res[1][3], reuse=res[1][5], changes=res[1][6])
        out["steps"].append(rec)
        print(rec, flush=True)
out["all_ok"] = all(r["equal_to_dense"] and r["valid_perm"] and r["tree_aware_symbolic"] == [2, 2] for r in out["steps"])
print(json.dumps({"all_ok": out["all_ok"], "n": n, "updates": len(out["steps"])}))
