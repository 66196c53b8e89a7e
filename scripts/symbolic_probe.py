"""developer probe: stage (d) symbolic analysis at n^3 -- device time (etree part / total) beside the CPU oracle"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_parth
from parth_b200 import api, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
check = "--check" in sys.argv
st = workloads.PoissonUpdateStream(n=n, levels=8)
g = api.ParthAPI(); g.verbose = False; g.ND_levels = 8
g.import_tree(st.tree(), st.Mp, st.Mi)
g.setMesh(st.N, st.Mp, st.Mi)
rc, perm = g.computePermutation(1)
out = {"n": n, "N": st.N}
for rep in range(3):
    t0 = time.perf_counter()
    parent, cc, lnz = g.symbolic(None)
    wall = time.perf_counter() - t0
    s = g.stats()
    out[f"gpu_rep{rep}"] = {"wall_ms": 1e3 * wall, "symbolic_ms": s["symbolic_ms"], "etree_ms": s["symbolic_etree_ms"], "path": s["symbolic_path"]}
out["lnz"] = int(lnz)
if check:
    t0 = time.perf_counter()
    op = cpu_parth.etree(st.Mp, st.Mi, perm)
    t1 = time.perf_counter()
    olnz, occ = cpu_parth.colcount(st.Mp, st.Mi, perm, op)
    t2 = time.perf_counter()
    out["cpu_etree_s"] = t1 - t0; out["cpu_colcount_s"] = t2 - t1
    out["mismatch"] = {"parent": int((op != parent).sum()), "colcount": int((occ != cc).sum()), "lnz": [int(olnz), int(lnz)],
                       "first_parent": int(np.flatnonzero(op != parent)[0]) if (op != parent).any() else -1,
                       "first_cc": int(np.flatnonzero(occ != cc)[0]) if (occ != cc).any() else -1}
    bad = np.flatnonzero(op != parent)[:8]
    out["mismatch"]["cols"] = [[int(j), int(op[j]), int(parent[j])] for j in bad]
    out["bit_exact"] = bool(np.array_equal(op, parent) and np.array_equal(occ, cc) and int(olnz) == int(lnz))
print(json.dumps(out))
