#!/usr/bin/env python
"""scripts/repair_quality.py -- fill quality of the REPAIR policy.

The reference answers a dirty COARSE node by re-dissecting its sub-tree with METIS
(Integrator::rePartitionCoarseHMDNode, Integrator.cpp:812-849); the device relocates one endpoint of every broken
edge into the common separator and re-orders the touched nodes with AMD.  The two orderings
differ by construction, so they are compared by what an ordering is for: nnz(L) of P*A*P' (stage d, exact).

Both sides start from the same separator tree with every node ordered by AMD, receive the same update (1000
cousin-leaf edges under one level-(L-1) node, lagging off), and their permutations go through the same symbolic
analysis.  The reference runs from oracle/_ref (the unmodified reference compiled with the METIS bundled with CUDA).
Prints one JSON line per grid size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_parth  # noqa: E402
from parth_b200 import api, workloads  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [64, 100]
    kind = "reference" if cpu_parth.available("reference") else None
    for n in sizes:
        L = 6
        st = workloads.PoissonUpdateStream(n=n, levels=L)
        N = st.N
        g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = api.AMD; g.lagging = False
        g.setCoarsePolicy(api.COARSE_REPAIR)
        g.import_tree(st.tree(), st.Mp, st.Mi)
        g.setMesh(N, st.Mp, st.Mi)
        nodes = [x for x in range(st.nodes) if st.node_ptr[x + 1] > st.node_ptr[x]]
        g.stage_permute_fine(nodes)                     # every node ordered by AMD: the common starting point
        t0 = g.export_tree()
        tree0 = cpu_parth.Tree(N, L, st.nodes, t0["meta"], t0["node_ptr"], t0["dofs"], t0["labels"])
        base_perm = g.mesh_perm().copy()
        Ap, Ai = st.matrix(0)
        Mp1, Mi1 = st.mesh(0)
        # ours: REPAIR
        g.setMatrix(N, Ap, Ai, 1)
        t = time.perf_counter()
        rc, perm = g.computePermutation(1)
        ours_ms = 1e3 * (time.perf_counter() - t)
        _, _, lnz_ours = g.symbolic(perm)
        _, _, lnz_stale = g.symbolic(base_perm)        # the old ordering applied to the new matrix
        out = {"n": n, "N": N, "levels": L, "dirty_coarse": g.dirty()[1].tolist(), "reuse": g.getReuse(),
               "lnz_before_update_ordering": int(lnz_stale), "lnz_repair": int(lnz_ours), "repair_update_ms": round(ours_ms, 2)}
        if kind:
            o = cpu_parth.CpuParth(kind)
            o.set_options(reorder_type=1, nd_levels=L, lagging=False)
            o.import_tree(tree0, st.Mp, st.Mi)
            o.setMatrix(N, Ap, Ai, 1)
            t = time.perf_counter()
            orc, operm = o.computePermutation(1)
            ref_ms = 1e3 * (time.perf_counter() - t)
            _, _, lnz_ref = g.symbolic(operm)
            out.update(lnz_reference_metis_redissect=int(lnz_ref), reference_status=int(orc), reference_update_ms=round(ref_ms, 1),
                       repair_over_reference=round(lnz_ours / lnz_ref, 4))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
