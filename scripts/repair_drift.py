#!/usr/bin/env python
"""scripts/repair_drift.py -- does the fill of the device's answers to dirty COARSE nodes drift over many updates?
(reference: Integrator::rePartitionCoarseHMDNode, Integrator.cpp:812-849.)

K consecutive updates with lagging off on an n^3 grid: update k ADDS 1000 cousin-leaf edges under a level-(L-1) node
to everything inserted so far (the matrix keeps growing long-range couplings, the tree is only ever repaired).  After
each update nnz(L) of the permutation is computed exactly (stage d) for the two device policies
    REPAIR    relocation into the common separator + AMD of the separators that received DOFs
    RELOCATE  the same relocation, no re-ordering (joiners at the end of their separator)
and, every `every` updates, for a FRESH ordering of the same matrix:
    fresh_device     cold start of a new handle (built-in BFS level-structure dissection + AMD of every node)
    fresh_reference  cold computePermutation of the unmodified reference (oracle/_ref: METIS dissection + AMD)
Prints one JSON line per sampled update and a summary line.   usage: repair_drift.py [n] [K] [every] [levels]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    every = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    st = workloads.PoissonUpdateStream(n=n, levels=L)
    N = st.N
    hs = {}
    for name, pol in (("REPAIR", api.COARSE_REPAIR), ("RELOCATE", api.COARSE_RELOCATE)):
        g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = api.AMD; g.lagging = False
        g.setCoarsePolicy(pol)
        g.import_tree(st.tree(), st.Mp, st.Mi)
        g.setMesh(N, st.Mp, st.Mi)
        nodes = [x for x in range(st.nodes) if st.node_ptr[x + 1] > st.node_ptr[x]]
        g.stage_permute_fine(nodes)  # every node ordered by AMD: the common starting point
        g.acceptSnapshot()
        hs[name] = g
    have_ref = cpu_parth.available("reference")
    Ap, Ai = st.Ap, st.Ai
    rows = []
    ms = {k: [] for k in hs}
    for k in range(K):
        _, e = st.edges_of(k)
        cols = np.concatenate([e[:, 0], e[:, 1]]); rws = np.concatenate([e[:, 1], e[:, 0]])
        # keep only entries that are not present yet
        keep = np.array([r not in Ai[Ap[c]:Ap[c + 1]] for c, r in zip(cols.tolist(), rws.tolist())])
        pairs = np.unique(np.stack([cols[keep], rws[keep]], 1), axis=0)
        Ap, Ai = workloads.insert_entries(Ap, Ai, pairs[:, 0], pairs[:, 1])
        out = {"update": k + 1, "nnzA": int(Ap[-1])}
        for name, g in hs.items():
            g.setMatrix(N, Ap, Ai, 1)
            t = time.perf_counter()
            rc, perm = g.computePermutation(1)
            ms[name].append(1e3 * (time.perf_counter() - t))
            assert rc == 0
            out["lnz_" + name] = int(g.symbolic(perm)[2])
            out["reuse_" + name] = g.getReuse()
        if (k + 1) % every == 0 or k == 0:
            f = api.ParthAPI(); f.verbose = False; f.ND_levels = L; f.reorder_type = api.AMD
            f.useBuiltinDecomposer(True)
            f.setMatrix(N, Ap, Ai, 1)
            rc, perm = f.computePermutation(1)
            out["lnz_fresh_device"] = int(f.symbolic(perm)[2])
            del f
            if have_ref:
                o = cpu_parth.CpuParth("reference")
                o.set_options(reorder_type=1, nd_levels=L, lagging=False)
                o.setMatrix(N, Ap, Ai, 1)
                t = time.perf_counter()
                orc, operm = o.computePermutation(1)
                out["fresh_reference_s"] = round(time.perf_counter() - t, 2)
                out["lnz_fresh_reference"] = int(hs["REPAIR"].symbolic(operm)[2])
            print(json.dumps(out), flush=True)
        rows.append(out)
    last = rows[-1]
    summ = {"summary": True, "n": n, "N": N, "levels": L, "updates": K,
            "lnz_first": {k: rows[0]["lnz_" + k] for k in hs}, "lnz_last": {k: last["lnz_" + k] for k in hs},
            "relocate_over_repair_last": last["lnz_RELOCATE"] / last["lnz_REPAIR"],
            "ms_per_update_wall": {k: float(np.median(v)) for k, v in ms.items()}}
    for key in ("lnz_fresh_device", "lnz_fresh_reference"):
        if key in last:
            summ[key + "_last"] = last[key]
            summ["repair_over_" + key[4:] + "_last"] = last["lnz_REPAIR"] / last[key]
            summ["relocate_over_" + key[4:] + "_last"] = last["lnz_RELOCATE"] / last[key]
    print(json.dumps(summ), flush=True)


if __name__ == "__main__":
    main()
