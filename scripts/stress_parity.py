"""developer stress: random update walks at sizes whose arrays go through the staged host copies (> 4 MB, power-of-two
DOF counts included) -- host-pointer and device-pointer API against the CPU oracle, every step bit-compared
(changed edges, dirty sets, reuse; the permutation whenever the oracle finishes without METIS).
usage: python scripts/stress_parity.py [n ...]   (default 64 100 128)"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import cpu_parth as port
from parth_b200 import api, synth
from tests import scenario_defs as sd

sizes = [int(a) for a in sys.argv[1:]] or [64, 100, 128]
out = {}
for n in sizes:
    L = 6
    tree, Mp, Mi = sd.tree_of(dict(tree=("geo", n, L)))
    M = len(Mp) - 1
    for mode in ("host", "dev"):
        for lagging, aggr in ((True, False), (False, True)):
            rng = np.random.default_rng(n * 10 + lagging)
            g = api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = 1; g.lagging = lagging
            g.activate_aggressive_reuse = aggr
            g.setCoarsePolicy(api.COARSE_REPORT)
            o = port.CpuParth("port"); o.set_options(reorder_type=1, nd_levels=L, lagging=lagging, aggressive=aggr)
            g.import_tree(tree, Mp, Mi); o.import_tree(tree, Mp, Mi)
            cur, extra, bad = (Mp, Mi), [], []
            for step in range(6):
                kind = int(rng.integers(0, 4))
                if kind == 0 and extra:
                    k = int(rng.integers(1, len(extra) + 1))
                    cur = synth.remove_edges(cur[0], cur[1], extra[:k]); extra = extra[k:]
                elif kind in (1, 2):
                    e = rng.integers(0, M, size=(int(rng.integers(1, 40)), 2))
                    if step % 2: e[:, 1] = np.minimum(M - 1, e[:, 0] + rng.integers(1, 3, size=len(e)))  # local: fine nodes
                    e = [(int(a), int(b)) for a, b in e if a != b]
                    cur = synth.add_edges(cur[0], cur[1], e); extra += e
                N, Ap, Ai = synth.graph_to_matrix(*cur)
                o.setMatrix(N, Ap, Ai, 1)
                if mode == "host":
                    g.setMatrix(N, Ap, Ai, 1)
                    st, perm = g.computePermutation(1)
                else:
                    dAp = torch.from_numpy(np.ascontiguousarray(Ap, np.int32)).cuda()
                    dAi = torch.from_numpy(np.ascontiguousarray(Ai, np.int32)).cuda()
                    dperm = torch.empty(N, dtype=torch.int32, device="cuda")
                    torch.cuda.synchronize()
                    g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), int(Ai.size), 1)
                    st = g.computePermutationDev(dperm.data_ptr(), 1)
                    torch.cuda.synchronize()
                    perm = dperm.cpu().numpy()
                ost, operm = o.computePermutation(1)
                ok = (g.getNumChanges() == o.getNumChanges() and np.array_equal(g.changed_edges(), o.changed_edges())
                      and g.getReuse() == o.getReuse())
                gf, gc = g.dirty(); of, oc = o.dirty()
                ok = ok and gf.tolist() == of.tolist() and gc.tolist() == oc.tolist()
                Mp2, Mi2 = g.mesh()
                ok = ok and np.array_equal(Mp2, cur[0]) and np.array_equal(Mi2, cur[1])
                if ost == 0:
                    ok = ok and st == 0 and np.array_equal(perm, operm)
                else:
                    g.import_tree(tree, *cur); o.import_tree(tree, *cur)
                if not ok: bad.append(step)
            out[f"{n}_{mode}_{'lag' if lagging else 'aggr'}"] = "ok" if not bad else f"MISMATCH at steps {bad}"
            print(n, mode, lagging, out[f"{n}_{mode}_{'lag' if lagging else 'aggr'}"], flush=True)
print(json.dumps(out))
