"""Development aid (not bench.py): per-stage device times on the resident 200^3 workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from parth_b200 import api, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
what = sys.argv[2] if len(sys.argv) > 2 else "all"
st = workloads.PoissonUpdateStream(n=n, levels=8)
g = api.ParthAPI(); g.verbose = False; g.ND_levels = 8
g.import_tree(st.tree(), st.Mp, st.Mi)
mats = []
for k in range(4):
    Ap, Ai = st.matrix(k)
    mats.append((torch.from_numpy(Ap).cuda(), torch.from_numpy(Ai).cuda(), len(Ai)))
dperm = torch.empty(st.N, dtype=torch.int32, device="cuda")
for assume in (False, True):
    g.setAssumeSymmetric(assume)
    for lag in (True, False):
        g.lagging = lag
        for it in range(6):
            Ap, Ai, nnz = mats[it % 4]
            g.setMatrixDev(st.N, Ap.data_ptr(), Ai.data_ptr(), nnz, 1)
            s1 = g.stats()
            t0 = time.perf_counter()
            g.computePermutationDev(dperm.data_ptr(), 1)
            t1 = time.perf_counter()
            s2 = g.stats()
        alg = workloads.build_bytes(st.N, nnz, *g.mesh_dims(), 1)
        print(f"assume={int(assume)} lag={int(lag)} build={s1['build_ms']:.3f} kernel={s1['build_kernel_ms']:.3f} "
              f"({alg/s1['build_kernel_ms']/1e6:.0f} GB/s, {alg/s1['build_kernel_ms']/1e6/6545.6:.3f}) path={s1['build_path']} "
              f"detect={s2['detect_ms']:.3f} dirty={s2['dirty_ms']:.3f} reorder={s2['reorder_ms']:.3f} expand={s2['expand_ms']:.3f} "
              f"compute_wall={1e3*(t1-t0):.3f} changes={g.getNumChanges()} reuse={g.getReuse():.5f} dirty={g.dirty()[1].tolist()}")
if what == "all":
    for fresh in (False, True):
        if fresh:  # pristine tree + graph: the separator property holds everywhere
            g.import_tree(st.tree(), st.Mp, st.Mi); g.setMesh(st.N, st.Mp, st.Mi)
        t0 = time.perf_counter()
        parent, cc, lnz = g.symbolic(None)
        t1 = time.perf_counter()
        s = g.stats()
        print(f"symbolic (fresh={fresh}): nnz(L)={lnz} device={s['symbolic_ms']:.1f} ms etree={s['symbolic_etree_ms']:.1f} ms "
              f"path={s['symbolic_path']} wall={1e3*(t1-t0):.1f} ms")
