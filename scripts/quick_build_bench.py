"""Scratch micro-benchmark of stage (a) on a resident 200^3 matrix (development aid, not bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from parth_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
assume = len(sys.argv) > 2 and sys.argv[2] == "assume"
N, Ap, Ai = synth.poisson3d(n)
dAp = torch.from_numpy(Ap).cuda(); dAi = torch.from_numpy(Ai).cuda()
g = api.ParthAPI()
g.setAssumeSymmetric(assume)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(6):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.setMatrixDev(N, dAp.data_ptr(), dAi.data_ptr(), len(Ai), 1)
    t1 = time.perf_counter()
    st = g.stats()
    M, nnzM = g.mesh_dims()
    alg = 4 * ((N + 1) + len(Ai) + (M + 1) + nnzM)
    print(f"it{it} path={st['build_path']} wall={1e3*(t1-t0):.3f}ms stage={st['build_ms']:.3f}ms kernel={st['build_kernel_ms']:.3f}ms "
          f"alg={alg/1e6:.1f}MB  kernel_GBs={alg/st['build_kernel_ms']/1e6:.0f} frac={alg/st['build_kernel_ms']/1e6/6545.6:.3f}")
