"""developer probe: dim = 3 graph build kernel time with and without the full mirror proof"""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parth_b200 import api, synth, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dim = 3
Nv, Apv, Aiv = synth.lattice_pattern(n, synth.KUHN_DIRS)
N3, Ap3, Ai3 = synth.kron_blocks(Apv, Aiv, dim)
dAp, dAi = torch.from_numpy(Ap3).cuda(), torch.from_numpy(Ai3).cuda()
for assume in (False, True):
    g = api.ParthAPI(); g.verbose = False
    g.setAssumeSymmetric(assume)
    ms = []
    for rep in range(4):
        g.clearParth() if rep else None
        g.setMatrixDev(N3, dAp.data_ptr(), dAi.data_ptr(), len(Ai3), dim)
        s = g.stats(); ms.append(s["build_kernel_ms"] + s["proof_kernel_ms"]); proof = s["proof_kernel_ms"]
    M, nnzM = g.mesh_dims()
    alg = workloads.build_bytes(N3, len(Ai3), M, nnzM, dim)
    print(json.dumps({"n": n, "assume_symmetric": assume, "build_plus_proof_kernel_ms": ms, "proof_kernel_ms": proof, "frac": alg / min(ms) / 1e6 / 6545.6, "path": s["build_path"]}))
