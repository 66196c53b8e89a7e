#!/usr/bin/env python
"""scripts/amd_bench.py -- stage (c) ordering micro-benchmark on one B200: the batched AMD kernel on
leaf-like (3D grid), separator-like (2D grid) and tet-like (Kuhn) graphs, checked against the CPU
restatement (oracle/amd_restated.c) and timed beside it.  Prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_parth  # noqa: E402  (checker + CPU timing only)
from parth_b200 import api, synth  # noqa: E402


def csr(n, r, c):
    _, Mp, Mi = synth.coo_to_csc(n, r, c)
    return Mp, Mi


def grid2d(k):
    idx = np.arange(k * k).reshape(k, k)
    r, c = [], []
    for a, b in ((idx[:-1, :], idx[1:, :]), (idx[:, :-1], idx[:, 1:])):
        r += [a.ravel(), b.ravel()]
        c += [b.ravel(), a.ravel()]
    return csr(k * k, np.concatenate(r), np.concatenate(c))


def kuhn(k):
    idx = np.arange(k ** 3).reshape(k, k, k)
    r, c = [], []
    for d in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]:
        a = idx[:k - d[0], :k - d[1], :k - d[2]]
        b = idx[d[0]:, d[1]:, d[2]:]
        r += [a.ravel(), b.ravel()]
        c += [b.ravel(), a.ravel()]
    return csr(k ** 3, np.concatenate(r), np.concatenate(c))


def main():
    cases = [("grid2d_40", grid2d(40), 1), ("grid2d_128", grid2d(128), 1), ("grid3d_16", synth.grid_graph(16), 1),
             ("grid3d_31", synth.grid_graph(31), 1), ("kuhn_25", kuhn(25), 1), ("grid3d_31 x64", synth.grid_graph(31), 64),
             ("grid2d_40 x148", grid2d(40), 148), ("grid2d_128 x66", grid2d(128), 66), ("grid2d_128 x148", grid2d(128), 148)]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0].split()[0] in sys.argv[1:]]
    g = api.ParthAPI()
    for name, (Mp, Mi), copies in cases:
        t0 = time.perf_counter()
        want = cpu_parth.amd(Mp, Mi)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        graphs = [(Mp, Mi)] * copies
        g.amd_batch(graphs[:1])
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            got = g.amd_batch(graphs)
            best = min(best, 1e3 * (time.perf_counter() - t0))
        ok = all(np.array_equal(p, want) for p in got)
        print(json.dumps({"case": name, "n": len(Mp) - 1, "nnz": int(len(Mi)), "graphs": copies, "bit_exact": bool(ok),
                          "gpu_ms_wall": round(best, 3), "cpu_ms_one_graph": round(cpu_ms, 3),
                          "gpu_dofs_per_s": round(copies * (len(Mp) - 1) / (best * 1e-3))}), flush=True)


if __name__ == "__main__":
    main()
