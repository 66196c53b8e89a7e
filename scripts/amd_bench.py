"""Development aid: device time of the batched AMD kernel on a few graph shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parth_b200 import api, synth
g = api.ParthAPI()
def box(nx, ny, nz):
    idx = np.arange(nx * ny * nz); z = idx % nz; y = (idx // nz) % ny; x = idx // (nz * ny)
    r, c = [], []
    for d, m in ((1, z < nz - 1), (nz, y < ny - 1), (nz * ny, x < nx - 1)):
        r += [idx[m], idx[m] + d]; c += [idx[m] + d, idx[m]]
    return synth.coo_to_csc(nx * ny * nz, np.concatenate(r), np.concatenate(c))[1:]
for name, gr, reps in (("plane25x25", box(1, 25, 25), 1), ("grid16^3", box(16, 16, 16), 1), ("leaf25x25x50", box(25, 25, 50), 1),
                       ("148 x leaf25x25x50", box(25, 25, 50), 148)):
    t0 = time.perf_counter(); p = g.amd_batch([gr] * reps); t1 = time.perf_counter()
    n = len(gr[0]) - 1
    print(f"{name}: n={n} nnz={len(gr[1])} graphs={reps} wall={1e3*(t1-t0):.1f} ms  -> {1e6*(t1-t0)/n:.2f} us/pivot/graph", flush=True)
