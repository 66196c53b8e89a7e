#!/usr/bin/env python
"""scripts/morton_probe.py [n] -- ReorderingType MORTON_CODE on n points uniform in [0,1)^3 (default 50 M, BASELINE
config 4): device time of set_coordinates (box + keys + radix sort + rank), validity of the order.  Run under ncu for
the launch list / full captures kept in profiles/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parth_b200 import api  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6545.6) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6545.6


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    xyz = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=gen)
    torch.cuda.synchronize()
    g = api.ParthAPI()
    ms = []
    for _ in range(reps):
        g.setCoordinatesDev(xyz.data_ptr(), n)
        ms.append(g.stats()["morton_ms"])
    passes = g.stats()["morton_passes"]
    perm = torch.from_numpy(g.mortonOrder()).cuda()
    keys = torch.from_numpy(g.mortonKeys().view(np.int64)).cuda()
    seen = torch.zeros(n, dtype=torch.int32, device="cuda")
    seen.index_add_(0, perm.long(), torch.ones(n, dtype=torch.int32, device="cuda"))
    tie = keys[1:] == keys[:-1]
    ok = bool((seen == 1).all()) and bool((keys[1:] >= keys[:-1]).all()) and bool((perm[1:][tie] > perm[:-1][tie]).all())
    # algorithmic bytes: box 24n, keys 24n + 12n, histogram 8n, scatter 24n per pass
    alg = n * (24 + 36 + 8 + passes * 24)
    t = float(np.median(ms[1:])) if len(ms) > 1 else ms[0]
    print(json.dumps({"stage": "morton_order", "ok": ok, "n": n, "bits": 21, "radix_passes": passes, "morton_ms": t,
                      "all_ms": ms, "Mpoints_per_s": n / t / 1e3, "algorithmic_bytes": alg, "GBs": alg / t / 1e6,
                      "frac": alg / t / 1e6 / PEAK}), flush=True)


if __name__ == "__main__":
    main()
