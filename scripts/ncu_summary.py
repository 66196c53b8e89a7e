#!/usr/bin/env python
"""scripts/ncu_summary.py -- turn ncu outputs brought back in gpurun_out/ into the markdown tables kept under profiles/.
  python scripts/ncu_summary.py launches gpurun_out/x.csv          launch list (gpu__time_duration.sum per kernel)
  python scripts/ncu_summary.py full gpurun_out/a.ncu-rep [b ...]  selected metrics of `ncu --set full` captures"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1000 if d["Metric Unit"] == "ns" else (v * 1000 if d["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total µs | avg µs | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} % |")
    print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")


def full(paths):
    cols = []
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            cols.append({k: (r[h.index(k)], u[h.index(k)]) for k in WANT + ["Kernel Name"] if k in h})
    print("| metric | unit | " + " | ".join(c["Kernel Name"][0].split("(")[0].replace("void ", "") for c in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for k in WANT:
        if any(k in c for c in cols):
            unit = next(c[k][1] for c in cols if k in c)
            cells = [(c.get(k, ("", ""))[0] + ("" if c.get(k, ("", unit))[1] == unit else " " + c[k][1])) for c in cols]
            print(f"| {k} | {unit} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
