"""Digest an .ncu-rep: headline raw metrics + top source lines by instructions / stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        ]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:95s} {rows[1][i]:>16s} " + " ".join(r[i][:48] for r in rows[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h = r; continue
    if h and len(r) > 5 and r[0] != "":
        d = dict(zip(h, r))
        try: ie = int(d["Instructions Executed"]); smp = int(d["# Samples"])
        except Exception: continue
        out.append((ie, smp, cur, r[0], r[1][:100]))
ti = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("\n-- top lines by executed instructions (total %d warp-inst, %d samples)" % (ti, ts))
for o in sorted(out, reverse=True)[:topn]:
    print(f"{o[0]/ti*100:5.1f}% inst {o[1]/ts*100:5.1f}% smp {o[2]}:{o[3]}  {o[4]}")
print("\n-- top lines by stall samples")
for o in sorted(out, key=lambda x: -x[1])[:topn]:
    print(f"{o[0]/ti*100:5.1f}% inst {o[1]/ts*100:5.1f}% smp {o[2]}:{o[3]}  {o[4]}")
