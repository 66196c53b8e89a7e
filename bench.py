#!/usr/bin/env python
"""bench.py -- reuse-update throughput of the B200 path (BASELINE.json metric) on config 2.

One step = one warm reuse update = setMatrix (CSC -> mesh graph build) + computePermutation
(change detection against the snapshot, dirty sets, re-ordering of the dirty nodes, assembly,
dim expansion) on the 200^3 7-point Poisson matrix (8 M DOF, 55.76 M nnz, dim = 1) with ~1 % of
the DOFs dirtied per update (workloads.PoissonUpdateStream; SURVEY.md section 8d, variant 2a).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code (oracle/_ref)

`value` : updates/s with the inputs resident in HBM (device-pointer entry points)
`e2e`   : updates/s through the host-pointer C ABI (pinned host CSC in, host permutation out)
`roofline`: dominant kernel (build_pipe_kernel) algorithmic bytes / its CUDA-event time
`cpu_baseline`: the reference CPU implementation timed on this box's host cores (rank 0, N = 1)
`nnz_L*`: the metric's second half -- nnz(L) of our permutation from the device symbolic stage, from the CPU oracle,
and of the reference's permutation (nnz_l_block)
`configs`: configs 1, 3, 4 and 5 of BASELINE.json, one bounded measurement each with its own parity verdict, roofline
and CPU reference time (bench_configs.py; N = 1)
N > 1: ONE mesh over the N GPUs: row blocks of the graph build and of change detection, one NCCL
all-gather of {flags, changed edges, row diffs} per update, dirty sub-meshes by all-reduce -- total work is fixed
("scaling": "strong"); every rank uploads 1/N of the matrix and returns the whole permutation.  The N independent
meshes of round 1 (no collective, weak scaling) are reported beside it under "replicas".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one build_pipe_kernel launch from `ncu --set full` (profiles/), keyed by
# (grid edge, fused stage b)
TRAFFIC = {(200, False): 440306944, (200, True): 671709184}

METRIC = "reuse-update perms/s at 8M DOF 1% change"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=200, help="grid edge (200 -> 8 M DOF)")
    ap.add_argument("--variant", default="nolag", choices=["lag", "nolag", "nolag_amd", "aggr"],
                    help="nolag (headline, SURVEY 8d variant 2a): lagging off, the dirty level-7 sub-tree is repaired "
                         "(device: RELOCATE policy = broken edges relocated into the common separator, joiners at its end; "
                         "reference: METIS re-dissection); nolag_amd: the same with the REPAIR policy (relocation + AMD "
                         "of the receiving separators: stage c runs, same fill, 5x the time -- profiles/r07_repair_drift.md); "
                         "aggr: aggressive reuse on (relocation into the separator + AMD of the touched nodes, "
                         "bit-exact with the reference); lag: reference defaults (the dirty node is lag-dropped, reuse 1)")
    ap.add_argument("--no-variants", action="store_true", help="skip the secondary variants (reported under 'variants')")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="1,3,4,5",
                    help="the other BASELINE.json configs measured under the key 'configs' (N = 1 only; bench_configs.py)")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-nnzl", action="store_true", help="skip the nnz(L) comparison of config 2")
    ap.add_argument("--replicas-only", action="store_true",
                    help="N > 1: only the N independent meshes of round 1 (no collective)")
    return ap.parse_args()


VARIANT_TEXT = {
    "lag": "lagging=true (reference defaults: the dirty level-7 node is lag-dropped, reuse 1)",
    "aggr": "reference defaults + aggressive reuse (relocate level 8), ReorderingType AMD: broken edges are relocated "
            "into the common separator and the touched nodes are re-ordered with AMD",
    "nolag": "lagging=false: the dirty coarse node is repaired on the device (RELOCATE policy: one endpoint of every "
             "broken edge moves into the common separator, which keeps its order and takes the joiners at its end) where "
             "the reference re-dissects it with METIS",
    "nolag_amd": "lagging=false, REPAIR policy: the relocation of nolag + AMD re-ordering of the separators that received "
                 "DOFs (stage c runs; fill within 0.1-0.5 % of RELOCATE, profiles/r07_repair_drift.md)",
}


def workload_name(n, variant):
    return (f"poisson3d_{n}^3 dim=1 ({n**3/1e6:.1f}M DOF), warm reuse update, 1000 cousin-leaf edges under one "
            f"level-7 node per update (~0.8% of DOFs dirty), 8 ND levels, geometric tree fixture, " + VARIANT_TEXT[variant])


def apply_variant(h, variant, api=None):
    """the option set of a variant on either implementation (our ParthAPI mirror or oracle.CpuParth)"""
    if api is not None:  # our arm
        h.lagging = not variant.startswith("nolag")
        h.activate_aggressive_reuse = variant == "aggr"
        h.relocate_lvl = 8 if variant == "aggr" else 2
        h.reorder_type = api.AMD if variant == "aggr" else api.METIS
        h.setCoarsePolicy(api.COARSE_RELOCATE if variant == "nolag" else api.COARSE_REPAIR)
    else:
        h.set_options(reorder_type=1 if variant == "aggr" else 0, nd_levels=8, lagging=not variant.startswith("nolag"),
                      aggressive=variant == "aggr", relocate_lvl=8 if variant == "aggr" else 2)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        # index: one GPU, or a comma-separated list (rank 0 samples every GPU of the job in ONE nvidia-smi
        # process: N polling processes measurably slow the launch path of a sub-millisecond step)
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 7:
                    continue
                try:
                    sm.append(float(t[0])); mx.append(float(t[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------ reference arm (CPU)
def cpu_update_stream(kind, stream, variant, steps, warmup, first=0):
    """times `steps` updates (after `warmup`) of a CPU implementation on the stream's matrices"""
    from oracle import cpu_parth
    h = cpu_parth.CpuParth(kind)
    apply_variant(h, variant)
    h.import_tree(stream.tree(), stream.Mp, stream.Mi)
    times, info, first_info = [], {}, None
    for k in range(warmup + steps):
        Ap, Ai = stream.matrix(first + k)
        t0 = time.perf_counter()
        h.setMatrix(stream.N, Ap, Ai, 1)
        st, perm = h.computePermutation(1)
        t1 = time.perf_counter()
        if k >= warmup:
            times.append(t1 - t0)
        info = dict(status=int(st), reuse=h.getReuse(), changes=h.getNumChanges())
        if k == 0:  # METIS-independent results of the first update from the fixture tree (compared with ours)
            f, c = h.dirty()
            first_info = dict(changes=h.getNumChanges(), reuse=h.getReuse(), dirty_fine=f.tolist(), dirty_coarse=c.tolist(),
                              edges=h.changed_edges().copy())
    info["_perm"] = perm
    info["_first"] = first_info
    return times, info


def cpu_kind():
    from oracle import cpu_parth
    return "reference" if cpu_parth.available("reference") else "port"


def config_dict(n, variant, world):
    """identical on both arms (the driver compares the two config objects)"""
    return {"workload": workload_name(n, variant), "variant": variant,
            "l2": "inputs larger than L2: 701 MB touched by the build + detect pass per step, a different matrix every step",
            "parallelism": f"{world} GPU(s)" if world == 1 else PARALLELISM_TEXT.format(world=world)}


PARALLELISM_TEXT = ("one mesh over {world} GPUs: row blocks of build + change detection, 1 NCCL all-gather of "
                    "{{flags, changed edges, row diffs}} per update, dirty sub-meshes by all-reduce")


def run_reference(args):
    """the reference's own CPU implementation (oracle/_ref, else the port) on the SAME config: every step is one
    full update of the named 200^3 workload (no shrinking, no extrapolation)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_parth
    from parth_b200 import workloads
    kind = cpu_kind()
    if kind == "port":
        cpu_parth.build("port")
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    stream = workloads.PoissonUpdateStream(n=args.n, levels=8)
    nnz = int(stream.Ap[-1])
    times, info = cpu_update_stream(kind, stream, args.variant, args.steps, args.warmup)
    info.pop("_perm", None)
    info.pop("_first", None)
    per = float(np.mean(times))
    value = 1.0 / per
    sample = f"{args.steps} full updates (setMatrix + computePermutation) at {args.n}^3 after {args.warmup} warm-up updates"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args.n, args.variant, args.gpus),
        "gnnz_per_s": value * nnz / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "last_update": info,
    }
    print(json.dumps(line), flush=True)


def nnz_l_block(make_handle, stream, variant, ref_perm, kind, ref_first=None):
    """the metric's second half ("nnz(L) bit-exact vs ref"): the same two updates the cpu_baseline leg has just run,
    on a fresh handle; nnz(L) of OUR permutation from the device symbolic stage (stage d), the same number from the CPU
    oracle (etree as reference src/utils/etree.cpp:39-115 + column counts), and nnz(L) of the REFERENCE's permutation.
    aggr: the permutations are identical, so all three agree; nolag: the reference re-dissects the dirty sub-tree with
    METIS where the device relocates, so nnz_L_reference is a fill-quality comparison, not a parity value."""
    from oracle import cpu_parth
    cpu_parth.build("port")
    g = make_handle(variant)
    parity_first = None
    for k in range(2):
        Ap, Ai = stream.matrix(k)
        g.setMatrix(stream.N, Ap, Ai, 1)
        st, perm = g.computePermutation(1)
        if k == 0 and ref_first is not None:
            f, c = g.dirty()
            parity_first = {"changed_edges_equal": bool(g.getNumChanges() == ref_first["changes"]
                                                        and np.array_equal(g.changed_edges(), ref_first["edges"])),
                            "dirty_sets_equal": bool(f.tolist() == ref_first["dirty_fine"] and c.tolist() == ref_first["dirty_coarse"]),
                            "reuse_equal": bool(g.getReuse() == ref_first["reuse"])}
            parity_first["parity"] = all(parity_first.values())
    Mp, Mi = g.mesh()
    g.symbolic()  # first call: the stage's scratch buffers are carved out of the memory pool
    parent, cc, lnz = g.symbolic()
    sym_ms = g.stats()["symbolic_ms"]
    lnz_given = g.symbolic(perm)[2]
    sn = g.supernodes(parent, cc)  # (f4) the consumer's first step, on the device
    sn_ms = g.stats()["supernodes_ms"]
    par = cpu_parth.etree(Mp, Mi, perm)
    lnz_o, _ = cpu_parth.colcount(Mp, Mi, perm, par)
    par_r = cpu_parth.etree(Mp, Mi, ref_perm)
    lnz_r, _ = cpu_parth.colcount(Mp, Mi, ref_perm, par_r)
    return {"parity_first_update": dict(parity_first or {}, what="first update from the fixture tree, headline variant, full size: "
                                        f"changed-edge list, dirty sets and reuse ratio against the CPU path ({kind}); what follows "
                                        "(METIS re-dissection there, relocation here) is compared through nnz(L)"),
            "nnz_L": int(lnz), "nnz_L_oracle": int(lnz_o), "nnz_L_reference": int(lnz_r),
            "nnz_L_detail": {"what": "nnz(L) = sum of the column counts of P A P^T (lower triangle with diagonal) after 2 updates "
                                     "from the tree fixture; nnz_L: device stage (d) on our permutation (tree-aware path); "
                                     "nnz_L_oracle: CPU oracle on the same permutation; nnz_L_reference: CPU oracle on the "
                                     f"permutation of the reference's CPU path ({kind}) after the same 2 updates",
                             "device_equals_oracle": bool(lnz == lnz_o and lnz_given == lnz_o),
                             "permutation_equals_reference": bool(np.array_equal(perm, ref_perm)),
                             "fill_vs_reference": float(lnz) / float(lnz_r),
                             "cholmod_L_NNZ": int(2 * lnz - stream.N),
                             "device_symbolic_ms": sym_ms, "status": int(st),
                             "supernodes": {"nsuper": sn["nsuper"], "ssize": sn["ssize"], "xsize": sn["xsize"],
                                            "device_ms": sn_ms,
                                            "equals_oracle": bool(np.array_equal(sn["super"], cpu_parth.supernodes(parent, cc)["super"]))}}}


# ------------------------------------------------------------------ our arm (CUDA)
def _init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def run_ours(args, emit=True):
    """N independent meshes, one per GPU (N = 1: the single-GPU line)"""
    import torch
    import torch.distributed as dist
    from parth_b200 import api, workloads

    world, rank, local = _init_dist()
    dev = torch.device("cuda", local)

    stream = workloads.PoissonUpdateStream(n=args.n, levels=8)
    N, M = stream.N, stream.N
    # distinct matrices, cycled; consecutive ones always differ.  The variants that repair the tree (aggr, nolag)
    # must never meet a matrix twice (its broken edges would already be repaired: nothing to re-order), so they get
    # one matrix per update they run; the headline variant changes nothing in the tree and cycles through 12.
    need = args.warmup + args.steps + min(args.steps, 8) + (0 if args.no_e2e else 2 + args.steps)
    D = min(args.steps + args.warmup, 12) if args.variant == "lag" else min(need, 64)
    first = rank * 17  # each rank updates its own sequence of nodes
    host_mats = []
    for k in range(D):
        Ap, Ai = stream.matrix(first + k)
        host_mats.append((torch.from_numpy(Ap).pin_memory(), torch.from_numpy(Ai).pin_memory()))
    dev_mats = [(a.to(dev), b.to(dev)) for a, b in host_mats]
    nnzA = int(host_mats[0][0][-1])

    def make_handle(variant):
        g = api.ParthAPI(local)
        g.verbose = False
        g.ND_levels = stream.levels
        apply_variant(g, variant, api)
        g.import_tree(stream.tree(), stream.Mp, stream.Mi)
        return g

    g = make_handle(args.variant)
    dperm = torch.empty(M, dtype=torch.int32, device=dev)
    hperm = torch.empty(M, dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(g, k):
        # the update with its ctypes arguments bound once per (handle, matrix): api.preparedUpdateDev
        prepared = g.__dict__.setdefault("_prepared", {})
        run = prepared.get(k % D)
        if run is None:
            Ap, Ai = dev_mats[k % D]
            run = prepared[k % D] = g.preparedUpdateDev(N, Ap.data_ptr(), Ai.data_ptr(), nnzA, 1, dperm.data_ptr(), 1)
        run()

    def step_e2e(g, k):
        Ap, Ai = host_mats[k % D]
        g.setMatrix(N, Ap.numpy(), Ai.numpy(), 1)
        g.computePermutation(1, out=hperm.numpy())

    def timed(g, fn, k0, steps):
        """the timed region holds nothing but the API calls a user makes (the device-pointer calls are
        stream-ordered: no statistics are read, nothing synchronises just to return)"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = g.stats()["kernels_launched"]
        with torch.cuda.stream(torch.cuda.ExternalStream(g.stream(), device=dev)):
            e0.record()
            for k in range(k0, k0 + steps):
                fn(g, k)
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), g.stats()["kernels_launched"] - launches0

    def stage_times(g, k0, steps):
        """per-stage device times of `steps` further updates (CUDA events inside the library), untimed"""
        kernel_ms, stages, diff = [], [], []
        for k in range(k0, k0 + steps):
            Ap, Ai = dev_mats[k % D]
            g.setMatrixDev(N, Ap.data_ptr(), Ai.data_ptr(), nnzA, 1)
            s = g.stats()
            g.computePermutationDev(dperm.data_ptr(), 1)
            st = g.stats()
            kernel_ms.append(s["build_kernel_ms"]); diff.append(s["diff_kernel_ms"])
            stages.append((s["build_ms"], st["detect_ms"], st["dirty_ms"], st["reorder_ms"], st["expand_ms"]))
        return kernel_ms, stages, float(np.mean(diff))

    def update_info(g):
        return dict(reuse=g.getReuse(), changes=g.getNumChanges(), dirty=[x.tolist() for x in g.dirty()])

    # ---- resident arm
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local)
    if rank == 0:
        sampler.start()  # samples every 50 ms from the warm-up to the end of the end-to-end arm
    for k in range(args.warmup):
        step_resident(g, k)
    ms, launches = timed(g, step_resident, args.warmup, args.steps)
    kernel_ms, stages, diff_ms = stage_times(g, args.warmup + args.steps, min(args.steps, 8))
    info = update_info(g)
    value = world * args.steps / (ms * 1e-3)

    # ---- end-to-end arm (host buffers in, host permutation out)
    e2e = None
    if not args.no_e2e:
        k0 = args.warmup + args.steps + 8
        for k in range(2):
            step_e2e(g, k0 + k)
        ms_e, _ = timed(g, step_e2e, k0 + 2, args.steps)
        e2e = {"value": world * args.steps / (ms_e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": 4 * (N + 1 + nnzA), "d2h_bytes_per_step": 4 * M,
               "ms_per_step": ms_e / args.steps}

    # ---- the other variants of the same workload (resident arm, a few steps each): what the update
    # costs when the dirty nodes really are re-ordered (stage c: extraction + batched AMD)
    variants = {}
    if not args.no_variants:
        for v in ("lag", "aggr", "nolag", "nolag_amd"):
            if v == args.variant:
                continue
            gv = make_handle(v)
            vs = max(2, min(args.steps, 5))
            for k in range(2):
                step_resident(gv, k)
            ms_v, l_v = timed(gv, step_resident, 2, vs)
            _, st_v, _ = stage_times(gv, 2 + vs, 2)
            sv = np.mean(np.array(st_v), axis=0)
            variants[v] = {"what": VARIANT_TEXT[v], "updates_per_s": world * vs / (ms_v * 1e-3), "ms_per_step": ms_v / vs,
                           "steps": vs, "gpu_launches": int(l_v),
                           "stage_ms": {"build": float(sv[0]), "detect": float(sv[1]), "dirty_sets": float(sv[2]),
                                        "reorder": float(sv[3]), "expand": float(sv[4])},
                           "last_update": update_info(gv)}
            del gv

    clocks = sampler.stop()
    Mn, nnzM = g.mesh_dims()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the pipeline kernel also does stage (b): it compares every tile with the snapshot in shared memory and writes
    # the changed-row bitmap.  Its algorithmic bytes count every array ONCE -- stage (a)'s input and output plus the
    # snapshot graph -- not SURVEY 8d's sum of the two unfused stages (that sum, which would put the fraction above 1
    # because the new graph is never re-read, is reported beside it as algorithmic_bytes_unfused)
    fused = bool(g.stats().get("build_fused_detect", 0))
    alg_build, alg_detect = workloads.build_bytes(N, nnzA, Mn, nnzM, 1), workloads.detect_bytes(Mn, nnzM)
    alg = alg_build + (alg_detect // 2 if fused else 0)
    kms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
    achieved = alg / (kms * 1e-3) / 1e9
    st = np.mean(np.array(stages), axis=0) if stages else np.zeros(5)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": config_dict(args.n, args.variant, world) if world == 1 else
        dict(config_dict(args.n, args.variant, 1), parallelism=f"{world} independent meshes (one per GPU), no data-path collective"),
        "gnnz_per_s": value * nnzA / 1e9,
        "roofline": {"bound": "hbm", "kernel": "build_pipe_kernel" + (" (stage a + fused stage b)" if fused else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/)
                     "traffic": TRAFFIC.get((args.n, fused)), "traffic_source": "profiles/r08_bench_ncu_summary.md (ncu --set full of "
                     "the same kernel and workload; not re-measured by this run)", "algorithmic_bytes": alg,
                     "algorithmic_bytes_stage_a": alg_build, "algorithmic_bytes_snapshot_read": alg_detect // 2 if fused else 0,
                     "algorithmic_bytes_unfused": alg_build + (alg_detect if fused else 0),
                     "kernel_ms": kms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s"},
        "roofline_detect": ({"kernel": "diff_rows_kernel", "kernel_ms": diff_ms,
                             "note": "stage b is fused into the build kernel; this launch only visits tiles the build "
                                     "kernel could not stage (none here)"} if fused else
                            {"bound": "hbm", "kernel": "diff_rows_kernel", "algorithmic_bytes": alg_detect,
                             "kernel_ms": diff_ms, "achieved": alg_detect / (diff_ms * 1e-3) / 1e9 if diff_ms else None,
                             "frac": alg_detect / (diff_ms * 1e-3) / 1e9 / peak if diff_ms else None,
                             "note": "runs inside setMatrix (incremental symmetry proof) and is reused by change detection"}),
        "stage_ms": {"build": float(st[0]), "detect": float(st[1]), "dirty_sets": float(st[2]),
                     "reorder": float(st[3]), "expand": float(st[4])},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "last_update": info, "variants": variants,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_parth
            kind = cpu_kind()
            if kind == "port":
                cpu_parth.build("port")
            cores = os.cpu_count() or 1
            times, cinfo = cpu_update_stream(kind, stream, args.variant, 2, 0)
            ref_perm = cinfo.pop("_perm")
            ref_first = cinfo.pop("_first")
            line["cpu_baseline"] = {"value": 1.0 / float(np.mean(times)), "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"2 full updates at {args.n}^3 on a freshly injected tree",
                                    "last_update": cinfo}
            if not args.no_nnzl:
                line.update(nnz_l_block(make_handle, stream, args.variant, ref_perm, kind, ref_first))
        except Exception as e:  # the checker must never take the product number down
            line.setdefault("cpu_baseline", {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)})
            line.setdefault("nnz_L", None)
            line["checker_error"] = repr(e)
    if rank == 0 and world == 1 and not args.no_configs and emit:
        import bench_configs
        del dev_mats, host_mats, g
        torch.cuda.empty_cache()
        line["configs"] = bench_configs.run_all(api, cpu_kind(), peak, tuple(args.configs.split(",")))
    if not emit:
        return line
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ our arm, N > 1: one mesh over the N GPUs
def run_sharded(args):
    import torch
    import torch.distributed as dist
    from parth_b200 import api, workloads

    world, rank, local = _init_dist()
    dev = torch.device("cuda", local)
    stream = workloads.PoissonUpdateStream(n=args.n, levels=8)
    N = M = stream.N

    g = api.ParthAPI(local)
    g.verbose = False
    g.ND_levels = stream.levels
    apply_variant(g, args.variant, api)
    uid = [api.ParthAPI.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.commInit(uid[0], rank, world)  # the library's own NCCL communicator (C++ host, csrc/comm.cu)
    g.import_tree(stream.tree())
    c0, c1 = g.blockRange(N, int(stream.Ap[-1]))

    def block_of(Ap, Ai):
        """this rank's block of a matrix: pinned host slices + device copies, addressed in GLOBAL indexing (the
        pointers are biased by -c0 / -Ap[c0]; the Ai slices are placed so that the biased pointer is 16-byte aligned)"""
        q0, q1, nnz = int(Ap[c0]), int(Ap[c1]), int(Ap[-1])
        pad = q0 & 3
        hAp = torch.from_numpy(np.ascontiguousarray(Ap[c0:c1 + 1])).pin_memory()
        hAi = torch.zeros(pad + (q1 - q0) + 8, dtype=torch.int32).pin_memory()
        hAi[pad:pad + q1 - q0] = torch.from_numpy(np.ascontiguousarray(Ai[q0:q1]))
        dAp, dAi = hAp.to(dev), hAi.to(dev)
        return dict(nnz=nnz, q0=q0, q1=q1, keep=(hAp, hAi, dAp, dAi),
                    h=(hAp.data_ptr() - 4 * c0, hAi.data_ptr() + 4 * pad - 4 * q0),
                    d=(dAp.data_ptr() - 4 * c0, dAi.data_ptr() + 4 * pad - 4 * q0))

    base = block_of(stream.Ap, stream.Ai)
    g.setMatrixBlockPtr(N, base["nnz"], base["h"][0], base["h"][1], 1)
    g.acceptSnapshot()  # the tree fixture belongs to the base matrix

    need = args.warmup + args.steps + min(args.steps, 8) + (0 if args.no_e2e else 2 + args.steps + max(2, args.steps // 2))
    D = min(args.steps + args.warmup, 12) if args.variant == "lag" else min(need, 96)  # blocks are 1 / world of a matrix
    mats = [block_of(*stream.matrix(k)) for k in range(D)]  # the same update stream on every rank
    dperm = torch.empty(M, dtype=torch.int32, device=dev)
    hperm = torch.empty(M, dtype=torch.int32).pin_memory()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k):
        b = mats[k % D]
        g.setMatrixBlockDev(N, b["nnz"], b["d"][0], b["d"][1], b["q0"], b["q1"], 1)
        g.computePermutationDev(dperm.data_ptr(), 1)

    def step_e2e(k):
        """distributed in, distributed out: this rank uploads its block of the matrix and downloads its block of the
        permutation (parth_b200_compute_permutation_block)"""
        b = mats[k % D]
        g.setMatrixBlockPtr(N, b["nnz"], b["h"][0], b["h"][1], 1)
        g.computePermutationBlock(1, out=hperm.numpy())

    def step_e2e_full(k):
        """the whole permutation on EVERY rank's host"""
        b = mats[k % D]
        g.setMatrixBlockPtr(N, b["nnz"], b["h"][0], b["h"][1], 1)
        g.computePermutation(1, out=hperm.numpy())

    def timed(fn, k0, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = g.stats()["kernels_launched"]
        with torch.cuda.stream(torch.cuda.ExternalStream(g.stream(), device=dev)):
            e0.record()
            for k in range(k0, k0 + steps):
                fn(k)
            e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), g.stats()["kernels_launched"] - launches0

    sampler = ClockSampler(",".join(str(i) for i in range(world)))
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        step_resident(k)
    ms, launches = timed(step_resident, args.warmup, args.steps)
    kernel_ms, stages = [], []
    for k in range(args.warmup + args.steps, args.warmup + args.steps + min(args.steps, 8)):
        b = mats[k % D]
        g.setMatrixBlockDev(N, b["nnz"], b["d"][0], b["d"][1], b["q0"], b["q1"], 1)
        s0 = g.stats()
        g.computePermutationDev(dperm.data_ptr(), 1)
        s1 = g.stats()
        kernel_ms.append(s0["build_kernel_ms"])
        stages.append((s0["build_ms"], s1["detect_ms"], s1["dirty_ms"], s1["reorder_ms"], s1["expand_ms"]))
    info = dict(reuse=g.getReuse(), changes=g.getNumChanges(), dirty=[x.tolist() for x in g.dirty()])
    value = args.steps / (ms * 1e-3)  # ONE mesh: the job's updates per second

    e2e = None
    if not args.no_e2e:
        k0 = args.warmup + args.steps + 8
        for k in range(2):
            step_e2e(k0 + k)
        ms_e, _ = timed(step_e2e, k0 + 2, args.steps)
        half = max(2, args.steps // 2)
        ms_f, _ = timed(step_e2e_full, k0 + 2 + args.steps, half)
        b = mats[0]
        mine = 4 * ((c1 - c0 + 1) + (b["q1"] - b["q0"]))
        t = torch.tensor([mine, 4 * (c1 - c0)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e = {"value": args.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(t[0].item()),
               "d2h_bytes_per_step": int(t[1].item()), "ms_per_step": ms_e / args.steps,
               "note": f"bytes are totals over the {world} ranks: every rank uploads its block of the matrix ({mine} bytes on "
                       f"rank 0) and downloads its block of the permutation ({4 * (c1 - c0)} bytes on rank 0); "
                       "parth_b200_compute_permutation_block",
               "full_result_on_every_rank": {"value": half / (ms_f * 1e-3), "ms_per_step": ms_f / half, "steps": half,
                                             "d2h_bytes_per_step": 4 * M * world,
                                             "note": "parth_b200_compute_permutation: every rank also downloads the whole "
                                                     "permutation"}}
    clocks = sampler.stop()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    b = mats[0]
    cols, qa = c1 - c0, b["q1"] - b["q0"]
    # this rank's launch: its block of Ap / Ai in, its rows of the graph out, its rows of the snapshot in (every array once)
    alg = 4 * ((cols + 1) + qa + (cols + 1) + (qa - cols) + (cols + 1) + (qa - cols))
    kms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
    achieved = alg / (kms * 1e-3) / 1e9
    st = np.mean(np.array(stages), axis=0) if stages else np.zeros(5)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": config_dict(args.n, args.variant, world),
        "gnnz_per_s": value * mats[0]["nnz"] / 1e9,
        "roofline": {"bound": "hbm", "kernel": "build_pipe_kernel (stage a + fused stage b), rank 0's row block",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "algorithmic_bytes": alg, "kernel_ms": kms, "columns": [c0, c1],
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s"},
        "stage_ms": {"build": float(st[0]), "detect": float(st[1]), "dirty_sets": float(st[2]),
                     "reorder": float(st[3]), "expand": float(st[4])},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "last_update": info,
        "collectives": {"library": "NCCL through the C++ host (parth_b200/csrc/comm.cu), communicator " + g.commInfo()[2],
                        "per_update": "1 all-gather of one slot per rank (header, node flags, changed edges, row diffs); "
                                      "a re-ordering update adds 2 all-reduces for the dirty sub-mesh (counts, entries) and, "
                                      "with more than one dirty graph, 1 for the local permutations"},
    }
    del g, mats
    torch.cuda.empty_cache()
    if not args.no_variants:
        # the N independent meshes of round 1 (weak scaling, no collective), a few steps
        import copy
        a2 = copy.copy(args)
        a2.steps, a2.warmup, a2.no_variants, a2.no_cpu_baseline = max(2, min(args.steps, 5)), 3, True, True
        rep = run_ours(a2, emit=False)
        line["replicas"] = {"what": f"{world} independent meshes, one per GPU, no collective (weak scaling)",
                            "value": rep["value"], "unit": UNIT, "ms_per_step": rep["ms_per_step"], "steps": a2.steps,
                            "e2e": rep["e2e"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def _own_stdout():
    """stdout carries ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints its
    "NCCL version ..." banner there on some boxes): descriptor 1 is pointed at stderr for the whole run and the JSON
    line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w", buffering=1)


if __name__ == "__main__":
    a = parse()
    _own_stdout()
    if a.impl == "reference":
        run_reference(a)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and not a.replicas_only:
        run_sharded(a)
    else:
        run_ours(a)
