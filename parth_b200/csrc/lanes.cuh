// parth_b200/csrc/lanes.cuh -- the warp / CTA vocabulary the cooperative kernels are written in.
//
// amd_core.cuh is written against these few functions only, so the same source compiles twice:
//   * with nvcc for sm_100a: every function is the warp intrinsic of the same meaning;
//   * with g++ for tests/amd_model (TEST INFRASTRUCTURE): tests/amd_model/lanes_host.h runs each
//     lane as an OS thread and implements the collectives with barriers, which lets the CPU suite
//     check the cooperative algorithm bit for bit against the oracle without a GPU.
// Every collective is called by ALL lanes of a warp (full mask), never from divergent code.
// Measured on B200 (scripts/warp_ops_probe.cu, dependent chain): shfl 36, ballot 27, redux.add 51,
// redux.min 22, LDS 34 cycles; match_any costs ~11 cycles PER DISTINCT VALUE (64 cycles for 4 groups,
// 373 for 32 distinct values) -- so idle lanes pass one shared dummy key, never a per-lane one.
#pragma once
#ifdef __CUDACC__

#define PB_DEV __device__ __forceinline__
#define PB_DEVFN __device__

namespace pb {
namespace lanes {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

PB_DEV int lane_id() { return threadIdx.x & 31; }
PB_DEV int thread_id() { return threadIdx.x; }
PB_DEV int num_threads() { return blockDim.x; }
PB_DEV void sync_warp() { __syncwarp(); }
PB_DEV void sync_cta() { __syncthreads(); }
PB_DEV unsigned ballot(bool p) { return __ballot_sync(FULL, p); }
PB_DEV int shfl(int v, int src) { return __shfl_sync(FULL, v, src); }
PB_DEV unsigned match_any(int v) { return __match_any_sync(FULL, v); }
PB_DEV int warp_sum(int v) { return __reduce_add_sync(FULL, v); }
PB_DEV int warp_min(int v) { return __reduce_min_sync(FULL, v); }
PB_DEV int warp_max(int v) { return __reduce_max_sync(FULL, v); }
PB_DEV unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }
PB_DEV unsigned lanemask_gt() { return lane_id() == 31 ? 0u : (FULL << (lane_id() + 1)); }
PB_DEV int popc(unsigned m) { return __popc(m); }
PB_DEV int lowest(unsigned m) { return __ffs(m) - 1; }          // m != 0
PB_DEV int highest(unsigned m) { return 31 - __clz(m); }        // m != 0
PB_DEV void atomic_add(int *p, int v) { atomicAdd(p, v); }
PB_DEV int atomic_add_ret(int *p, int v) { return atomicAdd(p, v); }
// inclusive prefix sum across the warp
PB_DEV int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FULL, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

}  // namespace lanes
}  // namespace pb

#endif  // __CUDACC__
