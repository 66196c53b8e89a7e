// parth_b200/csrc/common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace pb {

constexpr int kNumSMsB200 = 148;

struct CudaError : std::runtime_error {
  int code;
  CudaError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define PB_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw pb::CudaError((int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + \
                                       __FILE__ + ":" + std::to_string(__LINE__));           \
  } while (0)

// Grow-only device buffer: steady-state updates allocate nothing.  Growth goes through the
// stream-ordered allocator (cudaMallocAsync on the handle's stream, the device's default memory
// pool with its release threshold raised at handle creation): a bigger buffer is carved out of
// memory the pool already holds and the old one goes back to the pool without a device-wide
// synchronisation -- plain cudaMalloc / cudaFree were measured at up to 0.7 s per growth step when
// the dirty set of a mesh keeps growing (profiles/r02_configs.md).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void release() {
    if (p) cudaFree(p);  // also valid for stream-ordered allocations; synchronises
    p = nullptr;
    cap = 0;
  }
  // contents are NOT preserved across growth unless keep == true
  void reserve(size_t n, cudaStream_t s = 0, bool keep = false, size_t keep_n = 0) {
    if (n <= cap) return;
    size_t ncap = n + n / 2 + 64;  // geometric: a growing work-load re-allocates O(log) times
    T *np = nullptr;
    PB_CUDA(cudaMallocAsync(&np, ncap * sizeof(T), s));
    if (keep && p && keep_n)
      PB_CUDA(cudaMemcpyAsync(np, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) PB_CUDA(cudaFreeAsync(p, s));  // ordered behind everything already enqueued on s
    p = np;
    cap = ncap;
  }
  void swap(DevBuf &o) {
    T *tp = p; p = o.p; o.p = tp;
    size_t tc = cap; cap = o.cap; o.cap = tc;
  }
};

// Pinned host staging buffer (grow-only).
template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) cudaFreeHost(p);
    cap = n + n / 8 + 64;
    PB_CUDA(cudaMallocHost(&p, cap * sizeof(T)));
  }
};

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Programmatic dependent launch: the grid may be set up and its CTAs scheduled while the previous kernel of the
// stream is still draining; the kernel must call pdl_wait() before it touches anything that kernel wrote (the
// wait returns once the previous grid has completed and its writes are visible).  Hides the 2-3 us of launch
// latency between the small dependent kernels that follow the build kernel.  pdl_wait() is a no-op under a
// plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// streaming (read-once) 32-bit load: keep it out of L1 so mirror/gather traffic keeps its lines
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4 *p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(int *p, int v) {
  asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(int4 *p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = sum.
// smem: int[33].  All threads of the block must call.
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *smem, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WARPS = THREADS / 32;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < WARPS ? smem[lane] : 0;
    int wi = warp_incl_scan(w, lane);
    if (lane < WARPS) smem[lane] = wi - w;
    if (lane == WARPS - 1) smem[32] = wi;
  }
  __syncthreads();
  int res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

// ---- decoupled look-back tile prefix (single-pass chained scan) ----------------------------
// desc[tile]: bits 63..62 status (0 = empty, 1 = aggregate, 2 = inclusive prefix), bits 61..0 value.
constexpr unsigned long long kTileAgg = 1ull << 62;
constexpr unsigned long long kTilePre = 2ull << 62;
constexpr unsigned long long kTileVal = (1ull << 62) - 1;

// Sums are taken modulo 2^62, so callers may pack two independent fields into one value (the
// upper field then wraps modulo its width).
// Called by ALL lanes of warp 0 of the block (tile ids must be handed out in launch order by an
// atomic counter so predecessors are always resident or finished).  Returns the exclusive prefix
// of this tile in every lane and publishes the inclusive prefix.
__device__ __forceinline__ long long tile_lookback(unsigned long long *desc, int tile,
                                                   long long aggregate) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(desc, kTilePre | (unsigned long long)aggregate);
    return 0;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, kTileAgg | (unsigned long long)aggregate);
  long long excl = 0;
  int look = tile - 1;
  // A walker retires kLookWide*32 predecessors per L2 round trip.  The chained scan sustains at
  // most that many tiles per round trip chip-wide, so the window (and the tile size) must be
  // large enough that this bound sits below the HBM time of the kernel.
  constexpr int kLookWide = 4;
  bool done = false;
  while (!done) {
    unsigned long long d[kLookWide];
#pragma unroll
    for (int u = 0; u < kLookWide; u++) {
      const int idx = look - lane - 32 * u;
      // lanes past the front behave like a zero inclusive prefix
      d[u] = idx >= 0 ? ld_relaxed_u64(desc + idx) : kTilePre;
    }
#pragma unroll
    for (int u = 0; u < kLookWide; u++) {
      if (done) break;
      const int idx = look - lane - 32 * u;
      while ((d[u] >> 62) == 0) {  // predecessor has not published yet: back off, then re-read
        __nanosleep(32);
        d[u] = ld_relaxed_u64(desc + idx);
      }
      const unsigned full = __ballot_sync(0xffffffffu, (d[u] >> 62) == 2);
      const int first = full ? __ffs(full) - 1 : 32;  // nearest predecessor with an inclusive prefix
      long long v = (lane <= first) ? (long long)(d[u] & kTileVal) : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl = (excl + v) & (long long)kTileVal;
      if (full) done = true;
    }
    look -= 32 * kLookWide;
  }
  if (lane == 0)
    st_relaxed_u64(desc + tile, kTilePre | ((unsigned long long)(excl + aggregate) & kTileVal));
  return excl;
}

#endif  // __CUDACC__

}  // namespace pb
