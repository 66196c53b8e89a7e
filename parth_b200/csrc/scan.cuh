// parth_b200/csrc/scan.cuh -- device-wide exclusive scan (single pass, decoupled look-back).
#pragma once
#include "common.cuh"

namespace pb {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// ws: [0] = dynamic tile counter (as u64), [1..] tile descriptors; must be zeroed before launch.
// out[i] = sum_{j<i} f(in[j]) for i in [0, n]; out has n+1 entries (out[n] = total).
// MODE 0: f(x) = x;  MODE 1: f(x) = popc(x);  MODE 2: f(x) = (x != 0);  MODE 3: f(x) = (x == -1)
template <int MODE>
__global__ void __launch_bounds__(kScanThreads)
scan_excl_kernel(const int *in, int *out, int n, unsigned long long *ws) {
  __shared__ int s_scan[33];
  __shared__ int s_tile;
  __shared__ long long s_base;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ws, 1ull);
  __syncthreads();
  const int tile = s_tile;
  const long long base = (long long)tile * kScanTile;
  int v[kScanItems];
  int sum = 0;
  // full tiles of 16-byte aligned arrays move as 128-bit vectors (a thread owns 16 consecutive ints)
  const bool vec = base + kScanTile <= (long long)n &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  auto f = [](int x, bool inside) {
    if (MODE == 1) x = __popc((unsigned)x);
    if (MODE == 2) x = x != 0;
    if (MODE == 3) x = inside && x == -1;
    return x;
  };
  if (vec) {
    const int4 *src = reinterpret_cast<const int4 *>(in + base) + threadIdx.x * (kScanItems / 4);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; q++) {
      const int4 x = ld_stream4(src + q);
      v[4 * q] = f(x.x, true); v[4 * q + 1] = f(x.y, true); v[4 * q + 2] = f(x.z, true); v[4 * q + 3] = f(x.w, true);
    }
#pragma unroll
    for (int k = 0; k < kScanItems; k++) sum += v[k];
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + (long long)threadIdx.x * kScanItems + k;
      v[k] = f(i < n ? in[i] : 0, i < n);
      sum += v[k];
    }
  }
  int total;
  int excl = block_excl_scan<kScanThreads>(sum, s_scan, &total);
  if (threadIdx.x < 32) {
    long long b = tile_lookback(ws + 1, tile, total);
    if (threadIdx.x == 0) s_base = b;
  }
  __syncthreads();
  long long run = s_base + excl;
  if (vec) {
    int4 *dst = reinterpret_cast<int4 *>(out + base) + threadIdx.x * (kScanItems / 4);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; q++) {
      int4 o;
      o.x = (int)run; run += v[4 * q];
      o.y = (int)run; run += v[4 * q + 1];
      o.z = (int)run; run += v[4 * q + 2];
      o.w = (int)run; run += v[4 * q + 3];
      dst[q] = o;
    }
    if (base + kScanTile == (long long)n && threadIdx.x == kScanThreads - 1) out[n] = (int)run;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      long long i = base + (long long)threadIdx.x * kScanItems + k;
      if (i < n) out[i] = (int)run;
      run += v[k];
      if (i == n - 1) out[n] = (int)run;
    }
  }
  if (n == 0 && tile == 0 && threadIdx.x == 0) out[0] = 0;
}

inline size_t scan_ws_words(size_t n) { return 2 + (n + kScanTile - 1) / kScanTile; }

// ws must hold scan_ws_words(n) u64 words.  Returns number of kernels launched.
template <int MODE>
inline int exclusive_scan(const int *in, int *out, int n, unsigned long long *ws, cudaStream_t s) {
  int tiles = n > 0 ? (n + kScanTile - 1) / kScanTile : 1;
  PB_CUDA(cudaMemsetAsync(ws, 0, sizeof(unsigned long long) * (size_t)(tiles + 1), s));
  scan_excl_kernel<MODE><<<tiles, kScanThreads, 0, s>>>(in, out, n, ws);
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
