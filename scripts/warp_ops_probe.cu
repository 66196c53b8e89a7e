// developer probe: dependent-chain latency of the warp collectives the cooperative AMD leans on (one warp, sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(long long *out, int *sink) {
  __shared__ int sm[1024];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7 + 3) & 1023;
  __syncwarp();
  int v = lane;
  const int N = 2000;
  long long t0, t1;
#define RUN(k, body) t0 = clock64(); for (int it = 0; it < N; it++) { body; } t1 = clock64(); if (lane == 0) out[k] = (t1 - t0) / N;
  RUN(0, v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + 1)
  RUN(1, v = (int)(__ballot_sync(0xffffffffu, v & 1) & 0xffff) + lane)
  RUN(2, v = (int)(__match_any_sync(0xffffffffu, v & 3) & 0xffff) + lane)
  RUN(3, v = (int)(__match_any_sync(0xffffffffu, v) & 0xffff) + lane + it)
  RUN(4, v = __reduce_add_sync(0xffffffffu, v & 7) + lane)
  RUN(5, v = __reduce_min_sync(0xffffffffu, v) + lane)
  RUN(6, v = sm[v & 1023])
  RUN(7, __syncwarp(); v += 1)
  RUN(8, v = atomicAdd(&sm[v & 1023], 1))
  RUN(9, { int t = v; for (int d = 1; d < 32; d <<= 1) { int x = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += x; } v = t & 1023; })
  RUN(10, v = (int)((unsigned)(v + 12345) % 1601u))
  *sink = v;
}
int main() {
  long long *d; int *s; cudaMalloc(&d, 16 * 8); cudaMalloc(&s, 4);
  probe<<<1, 32>>>(d, s); cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
  const char *nm[] = {"shfl", "ballot", "match_any (4 groups)", "match_any (32 distinct)", "redux.add", "redux.min", "LDS dependent", "syncwarp+add", "ATOMS dependent", "5-step shfl_up scan", "u32 % const-unknown"};
  for (int i = 0; i < 11; i++) printf("%-28s %lld cycles per dependent op\n", nm[i], h[i]);
  return 0;
}
