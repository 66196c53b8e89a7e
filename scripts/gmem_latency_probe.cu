// developer probe: dependent-chain latency of GLOBAL memory accesses as the latency-bound kernels see them (one warp,
// sm_100a): pointer chase through buffers of different footprints with ld.global.ca / .cg / ld.relaxed.gpu, a
// dependent atomicAdd chain, store -> load visibility between two SMs.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ int ld_relaxed(const int *p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
template <int MODE>
__global__ void chase(const int *buf, int start, int n, long long *out, int *sink) {
  int v = start;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) v = MODE == 0 ? buf[v] : MODE == 1 ? __ldcg(buf + v) : ld_relaxed(buf + v);
  long long t1 = clock64();
  if (threadIdx.x == 0) { *out = (t1 - t0) / n; *sink = v; }
}
__global__ void atom_chain(int *buf, int mask, int n, long long *out, int *sink) {
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) v = (atomicAdd(buf + ((v * 977 + i * 131) & mask), 1) + v * 31 + i) ;
  long long t1 = clock64();
  if (threadIdx.x == 0) { *out = (t1 - t0) / n; *sink = v; }
}
// ping-pong between two CTAs (different SMs) through one word each: round trip = 2 x (store visible at the other SM)
__global__ void pingpong(int *flag, int n, long long *out) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  if (blockIdx.x == 0) {
    for (int i = 1; i <= n; i++) {
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(i) : "memory");
      while (ld_relaxed(flag + 32) != i) {}
    }
    *out = (clock64() - t0) / n;
  } else if (blockIdx.x == gridDim.x - 1) {
    for (int i = 1; i <= n; i++) {
      while (ld_relaxed(flag) != i) {}
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flag + 32), "r"(i) : "memory");
    }
  }
}
int main() {
  long long *d_out; int *d_sink; cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4);
  long long h;
  for (size_t mb : {1, 8, 64, 512, 4096}) {
    const size_t n = mb * 1024 * 1024 / 4;
    std::vector<int> hbuf(n);
    // random cyclic permutation with a 128-byte stride granularity (every hop a new line)
    const size_t lines = n / 32;
    std::vector<unsigned> perm(lines);
    for (size_t i = 0; i < lines; i++) perm[i] = (unsigned)i;
    srand(1);
    for (size_t i = lines - 1; i > 0; i--) { size_t j = ((size_t)rand() * RAND_MAX + rand()) % (i + 1); std::swap(perm[i], perm[j]); }
    for (size_t i = 0; i < lines; i++) hbuf[(size_t)perm[i] * 32] = (int)(perm[(i + 1) % lines] * 32);
    int *d; cudaMalloc(&d, n * 4); cudaMemcpy(d, hbuf.data(), n * 4, cudaMemcpyHostToDevice);
    const int hops = 20000;
    for (int rep = 0; rep < 2; rep++) {  // second run: whatever fits in L2 is resident (the first hops/lines of it)
      chase<0><<<1, 32>>>(d, 0, hops, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      long long a = h;
      chase<1><<<1, 32>>>(d, 0, hops, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      long long b = h;
      chase<2><<<1, 32>>>(d, 0, hops, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      printf("footprint %5zu MB (%d hops over %d lines, run %d): ld %lld  ld.cg %lld  ld.relaxed.gpu %lld cycles per dependent load\n",
             mb, hops, hops, rep, a, b, h);
    }
    cudaFree(d);
  }
  int *d; cudaMalloc(&d, 64 << 20); cudaMemset(d, 0, 64 << 20);
  for (int mask : {0, 1023, (16 << 20) - 1}) {
    atom_chain<<<1, 1>>>(d, mask, 5000, d_out, d_sink); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("dependent atomicAdd chain, addresses & 0x%x: %lld cycles per op\n", mask, h);
  }
  for (int grid : {2, 148}) {
    cudaMemset(d, 0, 1024);
    pingpong<<<grid, 32>>>(d, 5000, d_out); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("ping-pong CTA 0 <-> CTA %d through L2 (st.relaxed.gpu -> poll): %lld cycles per round trip\n", grid - 1, h);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("SM clock attr %d kHz\n", clk);
  return 0;
}
