// scripts/amd_prof.cu -- phase-level cycle profile of the cooperative AMD (developer tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPB_AMD_PROF -I parth_b200/csrc \
//        scripts/amd_prof.cu -o scripts/_bin/amd_prof ; scripts/_bin/amd_prof 2d 128 | 3d 31
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "lanes.cuh"
#include "amd_core.cuh"

using namespace pb;

__global__ void __launch_bounds__(256) prof_kernel(int n, const int *Cp, const int *Ci, int *slab, long long slab_ints,
                                                   int *P, int smem_ints, int force_global) {
  __shared__ int sh[36];
  extern __shared__ int amd_smem[];
  const int nz = Cp[n];
  const bool sym = amdk::rows_are_symmetric(n, Cp, 0, Ci, &sh[0]);
  const long long need = 8LL * n + nz + nz / 5 + n + 16;
  if (!force_global && need + n <= smem_ints) {
    amdk::order_graph(n, Cp, 0, Ci, sym, amd_smem, need, amd_smem + need, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) P[i] = amd_smem[need + i];
  } else {
    amdk::order_graph(n, Cp, 0, Ci, sym, slab, slab_ints, P, sh);
  }
}

int main(int argc, char **argv) {
  const bool d3 = argc > 1 && !strcmp(argv[1], "3d");
  const int k = argc > 2 ? atoi(argv[2]) : 40;
  const int force_global = argc > 3 ? atoi(argv[3]) : 0;
  const int n = d3 ? k * k * k : k * k;
  std::vector<int> Cp(n + 1, 0), Ci;
  for (int v = 0; v < n; v++) {
    const int x = v % k, y = (v / k) % k, z = v / (k * k);
    if (d3 && z > 0) Ci.push_back(v - k * k);
    if (y > 0) Ci.push_back(v - k);
    if (x > 0) Ci.push_back(v - 1);
    if (x < k - 1) Ci.push_back(v + 1);
    if (y < k - 1) Ci.push_back(v + k);
    if (d3 && z < k - 1) Ci.push_back(v + k * k);
    Cp[v + 1] = (int)Ci.size();
  }
  const int nz = (int)Ci.size();
  int *dCp, *dCi, *slab, *P;
  const long long slab_ints = 8LL * n + 2LL * nz + 2LL * nz / 5 + n + 16;
  cudaMalloc(&dCp, 4 * (n + 1)); cudaMalloc(&dCi, 4 * nz); cudaMalloc(&slab, 4 * slab_ints); cudaMalloc(&P, 4 * n);
  cudaMemcpy(dCp, Cp.data(), 4 * (n + 1), cudaMemcpyHostToDevice);
  cudaMemcpy(dCi, Ci.data(), 4 * nz, cudaMemcpyHostToDevice);
  int smem = 0;
  cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0);
  smem -= 1024;
  cudaFuncSetAttribute(prof_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; rep++) {
    long long zero[16] = {0};
#ifdef PB_AMD_PROF
    cudaMemcpyToSymbol(amdk::g_amd_prof, zero, sizeof(zero));
#endif
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    prof_kernel<<<1, 256, smem>>>(n, dCp, dCi, slab, slab_ints, P, smem / 4, force_global);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long pr[16] = {0};
#ifdef PB_AMD_PROF
    cudaMemcpyFromSymbol(pr, amdk::g_amd_prof, sizeof(pr));
#else
    pr[8] = 1;
    (void)zero;
#endif
    if (rep == 0) continue;
    const char *nm[8] = {"P0 pick", "P1 Lme", "P2 |Le-Lme|", "P3 degrees", "P4 supervar", "P5 reinsert", "init", "tail"};
    printf("%s k=%d n=%d nz=%d  %s  %.3f ms (%s)  pivots=%lld  avg|Lme|=%.1f avg elenme=%.2f\n", d3 ? "3d" : "2d", k, n, nz,
           cudaGetErrorString(e), ms, force_global ? "global" : "smem if it fits", pr[8], (double)pr[9] / pr[8], (double)pr[10] / pr[8]);
    long long tot = 0;
    for (int i = 0; i < 8; i++) tot += pr[i];
    for (int i = 0; i < 8; i++)
      printf("  %-12s %10.3f Mcycles  %5.1f%%  %8.0f cycles/pivot\n", nm[i], pr[i] / 1e6, 100.0 * pr[i] / tot, (double)pr[i] / pr[8]);
  }
  return 0;
}
