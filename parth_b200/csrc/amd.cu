// parth_b200/csrc/amd.cu -- stage (c), ordering: batched approximate-minimum-degree ordering of
// the extracted sub-meshes of dirty nodes, one CTA per sub-mesh.
//
// Replaces HMD::AMDNodeNDPermutation -> amd_order(N, Mp, Mi, perm, NULL, NULL)
// (reference src/Parth/HMD.cpp:58-75).  amd_order is SuiteSparse AMD (third party, absent from the
// reference tree, pinned v7.11.0); the algorithm here is the published one (Amestoy, Davis, Duff
// 1996 / ACM TOMS 837) with the default controls that Control == NULL selects (dense = 10,
// aggressive absorption), and it is held bit-exact to the CPU restatement oracle/amd_restated.c:
// same A+A' construction sweep (the entry order of every adjacency list is part of the
// tie-breaking), same degree lists with head insertion, element absorption, mass elimination,
// supervariable hashing, workspace compaction and assembly-tree post-order.
//
// Minimum degree is a pivot recurrence, so the pivots of one sub-mesh stay in order; the work
// inside a pivot (building the element, external degrees, list compaction, hashing, supervariable
// comparison, degree-list updates) is spread over the lanes of a warp, the set-up and the
// epilogue over the whole CTA (amd_core.cuh).  Many dirty nodes are in flight at once, one CTA
// each.  This stage is latency bound and is reported as nodes/s and DOFs/s, outside the
// HBM-roofline claim.
#include <climits>

#include "reorder.cuh"
#include "lanes.cuh"
#include "amd_core.cuh"

namespace pb {

constexpr int kAmdThreads = 256;

// ints of shared memory a graph needs to be ordered out of shared memory (8 n-arrays, Iw with
// amd_order's slack, the pivot array); nzaat = nnz for symmetric rows
__host__ __device__ inline long long amd_smem_need(int n, long long nzaat) {
  return 8 * (long long)n + nzaat + nzaat / 5 + n + 16 + n;
}

// One CTA per graph g.  Graph g has vertices [vtx_ptr[g], vtx_ptr[g+1]) of the concatenated vertex
// space; its CSR is (G + vtx_ptr[g]) with global entry offsets (subtract G[vtx_ptr[g]]) into
// sub_Mi.  Work-space slab of graph g: ws + ws_ptr[g], amd_slab_ints(n, nnz) ints.
// STAGED = true : graphs whose quotient graph fits smem_ints are ordered out of shared memory
//                 (separators, small leaves); the others are left to the second launch
// STAGED = false: the remaining graphs, work arrays in their slab (L2 / L1 resident), no dynamic
//                 shared memory so that several CTAs share an SM
// perm_out[vtx_ptr[g] + k] = k-th pivot (local id).  Zero-edge graphs get the identity
// (HMD.cpp:62-72).
template <bool STAGED>
__global__ void __launch_bounds__(kAmdThreads)
amd_batch_kernel(int count, const int *__restrict__ vtx_ptr, const int *__restrict__ G,
                 const int *__restrict__ sub_Mi, const long long *__restrict__ ws_ptr,
                 int *__restrict__ ws, int *__restrict__ perm_out, int smem_ints, const int *__restrict__ glist) {
  const int g = glist ? glist[blockIdx.x] : blockIdx.x;
  if (g >= count) return;
  const int v0 = vtx_ptr[g], n = vtx_ptr[g + 1] - v0;
  if (n <= 0) return;
  const int *Cp = G + v0;
  const int e0 = Cp[0];
  const int nz = Cp[n] - e0;
  const bool fits = amd_smem_need(n, nz) <= (long long)smem_ints;
  if (fits != STAGED) return;
  int *P = perm_out + v0;
  if (nz == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) P[i] = i;
    return;
  }
  const int *Ci = sub_Mi + e0;
  __shared__ int sh[36];
  extern __shared__ int amd_smem[];
  const bool sym = amdk::rows_are_symmetric(n, Cp, e0, Ci, &sh[0]);
  const long long need = amd_smem_need(n, sym ? (long long)nz : 2 * (long long)nz);
  if (STAGED && need <= (long long)smem_ints) {
    amdk::order_graph(n, Cp, e0, Ci, sym, amd_smem, need - n, amd_smem + (need - n), sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) P[i] = amd_smem[need - n + i];
  } else {
    amdk::order_graph(n, Cp, e0, Ci, sym, ws + ws_ptr[g], ws_ptr[g + 1] - ws_ptr[g], P, sh);
  }
}

// max_stage_ints: the largest amd_smem_need() among the graphs of the batch that fit the device's
// shared memory (0 = none fits / unknown -> stage with the device maximum)
int launch_amd_batch(int count, const int *vtx_ptr, const int *G, const int *sub_Mi,
                     const long long *ws_ptr, int *ws, int *perm_out, long long max_stage_ints, bool any_unstaged,
                     cudaStream_t s, const int *glist, int n_list) {
  if (count <= 0) return 0;
  const int grid = glist ? n_list : count;
  if (grid <= 0) return 0;
  int launches = 0;
  const int cap_bytes = amd_stage_capacity_ints() * 4;
  // the smaller the staging area, the more CTAs share an SM
  long long want = max_stage_ints > 0 ? max_stage_ints * 4 : cap_bytes;
  if (want > cap_bytes) want = cap_bytes;
  want = (want + 1023) / 1024 * 1024;
  const int smem_bytes = (int)(want > cap_bytes ? cap_bytes : want);
  if (max_stage_ints != 0) {
    PB_CUDA(cudaFuncSetAttribute(amd_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    amd_batch_kernel<true><<<grid, kAmdThreads, smem_bytes, s>>>(count, vtx_ptr, G, sub_Mi, ws_ptr, ws, perm_out,
                                                                 smem_bytes / 4, glist);
    PB_CUDA(cudaGetLastError());
    launches++;
  }
  if (any_unstaged) {
    amd_batch_kernel<false><<<grid, kAmdThreads, 0, s>>>(count, vtx_ptr, G, sub_Mi, ws_ptr, ws, perm_out,
                                                         smem_bytes / 4, glist);
    PB_CUDA(cudaGetLastError());
    launches++;
  }
  return launches;
}

int amd_stage_capacity_ints() {
  static int cap = -1;
  if (cap < 0) {
    int dev = 0, max_smem = 0;
    PB_CUDA(cudaGetDevice(&dev));
    PB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cap = (max_smem - 1024) / 4;
  }
  return cap;
}

long long amd_stage_need_ints(int n, int nnz) { return amd_smem_need(n, nnz); }

// host-side slab size of one graph (ints): 8 n-arrays + Iw sized for a pattern that is not
// symmetric (nzaat <= 2 nnz); symmetric rows simply get a roomier Iw (fewer compactions)
long long amd_slab_ints(int n, int nnz) {
  const long long nz = 2 * (long long)nnz;
  return 8 * (long long)n + nz + nz / 5 + (long long)n + 16;
}

}  // namespace pb
