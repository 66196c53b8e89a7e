// parth_b200/csrc/supernodes.cu -- SURVEY 8 f4: the consumer of stage (d).  The elimination tree and
// the column counts stage_symbolic leaves on the device are turned, on the device, into the
// fundamental supernode partition of L and its supernodal elimination tree -- the first thing a
// supernodal symbolic phase does with them (the reference's solver wrapper hands the permutation to
// cholmod_analyze_p with cm.supernodal = CHOLMOD_SUPERNODAL, src/LinSysSolver/CHOLMODSolver.cpp:108-151,
// which recomputes etree and counts on the host before it finds the supernodes).
//
// Definition (Liu, Ng, Peyton 1993; what CHOLMOD's supernodal analysis starts from): column j > 0
// CONTINUES the supernode of column j-1 iff parent[j-1] == j, colcount[j-1] == colcount[j] + 1 and j
// has exactly one child in the elimination tree; otherwise j starts a new supernode.  Then
// struct(L(:,j-1)) \ {j-1} == struct(L(:,j)) for every continued column, which is the property a
// supernodal factorisation needs (tests check it against a brute-force symbolic factorisation).
// CHOLMOD's relaxed amalgamation (nrelax / zrelax) is a tuning step on top of this partition and is
// not reproduced: SuiteSparse is absent from the reference tree, so it could not be pinned.
//
// All passes stream over n ints: child counts (atomics), start flags, one exclusive scan, one
// scatter (supernode of every column, first column of every supernode) and one pass over the
// supernodes (parent supernode, |L| structure sizes).
#include <algorithm>

#include "scan.cuh"
#include "stages.cuh"

namespace pb {

namespace {

inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

__global__ void sn_child_count_kernel(const int *__restrict__ parent, int n, int *__restrict__ cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int p = parent[j];
  if (p >= 0 && p < n) atomicAdd(cnt + p, 1);
}

__global__ void sn_flag_kernel(const int *__restrict__ parent, const int *__restrict__ colcount,
                               const int *__restrict__ cnt, int n, int *__restrict__ flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int start = 1;
  if (j > 0) start = parent[j - 1] != j || colcount[j - 1] != colcount[j] + 1 || cnt[j] > 1;
  flag[j] = start;
}

// pos[j] = number of supernode starts before column j (exclusive scan of the flags), pos[n] = nsuper
__global__ void sn_scatter_kernel(const int *__restrict__ flag, const int *__restrict__ pos, int n,
                                  int *__restrict__ super, int *__restrict__ supermap) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int f = flag[j], s = pos[j] + f - 1;
  if (supermap) supermap[j] = s;
  if (super) {
    if (f) super[s] = j;
    if (j == n - 1) super[s + 1] = n;
  }
}

// one thread per supernode s (grid-stride; nsuper = pos[n] is read on the device): its last column is
// super[s+1] - 1; totals[0] = nsuper, totals[1] = sum of colcount[first column] (row indices of the
// supernodal structure, CHOLMOD's ssize), totals[2] = sum of ncols * colcount[first column] (numerical
// entries of the supernodes, xsize)
__global__ void sn_parent_kernel(const int *__restrict__ parent, const int *__restrict__ colcount,
                                 const int *__restrict__ flag, const int *__restrict__ pos,
                                 const int *__restrict__ super, int n, int *__restrict__ sparent,
                                 unsigned long long *__restrict__ totals) {
  const int ns = pos[n];
  unsigned long long ss = 0, xs = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const int first = super[s], last = super[s + 1] - 1;
    const int p = parent[last];
    if (sparent) sparent[s] = (p < 0 || p >= n) ? -1 : pos[p] + flag[p] - 1;
    const unsigned long long cc = (unsigned long long)colcount[first];
    ss += cc;
    xs += cc * (unsigned long long)(last - first + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    xs += __shfl_xor_sync(0xffffffffu, xs, o);
  }
  if ((threadIdx.x & 31) == 0 && ss) {
    atomicAdd(totals + 1, ss);
    atomicAdd(totals + 2, xs);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) totals[0] = (unsigned long long)ns;
}

}  // namespace

void stage_supernodes(parth_b200 &h, int n, const int *parent_in, const int *colcount_in, int *super_out,
                      int *supermap_out, int *sparent_out, int *nsuper, long long *ssize, long long *xsize,
                      bool on_device) {
  cudaStream_t s = h.stream;
  if (n == 0) {
    if (nsuper) *nsuper = 0;
    if (ssize) *ssize = 0;
    if (xsize) *xsize = 0;
    if (super_out) {
      const int zero = 0;
      PB_CUDA(cudaMemcpyAsync(super_out, &zero, sizeof(int), on_device ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, s));
      PB_CUDA(cudaStreamSynchronize(s));
    }
    return;
  }
  StageTimer tm(h, &h.stats.supernodes_ms);
  DevBuf<int> par_b, cc_b, cnt, flag, pos, sup, smap, spar;
  DevBuf<unsigned long long> totals;
  const int *parent = parent_in, *colcount = colcount_in;
  if (!on_device) {
    par_b.reserve((size_t)n, s); cc_b.reserve((size_t)n, s);
    PB_CUDA(cudaMemcpyAsync(par_b.p, parent_in, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
    PB_CUDA(cudaMemcpyAsync(cc_b.p, colcount_in, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
    h.stats.h2d_bytes += 2ll * (long long)sizeof(int) * n;
    parent = par_b.p; colcount = cc_b.p;
  }
  int *d_super = super_out, *d_smap = supermap_out, *d_spar = sparent_out;
  if (!on_device || !super_out) { sup.reserve((size_t)n + 1, s); d_super = sup.p; }  // the parent pass reads it
  if (!on_device) {
    if (supermap_out) { smap.reserve((size_t)n, s); d_smap = smap.p; }
    if (sparent_out) { spar.reserve((size_t)n, s); d_spar = spar.p; }
  }
  cnt.reserve((size_t)n, s); flag.reserve((size_t)n + 1, s); pos.reserve((size_t)n + 2, s); totals.reserve(4, s);
  h.scan_ws.reserve(scan_ws_words((size_t)n + 1), s);
  PB_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * (size_t)n, s));
  PB_CUDA(cudaMemsetAsync(totals.p, 0, sizeof(unsigned long long) * 4, s));
  sn_child_count_kernel<<<nblk(n), 256, 0, s>>>(parent, n, cnt.p);
  sn_flag_kernel<<<nblk(n), 256, 0, s>>>(parent, colcount, cnt.p, n, flag.p);
  h.stats.kernels_launched += 2 + exclusive_scan<0>(flag.p, pos.p, n, h.scan_ws.p, s);
  sn_scatter_kernel<<<nblk(n), 256, 0, s>>>(flag.p, pos.p, n, d_super, d_smap);
  sn_parent_kernel<<<std::min(nblk(n), 4u * (unsigned)kNumSMsB200), 256, 0, s>>>(parent, colcount, flag.p, pos.p, d_super, n,
                                                                                 d_spar, totals.p);
  h.stats.kernels_launched += 2;
  PB_CUDA(cudaGetLastError());
  unsigned long long t[3] = {0, 0, 0};
  PB_CUDA(cudaMemcpyAsync(t, totals.p, sizeof(t), cudaMemcpyDeviceToHost, s));
  PB_CUDA(cudaStreamSynchronize(s));
  const int ns = (int)t[0];
  if (!on_device) {
    if (super_out) PB_CUDA(cudaMemcpyAsync(super_out, d_super, sizeof(int) * ((size_t)ns + 1), cudaMemcpyDeviceToHost, s));
    if (supermap_out) PB_CUDA(cudaMemcpyAsync(supermap_out, d_smap, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (sparent_out) PB_CUDA(cudaMemcpyAsync(sparent_out, d_spar, sizeof(int) * (size_t)ns, cudaMemcpyDeviceToHost, s));
    h.stats.d2h_bytes += (long long)sizeof(int) * ((long long)n + 2ll * ns + 1);
    PB_CUDA(cudaStreamSynchronize(s));
  }
  if (nsuper) *nsuper = ns;
  if (ssize) *ssize = (long long)t[1];
  if (xsize) *xsize = (long long)t[2];
  tm.stop();
}

}  // namespace pb
