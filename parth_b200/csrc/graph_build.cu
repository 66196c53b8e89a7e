// parth_b200/csrc/graph_build.cu -- stage (a): CSC matrix -> mesh graph on the device.
//
// Replaces ParthAPI::computeMeshFromMatrix (reference src/Parth/ParthAPI.cpp:366-418): sample
// block-column c = 0, dim, 2dim, ... at entries Ap[c], Ap[c]+dim, ...; mesh_r = Ai[p]/dim; emit
// (c,r) and (r,c) for r != c; sort; unique; CSR with strictly ascending rows.
//
// The reference materialises every pair and sorts 2*nnz tuples.  Here:
//   * "filter" path (build_sym_kernel): ONE streaming pass.  If every sampled column is
//     non-decreasing and the sampled pattern is structurally symmetric, the sorted-unique pair
//     list restricted to row r IS column r's off-diagonal samples, so the output is a stream
//     compaction of the input.  The kernel proves both properties while it compacts (ordering
//     against the previous sample; symmetry by finding the mirror (r,c) of every kept (c,r),
//     r > c, plus an upper/lower count match, which turns the injection into a bijection) and
//     raises a flag if the proof fails.  Work is split by merge-path over (columns, samples) so
//     every thread owns the same number of path items whatever the degree distribution; output
//     offsets come from a decoupled look-back chained scan, so input is read once and output
//     written once: algorithmic bytes = 4*[(N+1) + nnzA/dim + (M+1) + nnzM].
//   * general path (gen_* kernels): count -> scan -> scatter -> per-row bitonic sort + unique ->
//     scan -> gather.  Any triangle, unsorted columns, duplicates.  Same result as the filter
//     path wherever both apply (tests/test_gpu_build.py).
#include <climits>

#include "build.cuh"
#include "scan.cuh"

namespace pb {

constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 9;  // odd: the per-thread walk over shared memory is bank-conflict free
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;

// number of row-ends (V[c+1]) consumed before path diagonal `diag`
__device__ __forceinline__ int merge_path(long long diag, const int *row_end, int M, long long Q,
                                          long long q_base) {
  long long lo = diag > Q ? diag - Q : 0;
  long long hi = diag < M ? diag : M;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if ((long long)row_end[mid] <= q_base + diag - 1 - mid) lo = mid + 1; else hi = mid;
  }
  return (int)lo;
}

__global__ void build_partition_kernel(const int *__restrict__ V, int M, int Q, int num_tiles,
                                       int *__restrict__ tile_c) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > num_tiles) return;
  long long path = (long long)M + Q;
  long long d = (long long)t * K1_TILE;
  if (d > path) d = path;
  tile_c[t] = merge_path(d, V + 1, M, Q, 0);
}

// samples per block-column for dim > 1 (ParthAPI.cpp:373: r_ptr += dim)
__global__ void build_sample_count_kernel(const int *__restrict__ Ap, int M, int dim,
                                          int *__restrict__ ns) {
  int mc = blockIdx.x * blockDim.x + threadIdx.x;
  if (mc >= M) return;
  long long c = (long long)mc * dim;
  int len = Ap[c + 1] - Ap[c];
  ns[mc] = len > 0 ? (len + dim - 1) / dim : 0;
}

// is `target` among the sampled mesh rows of block-column `col`?
template <bool DIM1>
__device__ __forceinline__ bool mirror_in_global(const int *__restrict__ Ap,
                                                 const int *__restrict__ Ai,
                                                 const int *__restrict__ V, int dim, int col,
                                                 int target) {
  int lo = 0, hi, base;
  if (DIM1) {
    base = __ldg(Ap + col);
    hi = __ldg(Ap + col + 1) - base;
  } else {
    base = __ldg(Ap + (long long)col * dim);
    hi = __ldg(V + col + 1) - __ldg(V + col);
  }
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    int v = DIM1 ? __ldg(Ai + base + mid) : __ldg(Ai + base + (long long)mid * dim) / dim;
    if (v < target) lo = mid + 1; else hi = mid;
  }
  if (lo >= (DIM1 ? __ldg(Ap + col + 1) - base : __ldg(V + col + 1) - __ldg(V + col))) return false;
  int v = DIM1 ? __ldg(Ai + base + lo) : __ldg(Ai + base + (long long)lo * dim) / dim;
  return v == target;
}

template <bool DIM1>
__global__ void __launch_bounds__(K1_THREADS)
build_sym_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai, const int *__restrict__ V,
                 int dim, int M, int Q, const int *__restrict__ tile_c, int num_tiles,
                 int *__restrict__ Mp, int *__restrict__ Mi, BuildStatus *__restrict__ st,
                 unsigned long long *__restrict__ desc, int check_mirror) {
  __shared__ int s_vs[K1_TILE + 2];
  __shared__ int s_val[K1_TILE];
  __shared__ int s_out[K1_TILE];
  __shared__ int s_rank[K1_TILE + 1];
  __shared__ int s_scan[33];
  __shared__ int s_diff[K1_THREADS / 32];
  __shared__ int s_tile;
  __shared__ long long s_base;

  const int tid = threadIdx.x;
  if (tid == 0) s_tile = (int)atomicAdd(&st->tile_counter, 1ull);
  __syncthreads();
  const int tile = s_tile;
  const long long path = (long long)M + Q;
  const long long d0 = (long long)tile * K1_TILE;
  const long long d1 = d0 + K1_TILE < path ? d0 + K1_TILE : path;
  const int c0 = tile_c[tile], c1 = tile_c[tile + 1];
  const int q0 = (int)(d0 - c0), q1 = (int)(d1 - c1);
  const int nc = c1 - c0;  // columns that end inside this tile
  const int nq = q1 - q0;  // samples inside this tile
  const int total_items = (int)(d1 - d0);

  for (int i = tid; i <= nc + 1; i += K1_THREADS) {
    int c = c0 + i;
    s_vs[i] = c <= M ? V[c] : INT_MAX;
  }
  if (DIM1) {
    for (int j = tid; j < nq; j += K1_THREADS) s_val[j] = ld_stream(Ai + q0 + j);
  }
  __syncthreads();

  // thread-level merge-path split inside the tile
  const int diag = tid * K1_ITEMS < total_items ? tid * K1_ITEMS : total_items;
  int ci;
  {
    int lo = diag > nq ? diag - nq : 0;
    int hi = diag < nc ? diag : nc;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (s_vs[mid + 1] <= q0 + diag - 1 - mid) lo = mid + 1; else hi = mid;
    }
    ci = lo;
  }
  const int qi = diag - ci;

  // ---- pass 1: classify samples, prove ordering and symmetry
  unsigned keepmask = 0;
  int kept = 0, diff = 0;
  unsigned flags = 0;
  {
    int c = ci, q = qi;
#pragma unroll
    for (int it = 0; it < K1_ITEMS; it++) {
      if (diag + it >= total_items) break;
      const int rend = s_vs[c + 1];
      const int aq = q0 + q;
      if (aq < rend) {
        const int mc = c0 + c;
        const int rstart = s_vs[c];
        int mr, prev = -1;
        const bool has_prev = aq > rstart;
        if (DIM1) {
          mr = s_val[q];
          if (has_prev) prev = q > 0 ? s_val[q - 1] : __ldg(Ai + aq - 1);
        } else {
          const long long p = (long long)__ldg(Ap + (long long)mc * dim) + (long long)(aq - rstart) * dim;
          mr = __ldg(Ai + p) / dim;
          s_val[q] = mr;
          if (has_prev) prev = __ldg(Ai + p - dim) / dim;
        }
        bool keep = mr != mc;
        if (has_prev) {
          if (mr < prev) flags |= BUILD_UNSORTED;
          else if (mr == prev) keep = false;
        }
        if ((unsigned)mr >= (unsigned)M) { flags |= BUILD_RANGE; keep = false; }
        if (keep) {
          keepmask |= 1u << it;
          kept++;
          if (mr > mc) {
            diff++;
            if (check_mirror) {
              bool ok;
              const int rel = mr - c0;
              if (DIM1 && rel >= 1 && rel < nc) {
                // mirror column lies wholly inside this tile's staged samples
                int lo = s_vs[rel] - q0, hi = s_vs[rel + 1] - q0;
                const int end = hi;
                while (lo < hi) {
                  int mid = (lo + hi) >> 1;
                  if (s_val[mid] < mc) lo = mid + 1; else hi = mid;
                }
                ok = lo < end && s_val[lo] == mc;
              } else {
                ok = mirror_in_global<DIM1>(Ap, Ai, V, dim, mr, mc);
              }
              if (!ok) flags |= BUILD_ASYM;
            }
          } else {
            diff--;
          }
        }
        q++;
      } else {
        c++;
      }
    }
  }

  int total_kept;
  const int excl = block_excl_scan<K1_THREADS>(kept, s_scan, &total_kept);
  // block sum of (upper - lower)
  {
    int d = diff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((tid & 31) == 0) s_diff[tid >> 5] = d;
  }
  if (flags) atomicOr(&st->flags, flags);
  __syncthreads();
  if (tid < 32) {
    int d = tid < K1_THREADS / 32 ? s_diff[tid] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    // pack: low 31 bits kept count, next 31 bits (upper - lower) modulo 2^31
    long long agg = (long long)total_kept | ((long long)((unsigned)d & 0x7fffffffu) << 31);
    long long b = tile_lookback(desc, tile, agg);
    if (tid == 0) {
      s_base = b;
      if (tile == num_tiles - 1) {
        long long incl = (b + agg) & (long long)kTileVal;
        st->nnz_out = (int)(incl & 0x7fffffff);
        st->diff_mod = (unsigned)((incl >> 31) & 0x7fffffff);
      }
    }
  }
  __syncthreads();
  const int base = (int)(s_base & 0x7fffffff);

  // ---- pass 2: compact into shared memory, then coalesced stores
  {
    int c = ci, q = qi, r = excl;
#pragma unroll
    for (int it = 0; it < K1_ITEMS; it++) {
      if (diag + it >= total_items) break;
      if (q0 + q < s_vs[c + 1]) {
        if ((keepmask >> it) & 1u) s_out[r++] = s_val[q];
        q++;
      } else {
        s_rank[c] = r;
        c++;
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < total_kept; j += K1_THREADS) st_stream(Mi + base + j, s_out[j]);
  for (int i = tid; i < nc; i += K1_THREADS) Mp[c0 + 1 + i] = base + s_rank[i];
  if (tile == 0 && tid == 0) Mp[0] = 0;
}

// ------------------------------------------------------------------ general path
// one thread per block-column (fallback path; correctness first)
__global__ void gen_count_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai, int M,
                                 int dim, int *__restrict__ cnt, BuildStatus *st) {
  int mc = blockIdx.x * blockDim.x + threadIdx.x;
  if (mc >= M) return;
  long long c = (long long)mc * dim;
  int own = 0;
  for (int p = Ap[c]; p < Ap[c + 1]; p += dim) {
    int mr = Ai[p] / dim;
    if ((unsigned)mr >= (unsigned)M) { atomicOr(&st->flags, BUILD_RANGE); continue; }
    if (mr != mc) { own++; atomicAdd(cnt + mr, 1); }
  }
  if (own) atomicAdd(cnt + mc, own);
}

__global__ void gen_fill_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai, int M,
                                int dim, const int *__restrict__ ptr, int *__restrict__ fill,
                                int *__restrict__ tmp) {
  int mc = blockIdx.x * blockDim.x + threadIdx.x;
  if (mc >= M) return;
  long long c = (long long)mc * dim;
  for (int p = Ap[c]; p < Ap[c + 1]; p += dim) {
    int mr = Ai[p] / dim;
    if ((unsigned)mr >= (unsigned)M || mr == mc) continue;
    tmp[ptr[mc] + atomicAdd(fill + mc, 1)] = mr;
    tmp[ptr[mr] + atomicAdd(fill + mr, 1)] = mc;
  }
}

// one warp per row: in-place ascending bitonic sort of tmp[ptr[r] .. ptr[r+1]) (virtual +inf
// padding: every compare-exchange puts the minimum at the lower index, so indices >= len never
// move), then in-place unique.  uniq[r] = number of distinct entries.
__global__ void gen_sort_unique_kernel(const int *__restrict__ ptr, int M, int *__restrict__ tmp,
                                       int *__restrict__ uniq) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= M) return;
  int *a = tmp + ptr[r];
  const int len = ptr[r + 1] - ptr[r];
  if (len == 0) { if (lane == 0) uniq[r] = 0; return; }
  int n2 = 1;
  while (n2 < len) n2 <<= 1;
  for (int k = 2; k <= n2; k <<= 1) {
    // first step of each merge: mirror partner inside the block of size k
    for (int i = lane; i < n2; i += 32) {
      int l = i ^ (k - 1);
      if (l > i && l < len) {
        int x = a[i], y = a[l];
        if (x > y) { a[i] = y; a[l] = x; }
      }
    }
    __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = lane; i < n2; i += 32) {
        int l = i ^ j;
        if (l > i && l < len) {
          int x = a[i], y = a[l];
          if (x > y) { a[i] = y; a[l] = x; }
        }
      }
      __syncwarp();
    }
  }
  // unique, chunk by chunk (write index never passes the read index)
  int w = 0;
  for (int b = 0; b < len; b += 32) {
    int i = b + lane;
    int v = i < len ? a[i] : 0;
    int pv = (i > 0 && i < len) ? a[i - 1] : -1;
    bool first = i < len && (i == 0 || v != pv);
    unsigned m = __ballot_sync(0xffffffffu, first);
    __syncwarp();
    if (first) a[w + __popc(m & ((1u << lane) - 1))] = v;
    w += __popc(m);
    __syncwarp();
  }
  if (lane == 0) uniq[r] = w;
}

__global__ void gen_gather_kernel(const int *__restrict__ ptr, const int *__restrict__ Mp, int M,
                                  const int *__restrict__ tmp, int *__restrict__ Mi) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= M) return;
  const int src = ptr[r], dst = Mp[r], len = Mp[r + 1] - Mp[r];
  for (int j = lane; j < len; j += 32) Mi[dst + j] = tmp[src + j];
}

// ------------------------------------------------------------------ host drivers
int build_filter_launch(const BuildArgs &a, cudaStream_t s) {
  int launches = 0;
  const int *V = a.Ap;
  if (a.dim != 1) V = a.V;
  const long long path = (long long)a.M + a.Q;
  const int num_tiles = (int)((path + K1_TILE - 1) / K1_TILE);
  if (num_tiles == 0) return 0;
  PB_CUDA(cudaMemsetAsync(a.status, 0, sizeof(BuildStatus), s));
  PB_CUDA(cudaMemsetAsync(a.desc, 0, sizeof(unsigned long long) * (size_t)num_tiles, s));
  build_partition_kernel<<<(num_tiles + 1 + 255) / 256, 256, 0, s>>>(V, a.M, a.Q, num_tiles, a.tile_c);
  launches++;
  if (a.dim == 1)
    build_sym_kernel<true><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, 1, a.M, a.Q, a.tile_c,
                                                            num_tiles, a.Mp, a.Mi, a.status, a.desc,
                                                            a.check_mirror);
  else
    build_sym_kernel<false><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, a.dim, a.M, a.Q, a.tile_c,
                                                             num_tiles, a.Mp, a.Mi, a.status, a.desc,
                                                             a.check_mirror);
  launches++;
  PB_CUDA(cudaGetLastError());
  return launches;
}

int build_filter_tiles(int M, int Q) {
  return (int)(((long long)M + Q + K1_TILE - 1) / K1_TILE);
}

int build_sample_prefix_launch(const int *Ap, int M, int dim, int *ns_tmp, int *V,
                               unsigned long long *scan_ws, cudaStream_t s) {
  build_sample_count_kernel<<<(M + 255) / 256, 256, 0, s>>>(Ap, M, dim, ns_tmp);
  PB_CUDA(cudaGetLastError());
  return 1 + exclusive_scan<0>(ns_tmp, V, M, scan_ws, s);
}

int build_general_count_launch(const BuildArgs &a, int *cnt, int *ptr, unsigned long long *scan_ws,
                               cudaStream_t s) {
  PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)a.M, s));
  PB_CUDA(cudaMemsetAsync(a.status, 0, sizeof(BuildStatus), s));
  gen_count_kernel<<<(a.M + 255) / 256, 256, 0, s>>>(a.Ap, a.Ai, a.M, a.dim, cnt, a.status);
  PB_CUDA(cudaGetLastError());
  return 1 + exclusive_scan<0>(cnt, ptr, a.M, scan_ws, s);
}

int build_general_finish_launch(const BuildArgs &a, int *cnt, const int *ptr, int *tmp, int *uniq,
                                unsigned long long *scan_ws, cudaStream_t s) {
  int launches = 0;
  PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)a.M, s));
  gen_fill_kernel<<<(a.M + 255) / 256, 256, 0, s>>>(a.Ap, a.Ai, a.M, a.dim, ptr, cnt, tmp);
  launches++;
  const long long threads = (long long)a.M * 32;
  const int blocks = (int)((threads + 255) / 256);
  gen_sort_unique_kernel<<<blocks, 256, 0, s>>>(ptr, a.M, tmp, uniq);
  launches++;
  PB_CUDA(cudaGetLastError());
  launches += exclusive_scan<0>(uniq, a.Mp, a.M, scan_ws, s);
  return launches;
}

int build_general_gather_launch(const BuildArgs &a, const int *ptr, const int *tmp, cudaStream_t s) {
  const long long threads = (long long)a.M * 32;
  const int blocks = (int)((threads + 255) / 256);
  gen_gather_kernel<<<blocks, 256, 0, s>>>(ptr, a.Mp, a.M, tmp, a.Mi);
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
