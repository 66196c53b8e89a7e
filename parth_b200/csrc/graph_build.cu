// parth_b200/csrc/graph_build.cu -- stage (a): CSC matrix -> mesh graph on the device.
//
// Replaces ParthAPI::computeMeshFromMatrix (reference src/Parth/ParthAPI.cpp:366-418): sample
// block-column c = 0, dim, 2dim, ... at entries Ap[c], Ap[c]+dim, ...; mesh_r = Ai[p]/dim; emit
// (c,r) and (r,c) for r != c; sort; unique; CSR with strictly ascending rows.
//
// The reference materialises every pair and sorts 2*nnz tuples.  Here:
//   * "filter" path (build_filter_kernel): ONE streaming pass.  If every sampled column is
//     non-decreasing and the sampled pattern is structurally symmetric, the sorted-unique pair
//     list restricted to row r IS column r's off-diagonal samples, so the output is a stream
//     compaction of the input.  The kernel proves both properties while it compacts (ordering
//     against the previous sample; symmetry by finding the mirror (r,c) of every kept (c,r),
//     r > c, plus an upper/lower count match, which turns the injection into a bijection) and
//     raises a flag if the proof fails.  Work is tiled by samples (2048 per CTA), so every thread
//     owns the same number of samples whatever the degree distribution; output offsets come from
//     a decoupled look-back chained scan, so input is read once and output written once:
//     algorithmic bytes = 4*[(N+1) + nnzA/dim + (M+1) + nnzM].
//   * general path (gen_* kernels): count -> scan -> scatter -> per-row bitonic sort + unique ->
//     scan -> gather.  Any triangle, unsorted columns, duplicates.  Same result as the filter
//     path wherever both apply (tests/test_gpu_build.py).
#include <climits>

#include "build.cuh"
#include "scan.cuh"

namespace pb {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_ITEMS = 16;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;  // samples per tile
constexpr int K1_WTILE = K1_TILE / K1_WARPS;    // samples per warp (warp-contiguous layout)
constexpr int K1_SHORT = 32;                    // columns up to this length are filled by one thread

// tile_c[t] = column that holds sample t*K1_TILE (last c < M with V[c] <= q)
__global__ void build_partition_kernel(const int *__restrict__ V, int M, int num_tiles,
                                       int *__restrict__ tile_c) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  const long long q = (long long)t * K1_TILE;
  int lo = 0, hi = M;  // first c with V[c] > q
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((long long)V[mid] <= q) lo = mid + 1; else hi = mid;
  }
  tile_c[t] = lo - 1;
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

// Is `target` among the sampled mesh rows of block-column `col`?  [vs, ve) is the column's
// sample range.  Looks in the tile's staged samples when the whole column is staged (staged
// slots hold the mesh row, or ~row once the owning column has classified the slot as dropped).
template <bool DIM1>
__device__ __forceinline__ bool mirror_present(const int *__restrict__ Ap,
                                               const int *__restrict__ Ai, int dim, int col,
                                               int target, int vs, int ve, int q0, int q1,
                                               const volatile int *s_val) {
  int lo = 0, hi = ve - vs;
  const int n = hi;
  if (vs >= q0 && ve <= q1) {
    const volatile int *a = s_val + (vs - q0);
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      int x = a[mid];
      x = x < 0 ? ~x : x;
      if (x < target) lo = mid + 1; else hi = mid;
    }
    if (lo >= n) return false;
    int x = a[lo];
    return (x < 0 ? ~x : x) == target;
  }
  if (DIM1) {
    const int *a = Ai + vs;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(a + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(a + lo) == target;
  }
  const int *a = Ai + __ldg(Ap + (long long)col * dim);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + (long long)mid * dim) / dim < target) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(a + (long long)lo * dim) / dim == target;
}

// classification of one sample of column c (Integrator-independent; ParthAPI.cpp:374-381 plus the
// sort/unique that follows): returns true if the sample survives into the mesh row
template <bool DIM1>
__device__ __forceinline__ bool classify(int c, int v, bool has_prev, int prev, int M,
                                         const int *__restrict__ Ap, const int *__restrict__ Ai,
                                         const int *__restrict__ V, int dim, int q0, int q1,
                                         const volatile int *s_val, int check_mirror,
                                         unsigned &flags, int &diff) {
  bool keep = v != c;
  if (has_prev) {
    if (v < prev) flags |= BUILD_UNSORTED;
    else if (v == prev) keep = false;
  }
  if ((unsigned)v >= (unsigned)M) { flags |= BUILD_RANGE; return false; }
  if (keep) {
    if (v > c) {
      diff++;
      if (check_mirror) {
        const int vs = __ldg(V + v), ve = __ldg(V + v + 1);
        if (!mirror_present<DIM1>(Ap, Ai, dim, v, c, vs, ve, q0, q1, s_val)) flags |= BUILD_ASYM;
      }
    } else {
      diff--;
    }
  }
  return keep;
}

// One tile = K1_TILE consecutive samples.  Phases:
//   A  stage the samples in shared memory (DIM1: straight 128-bit copy of Ai)
//   B  one thread per column of the tile's column window walks that column's staged samples:
//      diagonal / duplicate / order / mirror checks with the previous sample in a register; a
//      dropped sample is overwritten by its complement (~v < 0).  Columns longer than K1_SHORT
//      samples are walked by a whole warp.
//   C  warp-contiguous pass over the samples: ballot ranks -> compaction into shared memory, and
//      the exclusive rank of every slot (needed for Mp at column starts).  Warp 0 resolves the
//      tile's global output offset by decoupled look-back meanwhile.
//   D  coalesced stores of Mi, and of Mp for the column starts this tile owns
template <bool DIM1>
__global__ void __launch_bounds__(K1_THREADS)
build_filter_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai,
                    const int *__restrict__ V, int dim, int M, int Q,
                    const int *__restrict__ tile_c, int num_tiles, int *__restrict__ Mp,
                    int *__restrict__ Mi, BuildStatus *__restrict__ st,
                    unsigned long long *__restrict__ desc, int check_mirror) {
  __shared__ __align__(16) int s_val[K1_TILE];  // staged samples; exclusive ranks after phase C
  __shared__ __align__(16) int s_out[K1_TILE];  // compacted output
  __shared__ int s_wsum[K1_WARPS], s_wdiff[K1_WARPS];
  __shared__ int s_long[K1_TILE / K1_SHORT + 1];
  __shared__ int s_nlong;
  __shared__ long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles are taken in blockIdx order: the hardware dispatches CTAs of a 1-D grid in that order,
  // so every predecessor of a resident tile is resident or finished (the look-back cannot starve)
  const int tile = blockIdx.x;
  const bool last = tile == num_tiles - 1;
  const int q0 = tile * K1_TILE;
  const int nq = Q - q0 < K1_TILE ? Q - q0 : K1_TILE;
  const int q1 = q0 + nq;
  const int cA = tile_c[tile];
  const int cEnd = last ? M : tile_c[tile + 1];
  if (tid == 0) s_nlong = 0;

  // ---- A
  if (DIM1) {
    if (nq == K1_TILE && ((reinterpret_cast<uintptr_t>(Ai) & 15) == 0)) {
      const int4 *src = reinterpret_cast<const int4 *>(Ai + q0);
#pragma unroll
      for (int k = 0; k < K1_ITEMS / 4; k++)
        reinterpret_cast<int4 *>(s_val)[tid + k * K1_THREADS] = ld_stream4(src + tid + k * K1_THREADS);
    } else {
      for (int j = tid; j < nq; j += K1_THREADS) s_val[j] = ld_stream(Ai + q0 + j);
    }
  } else {
    // dim > 1: each column's thread gathers its own strided samples
    for (int c = cA + tid; c <= cEnd && c < M; c += K1_THREADS) {
      const int vs = __ldg(V + c), ve = __ldg(V + c + 1);
      const int a = vs > q0 ? vs : q0, b = ve < q1 ? ve : q1;
      const int *col = Ai + __ldg(Ap + (long long)c * dim);
      for (int p = a; p < b; p++) s_val[p - q0] = __ldg(col + (long long)(p - vs) * dim) / dim;
    }
  }
  __syncthreads();

  // ---- B
  int diff = 0;
  unsigned flags = 0;
  const volatile int *vval = s_val;
  for (int c = cA + tid; c <= cEnd && c < M; c += K1_THREADS) {
    const int vs = __ldg(V + c), ve = __ldg(V + c + 1);
    const int a = vs > q0 ? vs : q0, b = ve < q1 ? ve : q1;
    if (b <= a) continue;
    if (b - a > K1_SHORT) { s_long[atomicAdd(&s_nlong, 1)] = c; continue; }
    bool has_prev = a > vs;  // only when the tile starts in the middle of this column
    int prev = 0;
    if (has_prev)
      prev = DIM1 ? __ldg(Ai + a - 1)
                  : __ldg(Ai + (long long)__ldg(Ap + (long long)c * dim) + (long long)(a - 1 - vs) * dim) / dim;
    for (int p = a; p < b; p++) {
      const int v = vval[p - q0];
      if (!classify<DIM1>(c, v, has_prev, prev, M, Ap, Ai, V, dim, q0, q1, vval, check_mirror, flags, diff))
        s_val[p - q0] = ~v;
      prev = v;
      has_prev = true;
    }
  }
  __syncthreads();
  for (int li = warp; li < s_nlong; li += K1_WARPS) {
    const int c = s_long[li];
    const int vs = __ldg(V + c), ve = __ldg(V + c + 1);
    const int a = vs > q0 ? vs : q0, b = ve < q1 ? ve : q1;
    int carry = 0;
    bool carry_ok = a > vs;
    if (carry_ok)
      carry = DIM1 ? __ldg(Ai + a - 1)
                   : __ldg(Ai + (long long)__ldg(Ap + (long long)c * dim) + (long long)(a - 1 - vs) * dim) / dim;
    for (int p0 = a; p0 < b; p0 += 32) {
      const int p = p0 + lane;
      const int v = p < b ? vval[p - q0] : 0;
      int prev = __shfl_up_sync(0xffffffffu, v, 1);
      bool has_prev = true;
      if (lane == 0) { prev = carry; has_prev = carry_ok; }
      if (p < b &&
          !classify<DIM1>(c, v, has_prev, prev, M, Ap, Ai, V, dim, q0, q1, vval, check_mirror, flags, diff))
        s_val[p - q0] = ~v;
      carry = __shfl_sync(0xffffffffu, v, 31);
      carry_ok = true;
    }
  }
  if (flags) atomicOr(&st->flags, flags);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
  __syncthreads();

  // ---- C: per-warp kept counts, then ranks
  unsigned keepbits = 0;
  int wcount = 0;
#pragma unroll
  for (int k = 0; k < K1_ITEMS; k++) {
    const int j = warp * K1_WTILE + k * 32 + lane;
    const bool keep = j < nq && s_val[j] >= 0;
    keepbits |= (keep ? 1u : 0u) << k;
    wcount += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (lane == 0) { s_wsum[warp] = wcount; s_wdiff[warp] = diff; }
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < K1_WARPS; w++) {
    const int x = s_wsum[w];
    if (w < warp) woff += x;
    total += x;
  }
  if (warp == 0) {
    int d = lane < K1_WARPS ? s_wdiff[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    // pack: low 31 bits kept count, next 31 bits (upper - lower) modulo 2^31
    const long long agg = (long long)total | ((long long)((unsigned)d & 0x7fffffffu) << 31);
    const long long b = tile_lookback(desc, tile, agg);
    if (lane == 0) {
      s_base = b;
      if (last) {
        const long long incl = (b + agg) & (long long)kTileVal;
        st->nnz_out = (int)(incl & 0x7fffffff);
        st->diff_mod = (unsigned)((incl >> 31) & 0x7fffffff);
      }
    }
  }
  {
    int run = woff;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++) {
      const int j = warp * K1_WTILE + k * 32 + lane;
      const bool keep = (keepbits >> k) & 1u;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int pre = run + __popc(m & ((1u << lane) - 1u));
      if (keep) s_out[pre] = s_val[j];
      s_val[j] = pre;  // the slot now holds its exclusive rank (its value has been consumed)
      run += __popc(m);
    }
  }
  __syncthreads();

  // ---- D
  const int base = (int)(s_base & 0x7fffffff);
  for (int i = tid; i < total; i += K1_THREADS) st_stream(Mi + base + i, s_out[i]);
  // column starts owned by this tile: q0 < V[c] <= q1 (tile 0 also owns V[c] == 0), so runs of
  // empty columns on a tile boundary and the closing Mp[M] are written exactly once
  for (int c = (tile == 0 ? 0 : cA) + tid; c <= cEnd; c += K1_THREADS) {
    const int vs = __ldg(V + c);
    if ((vs > q0 && vs <= q1) || (tile == 0 && vs == 0)) Mp[c] = base + (vs < q1 ? s_val[vs - q0] : total);
  }
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
int build_filter_tiles(int M, int Q) { (void)M; return (Q + K1_TILE - 1) / K1_TILE; }

int build_filter_launch(const BuildArgs &a, cudaStream_t s) {
  int launches = 0;
  const int *V = a.dim == 1 ? a.Ap : a.V;
  const int num_tiles = build_filter_tiles(a.M, a.Q);
  PB_CUDA(cudaMemsetAsync(a.status, 0, sizeof(BuildStatus), s));
  if (num_tiles == 0) {  // no samples at all: empty graph
    PB_CUDA(cudaMemsetAsync(a.Mp, 0, sizeof(int) * ((size_t)a.M + 1), s));
    return 0;
  }
  PB_CUDA(cudaMemsetAsync(a.desc, 0, sizeof(unsigned long long) * (size_t)num_tiles, s));
  build_partition_kernel<<<(num_tiles + 255) / 256, 256, 0, s>>>(V, a.M, num_tiles, a.tile_c);
  launches++;
  if (a.dim == 1)
    build_filter_kernel<true><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, 1, a.M, a.Q, a.tile_c,
                                                               num_tiles, a.Mp, a.Mi, a.status,
                                                               a.desc, a.check_mirror);
  else
    build_filter_kernel<false><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, a.dim, a.M, a.Q,
                                                                a.tile_c, num_tiles, a.Mp, a.Mi,
                                                                a.status, a.desc, a.check_mirror);
  launches++;
  PB_CUDA(cudaGetLastError());
  return launches;
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
