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

constexpr int K1_WORKERS = 256;                 // threads that process columns / samples
constexpr int K1_THREADS = K1_WORKERS + 32;     // + one warp that only resolves the look-back
constexpr int K1_ITEMS = 16;
constexpr int K1_TILE = K1_WORKERS * K1_ITEMS;  // samples per tile
constexpr int K1_WORDS = K1_TILE / 32;          // bitmap words per tile
constexpr int K1_WTILE = K1_TILE / (K1_WORKERS / 32);  // samples per worker warp
constexpr int K1_SHORT = 32;                    // up to this many mirror probes are done by one thread
constexpr int K1_CCAP = K1_TILE + 2;            // column-window capacity

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

__device__ __forceinline__ int slot_row(int x) { return x; }

// first position in [lo, hi) of column c's samples whose mesh row is >= target; the column is
// read from the staged tile when it lies wholly inside it, else from global memory
template <bool DIM1>
__device__ __forceinline__ int column_lower_bound(const int *__restrict__ Ap,
                                                  const int *__restrict__ Ai, int dim, int c, int vs,
                                                  int ve, int target, int q0, int q1,
                                                  const volatile int *s_val, int *found_row) {
  int lo = 0, hi = ve - vs;
  const int n = hi;
  int row = -1;
  if (vs >= q0 && ve <= q1) {
    const volatile int *a = s_val + (vs - q0);
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (slot_row(a[mid]) < target) lo = mid + 1; else hi = mid;
    }
    if (lo < n) row = slot_row(a[lo]);
  } else if (DIM1) {
    const int *a = Ai + vs;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(a + mid) < target) lo = mid + 1; else hi = mid;
    }
    if (lo < n) row = __ldg(a + lo);
  } else {
    const int *a = Ai + __ldg(Ap + (long long)c * dim);
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(a + (long long)mid * dim) / dim < target) lo = mid + 1; else hi = mid;
    }
    if (lo < n) row = __ldg(a + (long long)lo * dim) / dim;
  }
  *found_row = row;
  return vs + lo;
}

// One tile = K1_TILE consecutive samples.
//   A   stage the samples in shared memory (DIM1: straight 128-bit copy of Ai)
//   B   column side, one worker thread per column of the tile's column window: binary search for
//       the diagonal (columns are proven non-decreasing in C1, else the result is discarded),
//       mark its slot dropped, set the column's head bit, add (upper - lower) of the in-tile part
//       to the symmetry count, and probe the mirror (r,c) of every in-tile (c,r), r > c
//   C1  sample side, warp-contiguous: order / duplicate / range proof against the previous sample
//       (shuffle + head bit), dropped bitmap by ballot
//   C2  prefix of dropped counts per bitmap word (one warp); the extra warp then resolves the
//       tile's global offset by decoupled look-back while the workers compact:
//       rank(j) = j - dropped_before(j)
//   E   coalesced stores of Mi; Mp[c] = base + rank(V[c]) for the column starts this tile owns
// Duplicates inside a column, disorder, asymmetry, or a column window wider than K1_CCAP raise a
// flag: the host then runs the general path instead.
template <bool DIM1>
__global__ void __launch_bounds__(K1_THREADS)
build_filter_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai,
                    const int *__restrict__ V, int dim, int M, int Q,
                    const int *__restrict__ tile_c, int num_tiles, int *__restrict__ Mp,
                    int *__restrict__ Mi, BuildStatus *__restrict__ st,
                    unsigned long long *__restrict__ desc, int check_mirror) {
  __shared__ __align__(16) int s_val[K1_TILE];  // staged samples
  __shared__ __align__(16) int s_out[K1_TILE];  // compacted tile
  __shared__ unsigned s_head[K1_WORDS];         // bit j: sample j is the first of its column
  __shared__ unsigned s_drop[K1_WORDS];         // bit j: sample j does not survive
  __shared__ int s_dpre[K1_WORDS + 1];          // dropped samples before word w
  __shared__ int s_long[2 * (K1_TILE / K1_SHORT + 2)];
  __shared__ int s_nlong, s_diff, s_prev0;
  __shared__ long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool worker = tid < K1_WORKERS;
  // tiles are taken in blockIdx order: the hardware dispatches CTAs of a 1-D grid in that order,
  // so every predecessor of a resident tile is resident or finished (the look-back cannot starve)
  const int tile = blockIdx.x;
  const bool last = tile == num_tiles - 1;
  const int q0 = tile * K1_TILE;
  const int nq = Q - q0 < K1_TILE ? Q - q0 : K1_TILE;
  const int q1 = q0 + nq;
  const int cA = tile_c[tile];
  const int cEnd = last ? M : tile_c[tile + 1];
  const int w0 = tile == 0 ? 0 : cA;  // tile 0 also owns the empty columns in front of sample 0
  int wn = cEnd - w0 + 1;             // columns in the window
  if (tid == 0) { s_nlong = 0; s_diff = 0; }
  for (int w = tid; w < K1_WORDS; w += K1_THREADS) {
    s_head[w] = 0u;
    // slots past the end of the last tile count as dropped
    const int rem = nq - w * 32;
    s_drop[w] = rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : ~((1u << rem) - 1u));
  }

  // ---- A
  if (DIM1) {
    if (nq == K1_TILE && ((reinterpret_cast<uintptr_t>(Ai) & 15) == 0)) {
      const int4 *src = reinterpret_cast<const int4 *>(Ai + q0);
      for (int k = tid; k < K1_TILE / 4; k += K1_THREADS) reinterpret_cast<int4 *>(s_val)[k] = ld_stream4(src + k);
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
  if (tid == 0) {
    // sample just before the tile, if the tile starts in the middle of column cA
    const int vsA = __ldg(V + cA);
    int pv = -1;
    if (vsA < q0)
      pv = DIM1 ? __ldg(Ai + q0 - 1)
                : __ldg(Ai + (long long)__ldg(Ap + (long long)cA * dim) + (long long)(q0 - 1 - vsA) * dim) / dim;
    s_prev0 = pv;
  }
  __syncthreads();
  if (wn > K1_CCAP) {
    // a run of > K1_TILE empty columns inside one tile: leave it to the general path, but keep
    // the scan chain alive
    if (tid == 0) atomicOr(&st->flags, BUILD_WIDE);
    wn = 0;
  }

  const volatile int *vval = s_val;
  int diff = 0;
  unsigned flags = 0;
  // ---- B
  if (worker) {
    for (int lc = tid; lc < wn; lc += K1_WORKERS) {
      const int c = w0 + lc;
      if (c >= M) continue;
      const int vs = __ldg(V + c), ve = __ldg(V + c + 1);
      const int a = vs > q0 ? vs : q0, b = ve < q1 ? ve : q1;
      if (b <= a) continue;
      if (vs >= q0) atomicOr(&s_head[(vs - q0) >> 5], 1u << ((vs - q0) & 31));
      int row;
      const int dpos = column_lower_bound<DIM1>(Ap, Ai, dim, c, vs, ve, c, q0, q1, vval, &row);
      const int has_diag = row == c;
      if (has_diag && dpos >= a && dpos < b) atomicOr(&s_drop[(dpos - q0) >> 5], 1u << ((dpos - q0) & 31));
      int lo_end = dpos < a ? a : (dpos > b ? b : dpos);          // in-tile samples below the diagonal
      int up_beg = dpos + has_diag;
      up_beg = up_beg < a ? a : (up_beg > b ? b : up_beg);        // in-tile samples above it
      diff += (b - up_beg) - (lo_end - a);
      if (check_mirror && b > up_beg) {
        if (b - up_beg > K1_SHORT) {
          const int k = atomicAdd(&s_nlong, 1);
          s_long[2 * k] = c;
          s_long[2 * k + 1] = up_beg;
        } else {
          for (int p = up_beg; p < b; p++) {
            const int r = slot_row(vval[p - q0]);
            if ((unsigned)r < (unsigned)M) {
              int rr;
              const int rs = __ldg(V + r), re = __ldg(V + r + 1);
              column_lower_bound<DIM1>(Ap, Ai, dim, r, rs, re, c, q0, q1, vval, &rr);
              if (rr != c) flags |= BUILD_ASYM;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int nlong = s_nlong;
  if (worker) {
    // columns with many in-tile upper entries: one warp probes them
    for (int li = warp; li < nlong; li += K1_WORKERS / 32) {
      const int c = s_long[2 * li], up_beg = s_long[2 * li + 1];
      const int ve = __ldg(V + c + 1);
      const int b = ve < q1 ? ve : q1;
      for (int p = up_beg + lane; p < b; p += 32) {
        const int r = slot_row(vval[p - q0]);
        if ((unsigned)r < (unsigned)M) {
          int rr;
          const int rs = __ldg(V + r), re = __ldg(V + r + 1);
          column_lower_bound<DIM1>(Ap, Ai, dim, r, rs, re, c, q0, q1, vval, &rr);
          if (rr != c) flags |= BUILD_ASYM;
        }
      }
    }
    // ---- C1
    int carry = s_prev0;  // row of the sample before this warp's first sample
    if (warp > 0) carry = vval[warp * K1_WTILE - 1];
#pragma unroll 4
    for (int k = 0; k < K1_WTILE / 32; k++) {
      const int j = warp * K1_WTILE + k * 32 + lane;
      const int row = j < nq ? vval[j] : 0;
      int prev = __shfl_up_sync(0xffffffffu, row, 1);
      if (lane == 0) prev = carry;
      carry = __shfl_sync(0xffffffffu, row, 31);
      const unsigned hw = s_head[(warp * K1_WTILE >> 5) + k];
      const bool head = (hw >> lane) & 1u;
      if (j < nq) {
        if (!head && row <= prev) flags |= (row < prev) ? BUILD_UNSORTED : BUILD_DUP;
        if ((unsigned)row >= (unsigned)M) flags |= BUILD_RANGE;
      }
    }
  }
  // block sum of (upper - lower), flags
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
  if (lane == 0 && diff) atomicAdd(&s_diff, diff);
  if (flags) atomicOr(&st->flags, flags);
  __syncthreads();
  // ---- C2: prefix of dropped counts per word (warp 0: K1_WORDS/32 words per lane)
  if (warp == 0) {
    constexpr int PER = K1_WORDS / 32;
    int cnt[PER], sum = 0;
#pragma unroll
    for (int u = 0; u < PER; u++) { cnt[u] = __popc(s_drop[lane * PER + u]); sum += cnt[u]; }
    int run = warp_incl_scan(sum, lane) - sum;
#pragma unroll
    for (int u = 0; u < PER; u++) { s_dpre[lane * PER + u] = run; run += cnt[u]; }
    if (lane == 31) s_dpre[K1_WORDS] = run;
  }
  __syncthreads();
  // slots past nq were pre-marked dropped, so kept = K1_TILE - dropped
  const int total = K1_TILE - s_dpre[K1_WORDS];

  if (!worker) {
    // ---- look-back warp: global offset of the tile, overlapped with the compaction
    const int d = s_diff;
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
  } else {
#pragma unroll 4
    for (int k = 0; k < K1_WTILE / 32; k++) {
      const int w = (warp * K1_WTILE >> 5) + k;
      const int j = w * 32 + lane;
      const unsigned dm = s_drop[w];
      if (!((dm >> lane) & 1u)) s_out[j - s_dpre[w] - __popc(dm & ((1u << lane) - 1u))] = s_val[j];
    }
  }
  __syncthreads();

  // ---- E
  const int base = (int)(s_base & 0x7fffffff);
  for (int i = tid; i < total; i += K1_THREADS) st_stream(Mi + base + i, s_out[i]);
  // column starts owned by this tile: q0 < V[c] <= q1 (tile 0 also owns V[c] == 0), so runs of
  // empty columns on a tile boundary and the closing Mp[M] are written exactly once
  for (int lc = tid; lc < wn; lc += K1_THREADS) {
    const int c = w0 + lc;
    const int vs = __ldg(V + c);
    if ((vs > q0 && vs <= q1) || (tile == 0 && vs == 0)) {
      const int pos = vs - q0;
      int rank = total;
      if (pos < K1_TILE && vs < q1) rank = pos - s_dpre[pos >> 5] - __popc(s_drop[pos >> 5] & ((1u << (pos & 31)) - 1u));
      Mp[c] = base + rank;
    }
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
  if (a.ev_a) cudaEventRecord(a.ev_a, s);
  if (a.dim == 1)
    build_filter_kernel<true><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, 1, a.M, a.Q, a.tile_c,
                                                               num_tiles, a.Mp, a.Mi, a.status,
                                                               a.desc, a.check_mirror);
  else
    build_filter_kernel<false><<<num_tiles, K1_THREADS, 0, s>>>(a.Ap, a.Ai, V, a.dim, a.M, a.Q,
                                                                a.tile_c, num_tiles, a.Mp, a.Mi,
                                                                a.status, a.desc, a.check_mirror);
  if (a.ev_b) cudaEventRecord(a.ev_b, s);
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
