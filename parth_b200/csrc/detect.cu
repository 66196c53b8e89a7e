// parth_b200/csrc/detect.cu -- stage (b): change detection against the snapshot graph.
//
// Replaces Integrator::computeChangedEdges + computeTheAddedRemovedEdgesForDOF + 
// problematicEdgeCondition (reference src/Parth/Integrator.cpp:23-42,259-338,895-921):
//   row i is "changed" if its degree or any entry differs from the snapshot; a changed row reports
//   every (i, nbr > i) whose separator-tree nodes are neither equal nor ancestor-related, in
//   (row, stored position) order.
//
// Device formulation:
//   diff_rows         streaming comparison.  A CTA owns a block of rows; if the row-pointer shift
//                     Mp[i]-Mp_prev[i] is constant over the block and the two entry ranges are
//                     byte-identical, no row of the block changed (exact, not a heuristic) and the
//                     kernel is a shifted memcmp running at HBM speed: algorithmic bytes
//                     2*4*[(M+1)+nnzM].  Only blocks that fail do the per-row comparison.
//   list_changed_rows ordered compaction of the changed-row bitmap (popc scan outside)
//   count/emit        per changed row (few): problematic-edge count, exclusive scan, ordered emit
#include "scan.cuh"
#include "tree.cuh"

namespace pb {

constexpr int K2_THREADS = 256;
constexpr int K2_ROWS = 2048;  // rows per CTA (multiple of 32: one bitmap word per warp-row-group)

__device__ __forceinline__ bool is_separator_d(int sep, int node) {
  int a = (node - 1) / 2;
  while (a > sep && a > 0) a = (a - 1) / 2;
  return a == sep;
}

// Integrator::problematicEdgeCondition
__device__ __forceinline__ bool problematic_d(const int *__restrict__ dof2node, int a, int b) {
  int na = __ldg(dof2node + a), nb = __ldg(dof2node + b);
  int big = na, small = nb;
  if (small > big) { int t = big; big = small; small = t; }
  return !(small == big || is_separator_d(small, big));
}

__global__ void __launch_bounds__(K2_THREADS)
diff_rows_kernel(const int *__restrict__ Mp, const int *__restrict__ Mi,
                 const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev, int M,
                 unsigned *__restrict__ row_bits, DetectCounters *__restrict__ cnt) {
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31;
  const int r0 = blockIdx.x * K2_ROWS;
  const int r1 = r0 + K2_ROWS < M ? r0 + K2_ROWS : M;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const int p0 = Mp[r0], pp0 = Mp_prev[r0];
  const int delta = p0 - pp0;
  int bad = 0;
  // (1) constant shift of the row pointers over [r0, r1]
  for (int i = r0 + tid; i <= r1; i += K2_THREADS) bad |= (ld_stream(Mp + i) - ld_stream(Mp_prev + i)) != delta;
  // (2) identical entries (lengths match whenever (1) holds; guard with the new length otherwise)
  const int len = Mp[r1] - p0, lenp = Mp_prev[r1] - pp0;
  const int n = len < lenp ? len : lenp;
  const int *a = Mi + p0;
  const int *b = Mi_prev + pp0;
  int j = tid;
  for (; j + 3 * K2_THREADS < n; j += 4 * K2_THREADS) {
    int x0 = ld_stream(a + j), x1 = ld_stream(a + j + K2_THREADS), x2 = ld_stream(a + j + 2 * K2_THREADS),
        x3 = ld_stream(a + j + 3 * K2_THREADS);
    int y0 = ld_stream(b + j), y1 = ld_stream(b + j + K2_THREADS), y2 = ld_stream(b + j + 2 * K2_THREADS),
        y3 = ld_stream(b + j + 3 * K2_THREADS);
    bad |= (x0 != y0) | (x1 != y1) | (x2 != y2) | (x3 != y3);
  }
  for (; j < n; j += K2_THREADS) bad |= ld_stream(a + j) != ld_stream(b + j);
  if (bad) s_bad = 1;
  __syncthreads();
  const int word0 = r0 >> 5;
  const int n_words = (r1 - r0 + 31) >> 5;
  if (!s_bad) {
    for (int w = tid; w < n_words; w += K2_THREADS) row_bits[word0 + w] = 0u;
    return;
  }
  // per-row comparison (Integrator.cpp:269-288), one thread per row, one bitmap word per warp pass
  int changed_here = 0;
  for (int base = r0 + (tid & ~31); base < r1; base += K2_THREADS) {
    const int i = base + lane;
    bool ch = false;
    if (i < r1) {
      const int s = Mp[i], e = Mp[i + 1], sp = Mp_prev[i], ep = Mp_prev[i + 1];
      ch = (e - s) != (ep - sp);
      for (int k = 0; !ch && k < e - s; k++) ch = Mi[s + k] != Mi_prev[sp + k];
    }
    const unsigned m = __ballot_sync(0xffffffffu, ch);
    if (lane == 0) { row_bits[base >> 5] = m; changed_here += __popc(m); }
  }
  if (lane == 0 && changed_here) atomicAdd(&cnt->changed_rows, changed_here);
  if (tid == 0) atomicAdd(&cnt->dirty_tiles, 1);
}

// DOF-map variant (Integrator.cpp:290-337): row cur is compared with the surviving, re-labelled
// neighbours of its old row; added DOFs are always changed.  One thread per row.
__global__ void diff_rows_mapped_kernel(const int *__restrict__ Mp, const int *__restrict__ Mi,
                                        const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev,
                                        const int *__restrict__ new_to_old,
                                        const int *__restrict__ old_to_new, int M,
                                        unsigned *__restrict__ row_bits, DetectCounters *__restrict__ cnt) {
  const int cur = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool ch = false;
  if (cur < M) {
    const int old = new_to_old[cur];
    if (old == -1) ch = true;
    else {
      const int s = Mp[cur], len = Mp[cur + 1] - s;
      int j = 0;
      for (int p = Mp_prev[old]; p < Mp_prev[old + 1]; p++) {
        const int m = __ldg(old_to_new + Mi_prev[p]);
        if (m == -1) continue;
        if (j < len && Mi[s + j] != m) ch = true;
        j++;
      }
      if (j != len) ch = true;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, ch);
  if (lane == 0 && (cur >> 5) < ((M + 31) >> 5)) {
    row_bits[cur >> 5] = m;
    if (m) atomicAdd(&cnt->changed_rows, __popc(m));
  }
}

__global__ void list_changed_rows_kernel(const unsigned *__restrict__ row_bits,
                                         const int *__restrict__ word_prefix, int n_words, int M,
                                         int *__restrict__ rows, int cap) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  unsigned m = row_bits[w];
  int o = word_prefix[w];
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    if (o < cap) rows[o] = (w << 5) + b;
    o++;
  }
}

// computeTheAddedRemovedEdgesForDOF (Integrator.cpp:23-42): one warp per changed row
// n_rows_dev != nullptr: the row count is read on the device (speculative enqueue behind the build,
// the host has not seen it yet); n_rows is then the capacity: slots beyond the count get 0, and
// more rows than slots leave everything 0 (the host falls back to the normal path)
__global__ void count_row_edges_kernel(const int *__restrict__ rows, int n_rows, const int *__restrict__ n_rows_dev,
                                       const int *__restrict__ Mp, const int *__restrict__ Mi,
                                       const int *__restrict__ dof2node, int *__restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_rows) return;
  if (n_rows_dev) {
    const int have = *n_rows_dev;
    if (have > n_rows || k >= have) { if (lane == 0) counts[k] = 0; return; }
  }
  const int i = rows[k];
  int c = 0;
  for (int p = Mp[i] + lane; p < Mp[i + 1]; p += 32) {
    const int nb = Mi[p];
    c += (i < nb) && problematic_d(dof2node, i, nb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[k] = c;
}

__global__ void emit_row_edges_kernel(const int *__restrict__ rows, int n_rows, const int *__restrict__ n_rows_dev,
                                      const int *__restrict__ Mp, const int *__restrict__ Mi,
                                      const int *__restrict__ dof2node,
                                      const int *__restrict__ offsets, int *__restrict__ edges, int cap) {
  const int lane = threadIdx.x & 31;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_rows) return;
  if (n_rows_dev) {
    const int have = *n_rows_dev;
    if (have > n_rows || k >= have) return;
  }
  const int i = rows[k];
  int o = offsets[k];
  const int s = Mp[i], e = Mp[i + 1];
  for (int b = s; b < e; b += 32) {
    const int p = b + lane;
    bool hit = false;
    int nb = 0;
    if (p < e) {
      nb = Mi[p];
      hit = (i < nb) && problematic_d(dof2node, i, nb);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int at = o + __popc(m & ((1u << lane) - 1u));
      if (at < cap) {
        edges[2 * at] = i;
        edges[2 * at + 1] = nb;
      }
    }
    o += __popc(m);
  }
}

// Incremental symmetry proof (stage a).  The snapshot graph was proven symmetric and has the same
// vertex set; rows[] are exactly the rows whose adjacency differs.  Then the new graph is
// symmetric iff, for every changed row j: (1) every entry i of the new row has its mirror j in new
// row i, and (2) every entry i that the row lost is gone from new row i as well.  (An unchanged row
// x keeps its old entries, whose mirrors exist in the old graph: they can only be missing now if
// row y changed and lost x -- which (2) catches from y's side.)  One warp per changed row.
__device__ __forceinline__ bool row_contains(const int *__restrict__ Mi, int s, int e, int v) {
  int lo = s, hi = e;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(Mi + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo < e && __ldg(Mi + lo) == v;
}

// n_rows is read on the device (the host has not seen it yet); more than `cap` rows: nothing is
// checked and the host falls back to the streaming proof
// The two halves of a warp work on one changed row j at the same time: lanes 0-15 walk the new row
// (condition 1, plus the problematic-edge count of stage b), lanes 16-31 the snapshot row (condition
// 2) -- one instruction stream, so the dependent loads of both walks are in flight together.
// zero_out / n_zero: scratch the later kernels of the chain expect cleared (node flags), done by CTA 0.
__global__ void sym_check_rows_kernel(const int *__restrict__ rows, const int *__restrict__ n_rows_dev, int cap,
                                      const int *__restrict__ Mp, const int *__restrict__ Mi,
                                      const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev,
                                      int *__restrict__ asym, const int *__restrict__ dof2node,
                                      int *__restrict__ counts, int spec_cap, int *__restrict__ zero_out, int n_zero) {
  pdl_wait();
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  if (zero_out && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += blockDim.x) zero_out[i] = 0;
  const int n_rows = *n_rows_dev;
  if (n_rows > cap) return;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int *__restrict__ P = half ? Mp_prev : Mp;
  const int *__restrict__ I = half ? Mi_prev : Mi;
  bool bad = false;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_rows; k += warps) {
    const int j = rows[k];
    const int s = Mp[j], e = Mp[j + 1];
    const int s2 = P[j], e2 = P[j + 1];
    const bool want_count = counts != nullptr && k < spec_cap;  // computeTheAddedRemovedEdgesForDOF, fused in
    int c = 0;
    for (int p = s2 + hl; p < e2; p += 16) {
      const int i = I[p];
      const bool mirrored = row_contains(Mi, Mp[i], Mp[i + 1], j);
      if (!half) {
        if (!mirrored) bad = true;                                    // (1) new entry without its mirror
        if (want_count) c += (j < i) && problematic_d(dof2node, j, i);
      } else if (mirrored && !row_contains(Mi, s, e, i)) {
        bad = true;                                                   // (2) lost entry whose mirror stayed
      }
    }
    if (want_count) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) counts[k] = c;
    }
  }
  if (bad) *asym = 1;
}

int launch_sym_check_rows(const int *rows, const int *n_rows_dev, int cap, const int *Mp, const int *Mi,
                          const int *Mp_prev, const int *Mi_prev, int *asym, const int *dof2node, int *counts,
                          int spec_cap, cudaStream_t s, int *zero_out, int n_zero) {
  launch_pdl(sym_check_rows_kernel, dim3(2 * kNumSMsB200), dim3(256), 0, s, rows, n_rows_dev, cap, Mp, Mi, Mp_prev, Mi_prev,
             asym, dof2node, counts, spec_cap, zero_out, n_zero);
  PB_CUDA(cudaGetLastError());
  return 1;
}

// ------------------------------------------------------------------ fused steady state
// The pipeline build kernel (graph_build_pipe.cu) has written the changed-row bitmap and, per tile,
// the number of changed rows (-1: tile left unresolved -> the host falls back to diff_rows_kernel).
// Kernels A1 + A2: ascending list of the changed rows.  A scan over the tile counts (one CTA)
// replaces the scan over the whole bitmap; only the bitmap words of dirty tiles are read.
constexpr int KA_THREADS = 1024;

__device__ __forceinline__ int block_excl_scan_1024(int v, int *s_warp, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  *total = s_warp[31];
  const int base = warp > 0 ? s_warp[warp - 1] : 0;
  __syncthreads();  // s_warp may be reused by the caller
  return base + incl - v;
}

// Kernel A (grid, KR_TILES tiles per CTA): the changed rows of the dirty tiles, ascending.  The
// offset of a tile is the sum of the tile counts before it: a CTA that owns a dirty tile adds up the
// counts in front of its first tile itself (<= num_tiles ints out of L2) and scans its own -- no
// device-wide scan kernel in front.  With unresolved tiles (-1) among the counts the offsets are
// meaningless and the host discards the list -- only the bounds are guarded.
constexpr int KR_TILES = 256;

__global__ void __launch_bounds__(KA_THREADS)
rows_from_tiles_kernel(const int *__restrict__ tile_cnt, int num_tiles, int tile_cols,
                       const unsigned *__restrict__ row_bits, int M, int *__restrict__ rows, int cap) {
  __shared__ int s_warp[32];
  __shared__ int s_off[KR_TILES], s_list[KR_TILES];
  __shared__ int s_n, s_base;
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = blockIdx.x * KR_TILES;
  const int c = (tid < KR_TILES && t0 + tid < num_tiles) ? __ldg(tile_cnt + t0 + tid) : 0;
  if (tid == 0) s_n = 0;
  if (!__syncthreads_or(c > 0)) return;  // nothing changed in this CTA's tiles
  int sum = 0;
  for (int i = tid; i < t0; i += KA_THREADS) sum += __ldg(tile_cnt + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_warp[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) s_base = w;
  }
  __syncthreads();
  int total;
  const int excl = block_excl_scan_1024(c, s_warp, &total);
  if (tid < KR_TILES) s_off[tid] = s_base + excl;
  if (c > 0) s_list[atomicAdd(&s_n, 1)] = tid;
  __syncthreads();
  const int nd = s_n;
  const int wpt = tile_cols >> 5, n_words = (M + 31) >> 5;
  for (int li = warp; li < nd; li += KA_THREADS / 32) {
    const int tl = s_list[li];
    const int w0 = (int)(((long long)(t0 + tl) * tile_cols) >> 5);
    int o = s_off[tl];
    for (int wb = 0; wb < wpt; wb += 32) {
      const int w = w0 + wb + lane;
      unsigned m = (wb + lane < wpt && w < n_words) ? row_bits[w] : 0u;
      const int pc = __popc(m);
      int incl = pc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
      }
      int at = o + incl - pc;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if (at >= 0 && at < cap) rows[at] = (w << 5) + b;
        at++;
      }
      o += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// Kernel C (grid, one warp per changed row, 8 rows per CTA): ordered emission of the row's changed
// edges (emit_row_edges_kernel's body) and, while the endpoints are in registers, the dirty-node flags
// of Integrator::findRegionConnections (mark_dirty_nodes).  The offset of a row is the sum of the
// per-row counts before it: every CTA adds up the counts in front of its first row itself (at most
// kSpecRows ints out of L2) instead of waiting for a scan kernel.  out[0] = number of edges (written
// by the warp of the last row), -1 when more than kSpecRows rows changed or tiles were left
// unresolved (the host then uses the multi-kernel path); out[2..] were cleared by sym_check_rows_kernel.
constexpr int KC_THREADS = 256;
constexpr int KC_ROWS = KC_THREADS / 32;

__global__ void __launch_bounds__(KC_THREADS)
emit_mark_kernel(const int *__restrict__ rows, const DetectCounters *__restrict__ cnt,
                 const int *__restrict__ counts, const int *__restrict__ Mp,
                 const int *__restrict__ Mi, const int *__restrict__ dof2node,
                 const int *__restrict__ meta, int *__restrict__ edges, int cap_edges, int nn,
                 int *__restrict__ out) {
  pdl_wait();
  __shared__ int s_part[KC_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = cnt->changed_rows;
  const bool bail = n > kSpecRows || cnt->pad1 > 0;
  if (bail || n == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = bail ? -1 : 0;
    return;
  }
  const int k0 = blockIdx.x * KC_ROWS;
  if (k0 >= n) return;
  int sum = 0;
  for (int i = threadIdx.x; i < k0; i += KC_THREADS) sum += counts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_part[warp] = sum;
  const int own = (lane < KC_ROWS && k0 + lane < n) ? counts[k0 + lane] : 0;
  __syncthreads();
  int at0 = 0;
#pragma unroll
  for (int w = 0; w < KC_ROWS; w++) {
    at0 += s_part[w];
    const int v = __shfl_sync(0xffffffffu, own, w);
    if (w < warp) at0 += v;
  }
  const int k = k0 + warp;
  if (k >= n) return;
  const int i = rows[k];
  const int ni = __ldg(dof2node + i);
  const int s = Mp[i], e = Mp[i + 1];
  for (int b = s; b < e; b += 32) {
    const int p = b + lane;
    bool hit = false;
    int nb = 0, nnb = 0;
    if (p < e) {
      nb = Mi[p];
      if (i < nb) {
        nnb = __ldg(dof2node + nb);
        const int small = ni < nnb ? ni : nnb, big = ni < nnb ? nnb : ni;
        hit = !(small == big || is_separator_d(small, big));  // Integrator::problematicEdgeCondition
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int at = at0 + __popc(m & ((1u << lane) - 1u));
      if (at < cap_edges) {
        edges[2 * at] = i;
        edges[2 * at + 1] = nb;
        // a problematic edge joins two nodes that are neither equal nor ancestor-related: its LCA is dirty
        out[2 + nn + common_separator_dev(meta, ni < nnb ? ni : nnb, ni < nnb ? nnb : ni)] = 1;
      }
    }
    at0 += __popc(m);
  }
  if (k == n - 1 && lane == 0) out[0] = at0;
}

// ------------------------------------------------------------------ host launchers
int launch_diff_rows(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev, int M,
                     unsigned *row_bits, DetectCounters *cnt, cudaStream_t s) {
  PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(DetectCounters), s));
  const int blocks = (M + K2_ROWS - 1) / K2_ROWS;
  diff_rows_kernel<<<blocks, K2_THREADS, 0, s>>>(Mp, Mi, Mp_prev, Mi_prev, M, row_bits, cnt);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_diff_rows_mapped(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev,
                            const int *new_to_old, const int *old_to_new, int M, unsigned *row_bits,
                            DetectCounters *cnt, cudaStream_t s) {
  PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(DetectCounters), s));
  const int threads = ((M + 31) / 32) * 32;
  diff_rows_mapped_kernel<<<(threads + 255) / 256, 256, 0, s>>>(Mp, Mi, Mp_prev, Mi_prev, new_to_old,
                                                                old_to_new, M, row_bits, cnt);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_list_changed_rows(const unsigned *row_bits, const int *word_prefix, int n_words, int M,
                             int *rows, int cap, cudaStream_t s) {
  list_changed_rows_kernel<<<(n_words + 255) / 256, 256, 0, s>>>(row_bits, word_prefix, n_words, M, rows, cap);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_count_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                           const int *dof2node, int *counts, cudaStream_t s) {
  const long long threads = (long long)n_rows * 32;
  count_row_edges_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(rows, n_rows, n_rows_dev, Mp, Mi, dof2node,
                                                                          counts);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_emit_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                          const int *dof2node, const int *offsets, int *edges, int cap, cudaStream_t s) {
  const long long threads = (long long)n_rows * 32;
  emit_row_edges_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(rows, n_rows, n_rows_dev, Mp, Mi, dof2node,
                                                                         offsets, edges, cap);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_rows_from_tiles(const int *tile_cnt, int *tile_off, unsigned long long *scan_ws, int num_tiles, int tile_cols,
                           const unsigned *row_bits, int M, int *rows, int cap, cudaStream_t s) {
  (void)tile_off; (void)scan_ws;  // the offsets are computed inside the kernel
  launch_pdl(rows_from_tiles_kernel, dim3((unsigned)((num_tiles + KR_TILES - 1) / KR_TILES)), dim3(KA_THREADS), 0, s, tile_cnt,
             num_tiles, tile_cols, row_bits, M, rows, cap);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_emit_mark(const int *rows, const DetectCounters *cnt, const int *counts, const int *Mp, const int *Mi,
                     const int *dof2node, const int *meta, int *edges, int cap_edges, int nn, int *out, cudaStream_t s) {
  launch_pdl(emit_mark_kernel, dim3(kSpecRows / KC_ROWS), dim3(KC_THREADS), 0, s, rows, cnt, counts, Mp, Mi, dof2node, meta,
             edges, cap_edges, nn, out);
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
