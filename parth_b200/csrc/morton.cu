// parth_b200/csrc/morton.cu -- ReorderingType MORTON_CODE (reference src/Parth/include/ParthTypes.h:11).
//
// The reference declares the ordering type but does not implement it: HMD::Permute prints "Morton code is
// not implemented yet" and returns with perm unset (src/Parth/HMD.cpp:82-83), and ParthAPI has no coordinate
// input.  BASELINE config 4 nevertheless names the path, so the device implements the textbook definition
// (restated on the CPU in oracle/morton_oracle.py, PARITY UNPINNED: no reference implementation exists):
//   * bounding box of the points, each axis quantised to `bits` bits (<= 21) in IEEE double arithmetic with
//     explicitly rounded operations (no FMA contraction, so numpy reproduces the keys bit for bit);
//   * key = bit interleave, x most significant; order = ascending (key, point id);
//   * a fine separator-tree node is ordered by the global Morton rank of its DOFs (HMD::Permute's contract:
//     perm[new] = old local id).
// The sort is a hand-written stable LSD radix sort, 8 bits per pass:
//   radix_hist_kernel     ONE read of the keys gives the global digit counts of all passes (+ radix_bases_kernel)
//   radix_scatter_kernel  one launch per pass, single sweep: tile counts chained by a decoupled look-back per digit;
//                         ballot-ranked warps (8 votes give the peer mask of a lane's digit: match.any costs 11
//                         cycles per distinct value on B200, 30 distinct values in a warp of random digits),
//                         tile sorted in shared memory first so the global stores leave in digit runs
// Passes whose digit is constant over all keys (known from the OR / AND of the keys, folded into the key
// kernel) are skipped.  Hybrid: when the lower 32 bits hold two or more live digits only the upper 32 bits are
// radix-sorted and morton_fix_runs_kernel sorts the short runs of equal upper halves in place (one thread per run);
// clustered points whose runs exceed 16 fall back to all passes.  HBM bound: 8n bytes (histogram) + 24n bytes per pass (scatter); keys 24n + 12n.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "reorder.cuh"
#include "stages.cuh"

namespace pb {

namespace {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;

typedef unsigned long long u64;

// order-preserving map double -> u64 (for atomicMin / atomicMax on the bounding box)
__device__ __forceinline__ u64 ord_of(double x) {
  const u64 u = (u64)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_back(u64 o) {
  const u64 u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)u);
}

// box[0..2] = min, box[3..5] = max (ordered encoding); box must be initialised to {~0, ~0, ~0, 0, 0, 0}
__global__ void morton_bbox_kernel(const double *__restrict__ xyz, int n, long long ps, long long cs,
                                   u64 *__restrict__ box) {
  u64 lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int a = 0; a < 3; a++) {
      const double x = xyz[i * ps + a * cs];
      if (x != x) continue;  // NaN: no influence on the box, quantised to cell 0
      const u64 o = ord_of(x);
      lo[a] = o < lo[a] ? o : lo[a];
      hi[a] = o > hi[a] ? o : hi[a];
    }
  for (int a = 0; a < 3; a++) {
    for (int d = 16; d > 0; d >>= 1) {
      const u64 l2 = __shfl_xor_sync(0xffffffffu, lo[a], d), h2 = __shfl_xor_sync(0xffffffffu, hi[a], d);
      lo[a] = l2 < lo[a] ? l2 : lo[a];
      hi[a] = h2 > hi[a] ? h2 : hi[a];
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(box + a, lo[a]);
      atomicMax(box + 3 + a, hi[a]);
    }
  }
}

// spreads the low 21 bits of v: bit i -> bit 3i
__device__ __forceinline__ u64 spread3(u64 v) {
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x001f00000000ffffull;
  v = (v | (v << 16)) & 0x001f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

// keys[i] = Morton key of point i, vals[i] = i; orand[0] |= key, orand[1] &= key
__global__ void morton_keys_kernel(const double *__restrict__ xyz, int n, long long ps, long long cs,
                                   const u64 *__restrict__ box, int bits, u64 *__restrict__ keys,
                                   int *__restrict__ vals, u64 *__restrict__ orand) {
  const double scale = (double)(1ull << bits);
  const unsigned qmax = (unsigned)((1ull << bits) - 1);
  double lo[3], ext[3];
  for (int a = 0; a < 3; a++) {
    const u64 l = box[a], hgh = box[3 + a];
    lo[a] = l <= hgh ? ord_back(l) : 0.0;
    ext[a] = l <= hgh ? __dsub_rn(ord_back(hgh), lo[a]) : 0.0;
  }
  u64 o = 0, an = ~0ull;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    u64 key = 0;
    for (int a = 0; a < 3; a++) {
      const double x = xyz[i * ps + a * cs];
      unsigned q = 0;
      if (ext[a] > 0.0 && x == x) {
        const double t = __dmul_rn(__ddiv_rn(__dsub_rn(x, lo[a]), ext[a]), scale);
        q = t >= scale ? qmax : (unsigned)t;  // t >= 0 by construction; truncation = floor
      }
      key |= spread3(q) << (2 - a);
    }
    keys[i] = key;
    vals[i] = (int)i;
    o |= key;
    an &= key;
  }
  for (int d = 16; d > 0; d >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, d);
    an &= __shfl_xor_sync(0xffffffffu, an, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(orand, o);
    atomicAnd(orand + 1, an);
  }
}

// ghist[p * 256 + d] += number of keys whose digit p (bits 8p .. 8p+7) is d, for every pass p of pass_mask at once:
// the digit totals do not depend on the order of the keys, so ONE read of the keys serves all passes
__global__ void __launch_bounds__(kRsThreads) radix_hist_kernel(const u64 *__restrict__ keys, int n, unsigned pass_mask,
                                                               int *__restrict__ ghist) {
  __shared__ int cnt[8][256];
  for (int i = threadIdx.x; i < 8 * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)kRsThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRsThreads) {
    const u64 k = keys[i];
#pragma unroll
    for (int p = 0; p < 8; p++)
      if (pass_mask & (1u << p)) atomicAdd(&cnt[p][(unsigned)(k >> (8 * p)) & 255u], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 256; i += kRsThreads) {
    const int c = (&cnt[0][0])[i];
    if (c) atomicAdd(ghist + i, c);
  }
}

// block p: gbase[p * 256 + d] = number of keys whose digit p is smaller than d
__global__ void __launch_bounds__(256) radix_bases_kernel(const int *__restrict__ ghist, int *__restrict__ gbase) {
  __shared__ int sc[33];
  int total = 0;
  const int v = ghist[blockIdx.x * 256 + threadIdx.x];
  gbase[blockIdx.x * 256 + threadIdx.x] = block_excl_scan<256>(v, sc, &total);
}

template <int ITEMS>  // keys per thread and tile
struct RsSmem {
  static constexpr int kTile = kRsThreads * ITEMS;
  u64 key[kTile];
  int val[kTile];
  int cnt[kRsWarps][256];  // per-warp digit counters, then the warp's offset inside the tile's digit run
  int start[256];          // first slot of the digit inside the sorted tile
  int delta[256];          // global base of (digit, tile) minus start
  int scan[33];
  int tile;
};

// Stable scatter of one tile by the digit at `shift`, single pass over the keys ("onesweep"): the tile's digit
// counts are chained through desc[tile * 256 + digit] by a decoupled look-back, one thread per digit (status in
// bits 63..62 as in common.cuh: aggregate / inclusive prefix), so no count kernel and no device-wide scan run per
// pass.  gbase = the digit's global base (radix_bases_kernel).  desc[256 * tiles] is the dynamic tile counter; tile
// ids are handed out in launch order, so every predecessor is resident or finished.  All of desc must be zero.
template <int ITEMS, int CTAS>
__global__ void __launch_bounds__(kRsThreads, CTAS) radix_scatter_kernel(const u64 *__restrict__ keys_in,
                                                                  const int *__restrict__ vals_in,
                                                                  u64 *__restrict__ keys_out, int *__restrict__ vals_out,
                                                                  int n, int shift, const int *__restrict__ gbase,
                                                                  u64 *__restrict__ desc, int tiles) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  constexpr int kRsItems = ITEMS, kRsTile = kRsThreads * ITEMS, kRsChunk = kRsTile / kRsWarps;  // a warp ranks kRsChunk consecutive keys
  RsSmem<ITEMS> &sm = *reinterpret_cast<RsSmem<ITEMS> *>(rs_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) sm.tile = (int)atomicAdd(desc + (size_t)256 * tiles, 1ull);
  __syncthreads();
  const int tile = sm.tile;
  const long long t0 = (long long)tile * kRsTile;
  const int valid = (int)(n - t0 < kRsTile ? n - t0 : kRsTile);
  for (int d = lane; d < 256; d += 32) sm.cnt[w][d] = 0;
  __syncwarp();

  u64 key[kRsItems];
  int val[kRsItems], rank[kRsItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRsItems; r++) {
    const int loc = w * kRsChunk + r * 32 + lane;  // position inside the tile: warp chunks are contiguous
    const bool in = loc < valid;
    key[r] = in ? keys_in[t0 + loc] : ~0ull;  // padding sorts behind every real key of the tile (stable, last)
    val[r] = in ? vals_in[t0 + loc] : -1;
  }
#pragma unroll
  for (int r = 0; r < kRsItems; r++) {
    const unsigned d = (unsigned)(key[r] >> shift) & 255u;
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const unsigned vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
      peers &= ((d >> b) & 1u) ? vote : ~vote;
    }
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (lane == leader) {
      old = sm.cnt[w][d];
      sm.cnt[w][d] = old + __popc(peers);
    }
    __syncwarp();
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
  }
  __syncthreads();

  // thread d: offsets of the warps inside digit d's run and the tile's count of digit d, published at once as the
  // tile's aggregate; the four nearest predecessors' descriptors are requested now and read after the tile has
  // been sorted in shared memory, so their latency (a third of the kernel's stall samples when the walk was done
  // on the spot) hides behind that work and the predecessors have had time to publish their prefixes
  constexpr int kAhead = 4;
  const int d = tid;
  int run = 0;
#pragma unroll
  for (int x = 0; x < kRsWarps; x++) {
    const int c = sm.cnt[x][d];
    sm.cnt[x][d] = run;
    run += c;
  }
  u64 *mine = desc + (size_t)tile * 256 + d;
  u64 ahead[kAhead];
  if (tile > 0) st_relaxed_u64(mine, kTileAgg | (u64)run);
#pragma unroll
  for (int j = 0; j < kAhead; j++)
    ahead[j] = tile - 1 - j >= 0 ? ld_relaxed_u64(desc + (size_t)(tile - 1 - j) * 256 + d) : kTilePre;
  {
    int total = 0;
    sm.start[d] = block_excl_scan<kRsThreads>(run, sm.scan, &total);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; r++) {
    const unsigned dg = (unsigned)(key[r] >> shift) & 255u;
    const int pos = sm.start[dg] + sm.cnt[w][dg] + rank[r];
    sm.key[pos] = key[r];
    sm.val[pos] = val[r];
  }
  {
    // sum the predecessors' counts back to the nearest inclusive prefix (decoupled look-back, one digit per thread)
    long long before = 0;
    bool done = tile == 0;
#pragma unroll
    for (int j = 0; j < kAhead; j++) {
      if (done) break;
      u64 x = ahead[j];
      const u64 *theirs = desc + (size_t)(tile - 1 - j) * 256 + d;
      while ((x & ~kTileVal) == 0) x = ld_relaxed_u64(theirs);
      before += (long long)(x & kTileVal);
      done = (x & ~kTileVal) == kTilePre;
    }
    for (int t = tile - 1 - kAhead; !done; t -= kAhead) {  // rare: a longer walk, again four requests at a time
      u64 y[kAhead];
#pragma unroll
      for (int j = 0; j < kAhead; j++) y[j] = t - j >= 0 ? ld_relaxed_u64(desc + (size_t)(t - j) * 256 + d) : kTilePre;
#pragma unroll
      for (int j = 0; j < kAhead; j++) {
        if (done) break;
        u64 x = y[j];
        const u64 *theirs = desc + (size_t)(t - j) * 256 + d;
        while ((x & ~kTileVal) == 0) x = ld_relaxed_u64(theirs);
        before += (long long)(x & kTileVal);
        done = (x & ~kTileVal) == kTilePre;
      }
    }
    st_relaxed_u64(mine, kTilePre | (u64)(before + run));
    sm.delta[d] = (int)(__ldg(gbase + d) + before) - sm.start[d];
  }
  __syncthreads();
  for (int i = tid; i < valid; i += kRsThreads) {
    const u64 k = sm.key[i];
    const long long out = (long long)i + sm.delta[(unsigned)(k >> shift) & 255u];
    keys_out[out] = k;
    vals_out[out] = sm.val[i];
  }
}

// Hybrid sort, second half: the pairs are sorted by the upper 32 key bits with ties in ascending point id.  The thread
// of a run's first element sorts the run by the full key (stable insertion sort: equal keys keep ascending ids).  With
// keys that spread over the box the runs are pairs and triples; a run longer than kFixRun raises *overflow and the
// caller falls back to the full radix sort (whatever was already fixed stays a valid input for it).
constexpr int kFixRun = 16;
__device__ __noinline__ void morton_fix_one_run(u64 *__restrict__ keys, int *__restrict__ vals, int n, long long i,
                                                unsigned hi, int *__restrict__ overflow) {
  u64 rk[kFixRun];
  int rv[kFixRun];
  int len = 0;
  for (long long j = i; j < n; j++) {
    const u64 k = keys[j];
    if ((unsigned)(k >> 32) != hi) break;
    if (len == kFixRun) { atomicOr(overflow, 1); return; }
    rk[len] = k;
    rv[len] = vals[j];
    len++;
  }
  for (int a = 1; a < len; a++) {
    const u64 x = rk[a];
    const int y = rv[a];
    int b = a - 1;
    while (b >= 0 && rk[b] > x) { rk[b + 1] = rk[b]; rv[b + 1] = rv[b]; b--; }
    rk[b + 1] = x;
    rv[b + 1] = y;
  }
  for (int a = 0; a < len; a++) { keys[i + a] = rk[a]; vals[i + a] = rv[a]; }
}

// a thread looks at four consecutive keys (two 16-byte loads) and their two neighbours; run heads are rare
__global__ void morton_fix_runs_kernel(u64 *__restrict__ keys, int *__restrict__ vals, int n, int *__restrict__ overflow) {
  const long long i0 = 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
  if (i0 >= n) return;
  unsigned hi[6];  // upper halves of keys[i0 - 1 .. i0 + 4]; positions outside the array never equal a neighbour
  u64 k[4];
  if (i0 + 4 <= n) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0), b = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
    k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
  } else {
    for (int j = 0; j < 4; j++) k[j] = i0 + j < n ? keys[i0 + j] : 0;
  }
  for (int j = 0; j < 4; j++) hi[j + 1] = (unsigned)(k[j] >> 32);
  const bool has_prev = i0 > 0, has_next = i0 + 4 < n;
  hi[0] = has_prev ? (unsigned)(keys[i0 - 1] >> 32) : 0;
  hi[5] = has_next ? (unsigned)(keys[i0 + 4] >> 32) : 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const long long i = i0 + j;
    if (i >= n) break;
    const bool same_prev = (j > 0 || has_prev) && hi[j] == hi[j + 1];
    const bool same_next = i + 1 < n && (j < 3 || has_next) && hi[j + 2] == hi[j + 1];
    if (!same_prev && same_next) morton_fix_one_run(keys, vals, n, i, hi[j + 1], overflow);
  }
}

__global__ void rank_of_order_kernel(const int *__restrict__ order, int n, int *__restrict__ rank) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) rank[order[i]] = (int)i;
}

// concatenated vertex v of the batch: key = (graph << 32) | global Morton rank of its DOF, val = local id
__global__ void node_keys_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr, int count, int total,
                                 const int *__restrict__ node_ptr, const int *__restrict__ dofs,
                                 const int *__restrict__ rank, u64 *__restrict__ keys, int *__restrict__ vals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  int lo = 0, hi = count;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(vtx_ptr + mid) <= v) lo = mid + 1; else hi = mid;
  }
  const int g = lo - 1, k = v - __ldg(vtx_ptr + g);
  const int dof = __ldg(dofs + __ldg(node_ptr + __ldg(list + g)) + k);
  keys[v] = ((u64)(unsigned)g << 32) | (u64)(unsigned)__ldg(rank + dof);
  vals[v] = k;
}

inline int grid_for(long long n, int threads, int cap) {
  long long b = (n + threads - 1) / threads;
  return (int)std::max<long long>(1, std::min<long long>(b, cap));
}

// Stable LSD radix sort of (key, val) pairs on the digits where diff_mask has a set bit.  The pairs start in
// (k0, v0); returns 0 / 1 = the buffer pair that holds the sorted result.
int radix_sort_pairs(parth_b200 &h, u64 *k0, int *v0, u64 *k1, int *v1, int n, u64 diff_mask) {
  cudaStream_t s = h.stream;
  // tile shape: 16 keys per thread (3 CTAs / SM) or 8 keys per thread (5 CTAs / SM, twice the look-back chains)
  static const int items = [] { const char *e = std::getenv("PARTH_B200_RADIX_ITEMS"); return e ? std::atoi(e) : 16; }();
  const bool small = items == 8;
  auto kern = small ? radix_scatter_kernel<8, 5> : radix_scatter_kernel<16, 3>;
  const size_t smem = small ? sizeof(RsSmem<8>) : sizeof(RsSmem<16>);
  const int tile_keys = kRsThreads * (small ? 8 : 16);
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (n + tile_keys - 1) / tile_keys;
  unsigned pass_mask = 0;
  for (int p = 0; p < 8; p++) pass_mask |= (((diff_mask >> (8 * p)) & 255ull) != 0) << p;
  if (!pass_mask) return 0;
  const size_t dn = (size_t)256 * tiles + 1;
  h.m_hist.reserve(2 * 8 * 256, s);  // global digit counts, then their exclusive scans
  h.m_desc.reserve(dn, s);
  int *ghist = h.m_hist.p, *gbase = h.m_hist.p + 8 * 256;
  PB_CUDA(cudaMemsetAsync(ghist, 0, sizeof(int) * 8 * 256, s));
  radix_hist_kernel<<<grid_for(n, kRsThreads * 8, h.num_sms * 8), kRsThreads, 0, s>>>(k0, n, pass_mask, ghist);
  PB_CUDA(cudaGetLastError());
  radix_bases_kernel<<<8, 256, 0, s>>>(ghist, gbase);
  PB_CUDA(cudaGetLastError());
  h.stats.kernels_launched += 2;
  int cur = 0;
  for (int p = 0; p < 8; p++) {
    if (!(pass_mask & (1u << p))) continue;
    const u64 *ki = cur ? k1 : k0;
    const int *vi = cur ? v1 : v0;
    u64 *ko = cur ? k0 : k1;
    int *vo = cur ? v0 : v1;
    PB_CUDA(cudaMemsetAsync(h.m_desc.p, 0, sizeof(u64) * dn, s));
    kern<<<tiles, kRsThreads, smem, s>>>(ki, vi, ko, vo, n, 8 * p, gbase + 256 * p, h.m_desc.p, tiles);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched += 1;
    cur ^= 1;
  }
  return cur;
}

}  // namespace

// Coordinates of the current DOFs (device pointer, element strides): computes the Morton keys, the global
// Morton order (perm[new] = old) and its inverse (rank).  Stream-ordered except for one 16-byte read-back.
void stage_set_coordinates(parth_b200 &h, const double *dxyz, int n, long long ps, long long cs, int bits) {
  if (n < 0 || (n > 0 && !dxyz) || bits < 1 || bits > 21)
    throw StatusError(PARTH_B200_ERR_ARG, "set_coordinates: bad argument (bits must be 1..21)");
  cudaStream_t s = h.stream;
  StageTimer tm(h, &h.stats.morton_ms);
  h.morton_n = 0;
  h.morton_bits = bits;
  h.morton_passes = 0;
  if (n == 0) { tm.cancel(); return; }
  h.m_k0.reserve((size_t)n, s); h.m_k1.reserve((size_t)n, s);
  h.m_v0.reserve((size_t)n, s); h.m_v1.reserve((size_t)n, s);
  h.m_rank.reserve((size_t)n, s);
  h.m_box.reserve(8, s);
  const u64 init[8] = {~0ull, ~0ull, ~0ull, 0, 0, 0, 0, ~0ull};
  h.pin_u64.reserve(8);
  for (int i = 0; i < 8; i++) h.pin_u64.p[i] = init[i];
  PB_CUDA(cudaMemcpyAsync(h.m_box.p, h.pin_u64.p, sizeof(init), cudaMemcpyHostToDevice, s));
  const int grid = grid_for(n, 256, h.num_sms * 8);
  morton_bbox_kernel<<<grid, 256, 0, s>>>(dxyz, n, ps, cs, h.m_box.p);
  PB_CUDA(cudaGetLastError());
  morton_keys_kernel<<<grid, 256, 0, s>>>(dxyz, n, ps, cs, h.m_box.p, bits, h.m_k0.p, h.m_v0.p, h.m_box.p + 6);
  PB_CUDA(cudaGetLastError());
  h.stats.kernels_launched += 2;
  PB_CUDA(cudaMemcpyAsync(h.pin_u64.p, h.m_box.p + 6, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s));
  PB_CUDA(cudaStreamSynchronize(s));
  const u64 diff = h.pin_u64.p[0] ^ h.pin_u64.p[1];  // bits that are not constant over the keys
  auto passes_of = [](u64 m) { int c = 0; for (int sh = 0; sh < 64; sh += 8) c += ((m >> sh) & 255ull) != 0; return c; };
  auto sort_from = [&](int from, u64 mask) {  // the pairs sit in buffer pair `from`; returns the pair holding the result
    h.morton_passes += passes_of(mask);
    return from == 0 ? radix_sort_pairs(h, h.m_k0.p, h.m_v0.p, h.m_k1.p, h.m_v1.p, n, mask)
                     : 1 - radix_sort_pairs(h, h.m_k1.p, h.m_v1.p, h.m_k0.p, h.m_v0.p, n, mask);
  };
  const u64 hi_mask = diff & 0xffffffff00000000ull, lo_mask = diff & 0x00000000ffffffffull;
  static const bool no_hybrid = std::getenv("PARTH_B200_MORTON_NO_HYBRID") != nullptr;
  if (passes_of(lo_mask) >= 2 && hi_mask != 0 && !no_hybrid) {
    // Hybrid: radix passes over the upper 32 bits only, then the (short) runs of equal upper halves are sorted in place
    // by one thread each.  Keys that spread over the box share their upper half with one or two others at most, so the
    // passes over the lower digits are saved; clustered points overflow the run limit and take all passes below.
    int cur = sort_from(0, hi_mask);
    int *d_over = reinterpret_cast<int *>(h.m_box.p + 6);
    PB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), s));
    morton_fix_runs_kernel<<<(int)(((long long)n + 1023) / 1024), 256, 0, s>>>(cur ? h.m_k1.p : h.m_k0.p, cur ? h.m_v1.p : h.m_v0.p, n, d_over);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched += 1;
    PB_CUDA(cudaMemcpyAsync(h.pin_u64.p, d_over, sizeof(int), cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    if (*reinterpret_cast<int *>(h.pin_u64.p) != 0) cur = sort_from(cur, diff);  // stable: ties are still in ascending id
    h.morton_cur = cur;
  } else {
    h.morton_cur = sort_from(0, diff);
  }
  h.stats.morton_passes = h.morton_passes;
  h.morton_rank_valid = false;  // the inverse is a random scatter (1.5 ms at 50 M points): built when a node is ordered
  h.morton_n = n;
  tm.stop();
}

// HMD::Permute with ReorderingType MORTON_CODE for a batch of fine nodes: r_perm = local ids in ascending
// (Morton key, DOF id) order, node after node.  r_list / r_vtx hold the batch (uploaded by the caller).
void stage_morton_nodes(parth_b200 &h, int count, int total) {
  if (h.morton_n != h.M_n)
    throw StatusError(PARTH_B200_ERR_ARG, "MORTON_CODE: set_coordinates must supply one point per current DOF");
  if (total <= 0) return;
  cudaStream_t s = h.stream;
  if (!h.morton_rank_valid) {
    rank_of_order_kernel<<<(h.morton_n + 255) / 256, 256, 0, s>>>(h.morton_cur ? h.m_v1.p : h.m_v0.p, h.morton_n, h.m_rank.p);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched += 1;
    h.morton_rank_valid = true;
  }
  // the global order buffers stay untouched: the node sort has its own pairs
  h.m_nk0.reserve((size_t)total, s); h.m_nk1.reserve((size_t)total, s);
  h.m_nv0.reserve((size_t)total, s); h.m_nv1.reserve((size_t)total, s);
  h.r_perm.reserve((size_t)total, s);
  node_keys_kernel<<<(total + 255) / 256, 256, 0, s>>>(h.r_list.p, h.r_vtx.p, count, total, h.d_node_ptr.p,
                                                       h.d_dofs.p, h.m_rank.p, h.m_nk0.p, h.m_nv0.p);
  PB_CUDA(cudaGetLastError());
  h.stats.kernels_launched += 1;
  u64 mask = 0;  // digits that can differ: the rank's bits and the graph index's bits
  for (u64 x = (u64)std::max(h.M_n - 1, 0); x; x >>= 1) mask = (mask << 1) | 1ull;
  u64 gm = 0;
  for (u64 x = (u64)std::max(count - 1, 0); x; x >>= 1) gm = (gm << 1) | 1ull;
  mask |= gm << 32;
  const int cur = radix_sort_pairs(h, h.m_nk0.p, h.m_nv0.p, h.m_nk1.p, h.m_nv1.p, total, mask);
  PB_CUDA(cudaMemcpyAsync(h.r_perm.p, cur ? h.m_nv1.p : h.m_nv0.p, sizeof(int) * (size_t)total,
                          cudaMemcpyDeviceToDevice, s));
}

}  // namespace pb
