// parth_b200/csrc/graph_build_pipe.cu -- stage (a), dim == 1: persistent bulk-copy pipeline.
//
// Same mathematics as build_cols_spec_kernel (graph_build_cols.cu): when every column of a
// structurally symmetric, column-sorted matrix holds its diagonal, the mesh graph of
// ParthAPI::computeMeshFromMatrix (reference src/Parth/ParthAPI.cpp:366-418) is the input with
// one sample (the diagonal) removed per column, so Mp[c] = Ap[c] - c in closed form and tiles are
// independent.  What changes is how the bytes move:
//   * tiles are fixed runs of TC columns (TC chosen on the host from the average column length);
//   * a persistent CTA (3 per SM in the default shape) walks its tiles with a two-stage ring: a
//     producer warp issues cp.async.bulk (TMA engine, SASS UBLKCP) copies of the NEXT tile's Ai
//     window and Ap slice straight into shared memory and signals an mbarrier, while 256 consumer
//     threads work on the current tile (two columns each) -- DRAM latency is overlapped instead of
//     exposed per tile, and no register or LSU instruction is spent on staging;
//   * the compacted tile leaves through one bulk shared->global store (plus <= 6 scalar stores for
//     the unaligned head and tail), again off the LSU path, issued before the tile is compared with
//     the snapshot so that the store overlaps the comparison;
//   * stages are handed back by per-thread mbarrier arrivals; a tile costs two CTA barriers (after the
//     walk, and the OR-reduction of the comparison).
// Proof obligations are unchanged and raise the same flags: strict order, range, one diagonal per
// column (BUILD_NODIAG -> the chained-scan kernel takes over), structural symmetry by mirror
// probes plus the global upper/lower count, over-long tiles (BUILD_WIDE -> the roomier tile shape,
// then the sample-tiled kernel).
#include <cstdlib>
#include <string>

#include "build.cuh"
#include "tree.cuh"

namespace pb {

// Tile shape as a compile-time configuration.  CONSUMERS threads walk TC <= TCMAX columns of about TILE
// samples per tile (CAP staged at most); shared memory per CTA and the resident CTAs per SM follow.
template <int CONSUMERS_, int TILE_, int CAP_, int TCMAX_, int CTAS_, int OUTS_ = 2>
struct K9Cfg {
  static constexpr int OUTS = OUTS_;               // output staging buffers (1: the store drains before the next walk)
  static constexpr int CONSUMERS = CONSUMERS_;
  static constexpr int THREADS = CONSUMERS_ + 32;  // + producer warp
  static constexpr int TILE = TILE_;               // target samples per tile
  static constexpr int CAP = CAP_;                 // staged samples per tile
  static constexpr int TCMAX = TCMAX_;             // columns per tile, upper bound
  static constexpr int CTAS = CTAS_;               // resident CTAs per SM
};
// Measured at 200^3 (8 M columns, 7 samples per column), fused with change detection, B200: the kernel is bound
// by its instruction stream, and what a thread spends per TILE (waits, barriers, index arithmetic) weighs as much
// as the column walk itself, so two columns per thread and tile beat one: 256 consumers 141.9 us (default),
// 512 consumers + double tiles on one CTA per SM 143.1 us, 512 consumers (one column each) 153.8 us, 192: 151.8 us,
// 128 (four columns each, too few warps) 163.6 us, 4 x 256-thread CTAs on small tiles 158.3 us, 3 x 384: 155.1 us.
using K9Tri = K9Cfg<256, 3584, 4096, 512, 3, 1>;   // 72.5 KB, three CTAs per SM: one output buffer, 14 % tile headroom
using K9Half = K9Cfg<256, 3584, 4608, 1024, 2>;    // 105 KB, two CTAs per SM, two columns per thread and tile (default)
using K9Wide = K9Cfg<512, 3584, 4608, 1024, 2>;    // same tiles, one column per thread and tile
using K9One2 = K9Cfg<512, 7168, 9216, 1024, 1>;    // 197 KB, one CTA per SM, two columns per thread and tile

constexpr int K9_REG = 8;
constexpr int K9_LONG = 64;
constexpr int K9_STAGES = 2;

template <class C>
struct K9Stage {
  int val[C::CAP + 8];  // the 16-byte granular copy may fetch up to 3 samples on either side
  int ap[C::TCMAX + 8];
};
template <class C>
struct K9Smem {
  K9Stage<C> st[K9_STAGES];
  alignas(16) int out[C::OUTS][C::CAP + 8];
  unsigned long long full[K9_STAGES], empty[K9_STAGES];
  int par[K9_STAGES][4];  // q0, q1, wide
  int longs[C::CAP / K9_LONG + 2];
  int nlong;
  // fused change detection: the snapshot's rows of the tile being worked on (single stage: it is
  // fetched while the tile is walked and consumed right after)
  alignas(16) int pval[C::CAP + 8];
  alignas(16) int pap[C::TCMAX + 8];
  alignas(8) unsigned long long pfull, pempty;
  int ppar[4];  // Mp_prev[c0], staged?, Mp_prev[c1] - Mp_prev[c0]
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "K9_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra K9_DONE;\n"
      "bra K9_WAIT;\n"
      "K9_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680)  // suspend-time hint: a waiting thread sleeps instead of re-issuing the probe
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void consumer_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }
// barrier + OR over the consumer threads
template <int N>
__device__ __forceinline__ bool consumer_any_n(bool p) {
  int r;
  asm volatile(
      "{\n"
      ".reg .pred q, o;\n"
      "setp.ne.s32 q, %1, 0;\n"
      "bar.red.or.pred o, 1, %2, q;\n"
      "selp.s32 %0, 1, 0, o;\n"
      "}\n"
      : "=r"(r)
      : "r"((int)p), "n"(N)
      : "memory");
  return r != 0;
}

// is c in the ascending column [a, b)?  first the slot a symmetric stencil predicts, then a search
__device__ __forceinline__ bool k9_has_mirror(const int *__restrict__ Ai, const int *val, int sbase, int q0, int q1,
                                              int a, int b, int idx, int len, int c) {
  const bool in_tile = a >= q0 && b <= q1;
  const int g = a + len - 1 - idx;
  if (g >= a && g < b) {
    const int v = in_tile ? val[g - sbase] : __ldg(Ai + g);
    if (v == c) return true;
  }
  int lo = a, hi = b;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = in_tile ? val[mid - sbase] : __ldg(Ai + mid);
    if (v < c) lo = mid + 1; else hi = mid;
  }
  if (lo >= b) return false;
  return (in_tile ? val[lo - sbase] : __ldg(Ai + lo)) == c;
}

// The column walk for a column of exactly L samples without any proof obligation beyond the local
// ones (steady state: symmetry follows from the changed rows).  No predication on the length: L loads,
// strict order, range from the two ends, and -- the column being strictly ascending with its diagonal c
// in it -- output slot i takes r[i] while r[i] < c and r[i + 1] from the diagonal on.  Returns false
// (nothing written) when a check fails; the general walk then finds out which flag to raise.
template <int L>
__device__ __forceinline__ bool k9_walk_fixed(const int *sp, int *op, int c, int M) {
  int r[L];
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = sp[i];
  bool bad = r[0] < 0 || r[L - 1] >= M;
  bool has = false;
#pragma unroll
  for (int i = 0; i < L; i++) {
    if (i + 1 < L) bad |= r[i] >= r[i + 1];
    has |= r[i] == c;
  }
  if (bad || !has) return false;
#pragma unroll
  for (int i = 0; i + 1 < L; i++) op[i] = r[i] < c ? r[i] : r[i + 1];
  return true;
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
build_pipe_kernel(const int *__restrict__ Ap, const int *__restrict__ Ai, int M, int Q, int TC, int num_tiles,
                  int *__restrict__ Mp, int *__restrict__ Mi, BuildStatus *__restrict__ st, int check_mirror,
                  const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev, int nnz_prev,
                  int *__restrict__ tile_state, unsigned *__restrict__ row_bits, DetectCounters *__restrict__ dcnt,
                  int tile_begin, int M_rows) {
  // Row-block mode (one mesh over several GPUs, stage_sharded.cu): this launch owns tiles [tile_begin, num_tiles)
  // = columns [tile_begin * TC, M) of a matrix with M_rows columns.  Every array pointer is biased so that GLOBAL
  // column ids / entry offsets index it (only the block's own range is ever touched); M and Q are the ends of the
  // block's slices of Ap / Ai, nnz_prev the end of the snapshot's slice.  Whole matrix: tile_begin 0, M_rows == M.
  extern __shared__ __align__(128) unsigned char k9_raw[];
  static_assert(sizeof(K9Stage<C>) % 16 == 0, "bulk copies need 16-byte aligned stage buffers");
  constexpr int K9_CONSUMERS = C::CONSUMERS, K9_CAP = C::CAP;
  auto consumer_sync = [] { consumer_sync_n<C::CONSUMERS>(); };
  // the store that last used the output buffer about to be written has drained its shared-memory reads
  auto bulk_wait_read1 = [] { bulk_wait_read<C::OUTS - 1>(); };
  auto consumer_any = [](bool p) { return consumer_any_n<C::CONSUMERS>(p); };
  K9Smem<C> &S = *reinterpret_cast<K9Smem<C> *>(k9_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = tid >= K9_CONSUMERS;

  if (tid == 0) {
    // full: one arrival (the producer's expect_tx) + the copied bytes; empty: every consumer arrives on its own
    // after its last read of the stage -- no CTA barrier is spent on handing a stage back
    for (int s = 0; s < K9_STAGES; s++) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], K9_CONSUMERS); }
    mbar_init(&S.pfull, 1);
    mbar_init(&S.pempty, K9_CONSUMERS);
    S.nlong = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (producer) {
    if (lane != 0) return;
    int k = 0;
    for (int t = tile_begin + blockIdx.x; t < num_tiles; t += gridDim.x, k++) {
      const int s = k & 1;
      const int c0 = t * TC, c1 = min(M, c0 + TC), nc = c1 - c0;
      const int q0 = __ldg(Ap + c0), q1 = __ldg(Ap + c1);
      const int nq = q1 - q0;
      if (k >= K9_STAGES) mbar_wait(&S.empty[s], ((k / K9_STAGES) - 1) & 1);
      const bool wide = nq < 0 || nq > K9_CAP;
      S.par[s][0] = q0; S.par[s][1] = q1; S.par[s][2] = wide;
      unsigned bytes_ai = 0, bytes_ap = 0;
      const int sbase = q0 & ~3;
      if (!wide) {
        const int n4 = (q1 - sbase + 3) >> 2, safe4 = (Q - sbase) >> 2;
        bytes_ai = (unsigned)(n4 < safe4 ? n4 : safe4) * 16u;
        if (nq == 0) bytes_ai = 0;
        const int a4 = (nc + 1 + 3) >> 2, asafe4 = (M + 1 - c0) >> 2;
        bytes_ap = (unsigned)(a4 < asafe4 ? a4 : asafe4) * 16u;
      }
      mbar_expect_tx(&S.full[s], bytes_ai + bytes_ap);
      if (bytes_ai) bulk_g2s(S.st[s].val, Ai + sbase, bytes_ai, &S.full[s]);
      if (bytes_ap) bulk_g2s(S.st[s].ap, Ap + c0, bytes_ap, &S.full[s]);
      if (Mp_prev) {
        // the same rows of the snapshot, once the consumers are done with the previous tile's
        const int pp0 = __ldg(Mp_prev + c0), pp1 = __ldg(Mp_prev + c1);
        const int lenp = pp1 - pp0;
        const bool fits = !wide && pp0 >= 0 && pp1 <= nnz_prev && lenp >= 0 && lenp <= K9_CAP;
        if (k >= 1) mbar_wait(&S.pempty, (k - 1) & 1);
        S.ppar[0] = pp0; S.ppar[1] = fits; S.ppar[2] = lenp;
        unsigned bytes_pv = 0, bytes_pa = 0;
        if (fits) {
          const int pbase = pp0 & ~3;
          const int n4 = (pp1 - pbase + 3) >> 2, safe4 = (nnz_prev - pbase) >> 2;
          bytes_pv = lenp > 0 ? (unsigned)(n4 < safe4 ? n4 : safe4) * 16u : 0u;
          const int a4 = (nc + 1 + 3) >> 2, asafe4 = (M + 1 - c0) >> 2;
          bytes_pa = (unsigned)(a4 < asafe4 ? a4 : asafe4) * 16u;
        }
        mbar_expect_tx(&S.pfull, bytes_pv + bytes_pa);
        if (bytes_pv) bulk_g2s(S.pval, Mi_prev + (pp0 & ~3), bytes_pv, &S.pfull);
        if (bytes_pa) bulk_g2s(S.pap, Mp_prev + c0, bytes_pa, &S.pfull);
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  unsigned flags = 0;
  long long diff = 0;
  int k = 0;
  bool order_first = false;  // something of the previous tile must be ordered before this tile's walk
  for (int t = tile_begin + blockIdx.x; t < num_tiles; t += gridDim.x, k++) {
    const int s = k & 1;
    K9Stage<C> &B = S.st[s];
    int *const Bout = S.out[k % C::OUTS];
    if (tid == 0) {
      // out[s] is written again in this tile: the store that last used it must have drained its smem reads.
      // Fused mode waits for that before the OR-barrier of the previous tile instead (no barrier here).
      if (!Mp_prev) bulk_wait_read1();
      else tile_state[t] = 0;            // (ordered before the tile's atomicAdds by the consumer barriers below)
    }
    mbar_wait(&S.full[s], (k / K9_STAGES) & 1);
    const int c0 = t * TC, c1 = min(M, c0 + TC), nc = c1 - c0;
    const int q0 = S.par[s][0], q1 = S.par[s][1];
    const bool wide = S.par[s][2] != 0;
    const int nq = q1 - q0;
    const int sbase = q0 & ~3;
    bool by_threads = false;
    if (!wide) {
      // elements the 16-byte granular bulk copies could not fetch without leaving the arrays (next to their end only)
      const int got = (min((q1 - sbase + 3) >> 2, (Q - sbase) >> 2)) << 2;
      const int gota = (min((nc + 1 + 3) >> 2, (M + 1 - c0) >> 2)) << 2;
      by_threads = got < q1 - sbase || gota < nc + 1;
      if (by_threads) {
        for (int j = got + tid; j < q1 - sbase; j += K9_CONSUMERS) B.val[j] = ld_stream(Ai + sbase + j);
        for (int j = gota + tid; j < nc + 1; j += K9_CONSUMERS) B.ap[j] = __ldg(Ap + c0 + j);
      }
    }
    // every thread has seen the stage's mbarrier itself: a CTA barrier is needed only when threads staged
    // elements for each other, or the previous tile left something to order (CTA-uniform conditions)
    if (!Mp_prev || by_threads || order_first) consumer_sync();
    order_first = false;
    const long long base = (long long)q0 - c0;  // Mp[c0]
    const int mis = (int)(base & 3);              // out[] is laid out with the alignment of Mi + base
    int total = 0;
    if (wide) {
      if (tid == 0) atomicOr(&st->flags, BUILD_WIDE);
    } else {
      total = nq - nc;
      for (int lc = tid; lc < nc; lc += K9_CONSUMERS) {
        const int c = c0 + lc;
        const int vs = B.ap[lc], ve = B.ap[lc + 1];
        const int len = ve - vs;
        Mp[c] = vs - c;
        if (len <= 0) { flags |= len < 0 ? BUILD_RANGE : BUILD_NODIAG; continue; }
        const int o0 = (vs - q0) - lc;
        if (o0 < 0 || vs < q0 || ve > q1) { flags |= BUILD_NODIAG; continue; }
        const int *sp = B.val + (vs - sbase);
        int *op = Bout + mis + o0;
        bool walked = false;
        if (!check_mirror) {
          // threads of a warp mostly agree on the length (regular meshes): a branch per length, not a vote
          switch (len) {
            case 3: walked = k9_walk_fixed<3>(sp, op, c, M_rows); break;
            case 4: walked = k9_walk_fixed<4>(sp, op, c, M_rows); break;
            case 5: walked = k9_walk_fixed<5>(sp, op, c, M_rows); break;
            case 6: walked = k9_walk_fixed<6>(sp, op, c, M_rows); break;
            case 7: walked = k9_walk_fixed<7>(sp, op, c, M_rows); break;
            case 8: walked = k9_walk_fixed<8>(sp, op, c, M_rows); break;
            default: break;
          }
        }
        if (walked) {
        } else if (len <= K9_REG) {
          int r[K9_REG];
#pragma unroll
          for (int i = 0; i < K9_REG; i++) r[i] = i < len ? sp[i] : 0x7fffffff;
          int di = K9_REG, low = 0;
          bool bad = r[0] < 0;
#pragma unroll
          for (int i = 0; i < K9_REG; i++) {
            if (i + 1 < K9_REG) bad |= (i + 1 < len) & (r[i] >= r[i + 1]);
            if (r[i] == c) di = i;
          }
          if (check_mirror) {  // lower / upper split: only the full symmetry proof needs it
#pragma unroll
            for (int i = 0; i < K9_REG; i++) low += r[i] < c;
          }
          const int lastv = sp[len - 1];
          if (bad) {
            int prev = -1;
            for (int i = 0; i < len; i++) {
              if (sp[i] <= prev) flags |= sp[i] < 0 ? BUILD_RANGE : (sp[i] == prev ? BUILD_DUP : BUILD_UNSORTED);
              prev = sp[i];
            }
          }
          if (lastv >= M_rows) flags |= BUILD_RANGE;
          const int has = di < K9_REG;
          if (!has) flags |= BUILD_NODIAG;
          if (check_mirror) diff += (len - has - low) - low;
          // drop the diagonal by shifting the registers, then store the first len-1 (or len) slots
#pragma unroll
          for (int i = 0; i < K9_REG; i++) {
            const int v = (i >= di && i + 1 < K9_REG) ? r[i + 1] : r[i];
            if (i < len - has) op[i] = v;
          }
          if (check_mirror && !bad && lastv < M_rows) {
            for (int i = low + has; i < len; i += 4) {
              int a[4], b[4];
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const bool v = i + u < len;
                const int rr = v ? sp[i + u] : 0;
                a[u] = v ? __ldg(Ap + rr) : 0;
                b[u] = v ? __ldg(Ap + rr + 1) : 0;
              }
#pragma unroll
              for (int u = 0; u < 4; u++)
                if (i + u < len && !k9_has_mirror(Ai, B.val, sbase, q0, q1, a[u], b[u], i + u, len, c)) flags |= BUILD_ASYM;
            }
          }
        } else if (len <= K9_LONG) {
          int prev = -1, low = 0, dpos = -1;
          for (int i = 0; i < len; i++) {
            const int r = sp[i];
            if (r <= prev) flags |= (r < 0) ? BUILD_RANGE : (r == prev ? BUILD_DUP : BUILD_UNSORTED);
            if (r == c) dpos = i;
            low += r < c;
            prev = r;
          }
          if (prev >= M_rows) flags |= BUILD_RANGE;
          if (dpos < 0) flags |= BUILD_NODIAG;
          const int has = dpos >= 0;
          diff += (len - has - low) - low;
          for (int i = 0; i < len; i++)
            if (i != dpos) op[i - (has && i > dpos)] = sp[i];
          if (check_mirror && !(flags & (BUILD_RANGE | BUILD_UNSORTED)))
            for (int i = low + has; i < len; i++) {
              const int r = sp[i];
              if (!k9_has_mirror(Ai, B.val, sbase, q0, q1, __ldg(Ap + r), __ldg(Ap + r + 1), i, len, c)) flags |= BUILD_ASYM;
            }
        } else {
          S.longs[atomicAdd(&S.nlong, 1)] = lc;
        }
      }
    }
    fence_proxy_async();
    consumer_sync();
    const int nlong = S.nlong;
    if (nlong > 0) {
      for (int li = warp; li < nlong; li += K9_CONSUMERS / 32) {
        const int lc = S.longs[li];
        const int c = c0 + lc;
        const int vs = B.ap[lc], ve = B.ap[lc + 1], len = ve - vs;
        const int o0 = (vs - q0) - lc;
        if (o0 < 0 || vs < q0 || ve > q1) { flags |= BUILD_NODIAG; continue; }
        const int *sp = B.val + (vs - sbase);
        int *op = Bout + mis + o0;
        int dp = -1, low = 0;
        unsigned f = 0;
        for (int i0 = 0; i0 < len; i0 += 32) {
          const int i = i0 + lane;
          const int r = i < len ? sp[i] : 0x7fffffff;
          int prev = __shfl_up_sync(0xffffffffu, r, 1);
          if (lane == 0) prev = i0 > 0 ? sp[i0 - 1] : -1;
          if (i < len) {
            if (r <= prev) f |= (r < 0) ? BUILD_RANGE : (r == prev ? BUILD_DUP : BUILD_UNSORTED);
            if (r >= M_rows) f |= BUILD_RANGE;
            if (r == c) dp = i;
            low += r < c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          low += __shfl_xor_sync(0xffffffffu, low, o);
          dp = max(dp, __shfl_xor_sync(0xffffffffu, dp, o));
          f |= __shfl_xor_sync(0xffffffffu, f, o);
        }
        const int has = dp >= 0;
        if (!has) f |= BUILD_NODIAG;
        for (int i = lane; i < len; i += 32)
          if (i != dp) op[i - (has && i > dp)] = sp[i];
        if (check_mirror && !(f & (BUILD_RANGE | BUILD_UNSORTED)))
          for (int i = low + has + lane; i < len; i += 32) {
            const int r = sp[i];
            if (!k9_has_mirror(Ai, B.val, sbase, q0, q1, __ldg(Ap + r), __ldg(Ap + r + 1), i, len, c)) f |= BUILD_ASYM;
          }
        flags |= f;
        if (lane == 0) diff += (len - has - low) - low;
      }
      fence_proxy_async();
      consumer_sync();
      if (tid == 0) S.nlong = 0;
      order_first = true;  // the reset must be seen before the next walk appends
    }
    // ---- the compacted tile leaves through the bulk-store path, now: the store overlaps change detection
    if (total > 0) {
      int *dst = Mi + base;
      const int head = min((4 - mis) & 3, total);
      const int n16 = (total - head) >> 2;
      const int tail0 = head + 4 * n16;
      if (tid == 0 && n16 > 0) bulk_s2g(dst + head, Bout + mis + head, (unsigned)n16 * 16u);
      if (tid >= 32 && tid < 32 + head) st_stream(dst + (tid - 32), Bout[mis + tid - 32]);
      if (tid >= 64 && tid < 64 + (total - tail0)) st_stream(dst + tail0 + (tid - 64), Bout[mis + tail0 + tid - 64]);
    }
    if (tid == 0) bulk_commit();  // one group per tile, even when empty: wait_group counts groups
    if (Mp_prev) {
      // ---- fused change detection (stage b): this tile against the same rows of the snapshot,
      // both sides in shared memory.  Constant row-pointer shift + identical entries <=> no row of
      // the tile changed; otherwise one thread per row compares (Integrator.cpp:269-288) and the
      // changed-row bitmap words of the tile are written here.  tile_state[t] (zeroed before the
      // launch): number of changed rows of the tile, or -1 = not compared here (snapshot tile too
      // large / over-long columns; counted in dcnt->pad1, the host then runs diff_rows_kernel).
      mbar_wait(&S.pfull, k & 1);
      const int pp0 = S.ppar[0], lenp = S.ppar[2];
      const bool resolved = S.ppar[1] != 0 && nlong == 0 && !wide;
      if (!resolved) {
        if (tid == 0) { tile_state[t] = -1; atomicAdd(&dcnt->pad1, 1); bulk_wait_read1(); }
        consumer_sync();  // (stands in for the OR-barrier below: orders the drained store before the next tile)
      } else {
        const int pbase = pp0 & ~3, off = pp0 - pbase;
        const int gotp = lenp > 0 ? (min((pp0 + lenp - pbase + 3) >> 2, (nnz_prev - pbase) >> 2)) << 2 : 0;
        const int gotpa = (min((nc + 1 + 3) >> 2, (M + 1 - c0) >> 2)) << 2;
        auto prev_ptr = [&](int lc) { return lc < gotpa ? S.pap[lc] : __ldg(Mp_prev + c0 + lc); };
        auto prev_val = [&](int j) { return off + j < gotp ? S.pval[off + j] : __ldg(Mi_prev + pp0 + j); };
        const int *ov = Bout + mis;
        bool bad = lenp != total;
        if (!bad) {
          const int shift = (int)base - pp0;
          if (off + lenp <= gotp && nc + 1 <= gotpa) {
            // everything was staged (always, except next to the end of the arrays): bare compare
            const int want = shift + c0;
#pragma unroll 1
            for (int lc = tid; lc <= nc; lc += K9_CONSUMERS) bad |= (B.ap[lc] - lc) - S.pap[lc] != want;
            const int *pv = S.pval + off;
#pragma unroll 4
            for (int i = tid; i < total; i += K9_CONSUMERS) bad |= pv[i] != ov[i];
          } else {
            for (int lc = tid; lc <= nc; lc += K9_CONSUMERS) bad |= (B.ap[lc] - (c0 + lc)) - prev_ptr(lc) != shift;
            for (int i = tid; i < total; i += K9_CONSUMERS) bad |= prev_val(i) != ov[i];
          }
        }
        // the store of the PREVIOUS tile has drained (this tile's may be pending): its out[] stage is written
        // again by the next tile, and the OR-barrier below is what orders that for every thread
        if (tid == 0) bulk_wait_read1();
        if (consumer_any(bad)) {
          int changed_here = 0;
          for (int lb = tid & ~31; lb < nc; lb += K9_CONSUMERS) {
            const int lc = lb + lane;
            bool ch = false;
            if (lc < nc) {
              const int vs = B.ap[lc], nlen = B.ap[lc + 1] - vs - 1, o0 = (vs - q0) - lc;
              const int ps = prev_ptr(lc) - pp0, plen = prev_ptr(lc + 1) - pp0 - ps;
              ch = nlen != plen;
              for (int i = 0; !ch && i < nlen; i++) ch = ov[o0 + i] != prev_val(ps + i);
            }
            const unsigned m = __ballot_sync(0xffffffffu, ch);
            if (lane == 0) { row_bits[(c0 + lb) >> 5] = m; changed_here += __popc(m); }
          }
          if (lane == 0 && changed_here) { atomicAdd(&dcnt->changed_rows, changed_here); atomicAdd(&tile_state[t], changed_here); }
          if (tid == 0) atomicAdd(&dcnt->dirty_tiles, 1);
          if (C::OUTS == 1) consumer_sync();  // the row compare read the one output buffer the next walk overwrites
        } else {
          for (int w = tid; w < (nc + 31) >> 5; w += K9_CONSUMERS) row_bits[(c0 >> 5) + w] = 0u;
        }
      }
      mbar_arrive(&S.pempty);  // this thread is done with the snapshot stage
    }
    mbar_arrive(&S.empty[s]);  // ... and with the staged input
  }
  if (tid == 0) {
    bulk_wait_all();
    if (blockIdx.x == (num_tiles - 1 - tile_begin) % gridDim.x) {
      Mp[M] = Q - M;
      st->nnz_out = Q - M;
    }
  }
  if (flags) atomicOr(&st->flags, flags);
  if (check_mirror) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
    if (lane == 0 && diff) atomicAdd(&st->diff_sum, (unsigned long long)diff);
  }
}

int build_pipe_supported(const BuildArgs &a) {
  // bulk copies need 16-byte aligned sources / destinations
  return a.dim == 1 && ((reinterpret_cast<uintptr_t>(a.Ai) | reinterpret_cast<uintptr_t>(a.Ap) |
                         reinterpret_cast<uintptr_t>(a.Mi)) & 15) == 0 && a.M > 0 && a.Q > 0;
}

// Tile shape (BuildArgs::pipe_cfg): 0 = three compact CTAs per SM (default), 1 = two CTAs with full head-room (taken
// when a tile of the compact shape overflows), 2 / 3 = experiments.  PARTH_B200_PIPE = tri | half | wide | one2 sets
// the starting shape of new handles.
int build_pipe_default_cfg() {
  static const int v = [] {
    const char *e = std::getenv("PARTH_B200_PIPE");
    if (!e) return 0;
    const std::string x(e);
    return x == "half" ? 1 : x == "wide" ? 2 : x == "one2" ? 3 : 0;
  }();
  return v;
}

template <class C>
static int tile_cols_of(const BuildArgs &a) {
  // columns per tile from the average column length, a multiple of 32 so that Ap slices stay aligned
  long long tc = (long long)C::TILE * a.M / (a.Q > 0 ? a.Q : 1);
  tc = (tc / 32) * 32;
  if (tc < 32) tc = 32;
  if (tc > C::TCMAX) tc = C::TCMAX;
  return (int)tc;
}

int build_pipe_tile_cols(const BuildArgs &a) {
  if (a.blk_TC > 0) return a.blk_TC;  // row-block mode: the tile shape of the WHOLE matrix, fixed by the caller
  switch (a.pipe_cfg) {
    case 1: return tile_cols_of<K9Half>(a);
    case 2: return tile_cols_of<K9Wide>(a);
    case 3: return tile_cols_of<K9One2>(a);
    default: return tile_cols_of<K9Tri>(a);
  }
}

int build_pipe_num_tiles(const BuildArgs &a) {
  const int TC = build_pipe_tile_cols(a);
  return ((a.blk_TC > 0 ? a.M_rows : a.M) + TC - 1) / TC;
}

template <class C>
static void launch_cfg(const BuildArgs &a, int TC, int num_tiles, bool fused, int num_sms, cudaStream_t s) {
  // once per device: the call is a driver round trip on the critical path between two updates
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    PB_CUDA(cudaFuncSetAttribute(build_pipe_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K9Smem<C>)));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int t0 = a.blk_TC > 0 ? a.blk_t0 : 0;
  const int own = num_tiles - t0;
  const int grid = own < C::CTAS * num_sms ? own : C::CTAS * num_sms;
  if (a.ev_a) cudaEventRecord(a.ev_a, s);
  build_pipe_kernel<C><<<grid, C::THREADS, sizeof(K9Smem<C>), s>>>(a.Ap, a.Ai, a.M, a.Q, TC, num_tiles, a.Mp, a.Mi, a.status,
                                                                  a.check_mirror, fused ? a.Mp_prev : nullptr,
                                                                  fused ? a.Mi_prev : nullptr, a.nnz_prev,
                                                                  fused ? a.tile_dirty : nullptr, a.row_bits, a.dcnt, t0,
                                                                  a.blk_TC > 0 ? a.M_rows : a.M);
  if (a.ev_b) cudaEventRecord(a.ev_b, s);
}

int build_pipe_launch(const BuildArgs &a, int num_sms, cudaStream_t s) {
  const int TC = build_pipe_tile_cols(a);
  const int num_tiles = a.blk_TC > 0 ? a.blk_t1 : (a.M + TC - 1) / TC;
  if (num_tiles - (a.blk_TC > 0 ? a.blk_t0 : 0) <= 0) return 0;
  const bool fused = a.Mp_prev && a.Mi_prev && a.tile_dirty && a.row_bits && a.dcnt;
  // one memset for the whole mailbox when status and counters are its two halves (every stream
  // operation costs ~2 us of stream time); the per-tile counts are zeroed by the kernel itself
  char *st0 = reinterpret_cast<char *>(a.status), *dc0 = reinterpret_cast<char *>(a.dcnt);
  if (fused && a.mail_prezeroed) {
    // zeroed behind the previous update's mailbox copy (stage_build): nothing to do on the critical path
  } else if (fused && dc0 > st0 && dc0 - st0 <= 64) {
    PB_CUDA(cudaMemsetAsync(st0, 0, (size_t)(dc0 - st0) + sizeof(DetectCounters), s));
  } else {
    PB_CUDA(cudaMemsetAsync(a.status, 0, sizeof(BuildStatus), s));
    if (fused) PB_CUDA(cudaMemsetAsync(a.dcnt, 0, sizeof(DetectCounters), s));
  }
  switch (a.pipe_cfg) {
    case 1: launch_cfg<K9Half>(a, TC, num_tiles, fused, num_sms, s); break;
    case 2: launch_cfg<K9Wide>(a, TC, num_tiles, fused, num_sms, s); break;
    case 3: launch_cfg<K9One2>(a, TC, num_tiles, fused, num_sms, s); break;
    default: launch_cfg<K9Tri>(a, TC, num_tiles, fused, num_sms, s); break;
  }
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
