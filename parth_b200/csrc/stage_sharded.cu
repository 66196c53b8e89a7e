// parth_b200/csrc/stage_sharded.cu -- ONE mesh over the ranks of a communicator: row blocks of the graph build and
// of change detection (reference src/Parth/ParthAPI.cpp:366-418, src/Parth/Integrator.cpp:259-338).
//
// Partition.  The matrix is cut into the tiles of the pipeline build kernel (graph_build_pipe.cu; the tile shape is
// fixed when the handle sees its first matrix and never moves, so a rank's rows of the snapshot stay its rows);
// rank r owns tiles [T*r/world, T*(r+1)/world) = columns [c0, c1).  It reads Ap[c0..c1] and Ai[Ap[c0]..Ap[c1]) only
// (a host caller uploads 1/world of the matrix over its own PCIe link), builds rows [c0, c1) of the mesh graph
// (Mp[c] = Ap[c] - c in closed form: blocks are independent) and keeps the same rows of the snapshot.  The separator
// tree and every per-DOF array (dofs, labels, dof2node, mesh_perm) are replicated: they are what the update edits.
//
// Steady-state update, per rank: build_pipe_kernel over the own tiles, fused with the comparison against the own rows
// of the snapshot -> ascending list of changed rows -> per changed row (both versions of the row are local):
//   * the changed edges the reference reports for it (Integrator.cpp:289-338: entries nbr > row whose tree nodes are
//     neither equal nor ancestor-related) and the dirty-node flags they raise,
//   * the DIRECTED DIFF of the row: (row, i, +) for every entry it gained, (row, i, -) for every entry it lost.
// One all-gather of a fixed-size slot per rank {header, node flags, edges, diffs}; then every rank reduces the
// gathered slots identically: flags are OR-ed, edge lists concatenated in rank order (= row order, the reference's
// order), and the symmetry of the new graph is proven from the diffs alone -- the snapshot was proven symmetric,
// so the new graph is symmetric iff the set of directed diffs is: (j, i, s) present <=> (i, j, s) present.
// (Same argument as sym_check_rows_kernel, detect.cu, with the mirror row looked up in the gathered diff list
// instead of in a remote rank's memory.)  No rank ever needs another rank's rows for detection.
//
// Re-ordering of dirty nodes needs their sub-meshes, whose rows are spread over the blocks: each rank counts / fills
// the rows it owns (extract.cu, row range) and two all-reduces (counts, entries) give every rank the sub-mesh CSR;
// the graphs are then dealt to the ranks (largest first) and the local permutations all-reduced.  Everything after
// that (labels, mesh_perm, expansion) is replicated work on replicated arrays.
#include <algorithm>
#include <cstring>

#include "build.cuh"
#include "scan.cuh"
#include "stages.cuh"
#include "tree.cuh"

namespace pb {

namespace {

constexpr int XH = 16;           // header ints of a slot
constexpr int kBlkRows = 16384;  // changed rows per rank the incremental path covers

// header fields
enum { XF_FLAGS = 0, XF_ROWS, XF_EDGES, XF_DIFFS, XF_UNRESOLVED, XF_C0, XF_C1, XF_NNZ_END, XF_ROWS_OVER };

__device__ __forceinline__ bool is_sep_d(int sep, int node) {
  int a = (node - 1) / 2;
  while (a > sep && a > 0) a = (a - 1) / 2;
  return a == sep;
}
__device__ __forceinline__ bool problematic_nodes(int na, int nb) {
  const int small = na < nb ? na : nb, big = na < nb ? nb : na;
  return !(small == big || is_sep_d(small, big));
}
__device__ __forceinline__ bool row_has(const int *__restrict__ Mi, int s, int e, int v) {
  int lo = s, hi = e;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(Mi + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo < e && __ldg(Mi + lo) == v;
}

// One warp per changed row k (grid-stride): cnt[k] = reported edges, cnt[R + k] = directed diffs; rows beyond the
// device-side count get 0 so that one scan over 2R ints gives both offset arrays.  Lanes 0-15 walk the new row,
// lanes 16-31 the snapshot row.
__global__ void blk_count_kernel(const int *__restrict__ rows, const int *__restrict__ n_rows_dev, int R,
                                 const int *__restrict__ Mp, const int *__restrict__ Mi,
                                 const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev,
                                 const int *__restrict__ dof2node, int *__restrict__ cnt) {
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n = min(*n_rows_dev, R);
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < R; k += warps) {
    int ce = 0, cd = 0;
    if (k < n) {
      const int j = rows[k];
      const int s = Mp[j], e = Mp[j + 1], s2 = Mp_prev[j], e2 = Mp_prev[j + 1];
      if (!half) {
        const int nj = __ldg(dof2node + j);
        for (int p = s + hl; p < e; p += 16) {
          const int i = Mi[p];
          ce += (j < i) && problematic_nodes(nj, __ldg(dof2node + i));
          cd += !row_has(Mi_prev, s2, e2, i);
        }
      } else {
        for (int p = s2 + hl; p < e2; p += 16) cd += !row_has(Mi, s, e, Mi_prev[p]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ce += __shfl_xor_sync(0xffffffffu, ce, o);
      cd += __shfl_xor_sync(0xffffffffu, cd, o);
    }
    if (lane == 0) { cnt[k] = ce; cnt[R + k] = cd; }
  }
}

// One warp per changed row: ordered emission into the rank's slot.  off = exclusive scan of cnt (2R + 1 entries).
// Edges in (row, stored position) order; diffs per row: gained entries in stored order, then lost entries in the
// snapshot's stored order, each as (row, 2 * i + (lost ? 1 : 0)).  Thread 0 of CTA 0 writes the header.
__global__ void blk_emit_kernel(const int *__restrict__ rows, const int *__restrict__ n_rows_dev, int R,
                                const int *__restrict__ Mp, const int *__restrict__ Mi,
                                const int *__restrict__ Mp_prev, const int *__restrict__ Mi_prev,
                                const int *__restrict__ dof2node, const int *__restrict__ meta,
                                const int *__restrict__ off, const BuildStatus *__restrict__ st,
                                const DetectCounters *__restrict__ dc, int c0, int c1, int nn, int capE, int capD,
                                int *__restrict__ slot) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_all = *n_rows_dev;
  const int n = min(n_all, R);
  int *const flags = slot + XH;
  int *const edges = flags + 2 * nn;
  int *const diffs = edges + 2 * capE;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    slot[XF_FLAGS] = (int)st->flags;
    slot[XF_ROWS] = n_all;
    slot[XF_EDGES] = off[R];
    slot[XF_DIFFS] = off[2 * R] - off[R];
    slot[XF_UNRESOLVED] = dc->pad1;
    slot[XF_C0] = c0;
    slot[XF_C1] = c1;
    slot[XF_NNZ_END] = st->nnz_out;
    slot[XF_ROWS_OVER] = n_all > R;
  }
  const int d_base = off[R];
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
    const int j = rows[k];
    const int nj = __ldg(dof2node + j);
    const int s = Mp[j], e = Mp[j + 1], s2 = Mp_prev[j], e2 = Mp_prev[j + 1];
    int ae = off[k], ad = off[R + k] - d_base;
    for (int b = s; b < e; b += 32) {
      const int p = b + lane;
      bool hit = false, gained = false;
      int i = 0, ni = 0;
      if (p < e) {
        i = Mi[p];
        if (j < i) { ni = __ldg(dof2node + i); hit = problematic_nodes(nj, ni); }
        gained = !row_has(Mi_prev, s2, e2, i);
      }
      const unsigned mh = __ballot_sync(0xffffffffu, hit), mg = __ballot_sync(0xffffffffu, gained);
      const unsigned below = (1u << lane) - 1u;
      if (hit) {
        const int at = ae + __popc(mh & below);
        if (at < capE) { edges[2 * at] = j; edges[2 * at + 1] = i; }
        // a problematic edge joins two nodes that are neither equal nor ancestor-related: their LCA is dirty
        flags[nn + common_separator_dev(meta, nj < ni ? nj : ni, nj < ni ? ni : nj)] = 1;
      }
      if (gained) {
        const int at = ad + __popc(mg & below);
        if (at < capD) { diffs[2 * at] = j; diffs[2 * at + 1] = 2 * i; }
      }
      ae += __popc(mh);
      ad += __popc(mg);
    }
    for (int b = s2; b < e2; b += 32) {
      const int p = b + lane;
      bool lost = false;
      int i = 0;
      if (p < e2) { i = Mi_prev[p]; lost = !row_has(Mi, s, e, i); }
      const unsigned ml = __ballot_sync(0xffffffffu, lost);
      if (lost) {
        const int at = ad + __popc(ml & ((1u << lane) - 1u));
        if (at < capD) { diffs[2 * at] = j; diffs[2 * at + 1] = 2 * i + 1; }
      }
      ad += __popc(ml);
    }
  }
}

// Reduction of the gathered slots, identical on every rank.  out: [0] flags OR, [1] changed rows, [2] edges,
// [3] diffs, [4] unresolved tiles, [5] asymmetric diffs, [6] a rank overflowed its edge capacity, [7] ... its diff
// capacity, [8] ... its row capacity; [XH, XH + 2nn) node flags OR-ed.  edges_out: the edge lists in rank order.
__global__ void blk_reduce_kernel(const int *__restrict__ all, int world, int S, int nn, int capE, int capD,
                                  int *__restrict__ out, int *__restrict__ edges_out, int cap_out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  auto slot = [&](int r) { return all + (size_t)r * S; };
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      int f = 0, rows = 0, ed = 0, df = 0, un = 0, oe = 0, od = 0, orow = 0;
      for (int r = 0; r < world; r++) {
        const int *h = slot(r);
        f |= h[XF_FLAGS]; rows += h[XF_ROWS]; ed += h[XF_EDGES]; df += h[XF_DIFFS]; un += h[XF_UNRESOLVED];
        oe |= h[XF_EDGES] > capE; od |= h[XF_DIFFS] > capD; orow |= h[XF_ROWS_OVER];
      }
      out[0] = f; out[1] = rows; out[2] = ed; out[3] = df; out[4] = un; out[6] = oe; out[7] = od; out[8] = orow;
    }
    for (int x = threadIdx.x; x < 2 * nn; x += blockDim.x) {
      int v = 0;
      for (int r = 0; r < world; r++) v |= slot(r)[XH + x];
      out[XH + x] = v;
    }
  }
  // edges, rank order
  int base = 0;
  for (int r = 0; r < world; r++) {
    const int *h = slot(r);
    const int ne = min(h[XF_EDGES], capE);
    const int *e = h + XH + 2 * nn;
    for (int k = tid; k < 2 * ne; k += nthreads)
      if (base + k < 2 * cap_out) edges_out[base + k] = e[k];
    base += 2 * ne;
  }
  // symmetry of the directed diffs: (j, i, s) needs (i, j, s) in the slot of the rank that owns row i
  for (int r = 0; r < world; r++) {
    const int *h = slot(r);
    const int nd = min(h[XF_DIFFS], capD);
    const int *d = h + XH + 2 * nn + 2 * capE;
    for (int k = tid; k < nd; k += nthreads) {
      const int j = d[2 * k], v = d[2 * k + 1], i = v >> 1, sgn = v & 1;
      bool found = false;
      for (int q = 0; q < world && !found; q++) {
        const int *hq = slot(q);
        if (i < hq[XF_C0] || i >= hq[XF_C1]) continue;
        const int nq = min(hq[XF_DIFFS], capD);
        const int *dq = hq + XH + 2 * nn + 2 * capE;
        int lo = 0, hi = nq;  // first diff of row i (rows ascend inside a slot)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (dq[2 * mid] < i) lo = mid + 1; else hi = mid;
        }
        for (; lo < nq && dq[2 * lo] == i; lo++)
          if (dq[2 * lo + 1] == 2 * j + sgn) { found = true; break; }
        break;
      }
      if (!found) out[5] = 1;
    }
  }
}

// ---- first build of a handle in block mode: full symmetry proof.  Every entry (j -> i) of the own rows is checked:
// a local mirror row is probed here, a remote one goes to the cross list (i, j) that its owner checks after the
// exchange.  One warp per row.
__global__ void blk_sym_local_kernel(const int *__restrict__ Mp, const int *__restrict__ Mi, int c0, int c1,
                                     int *__restrict__ cross, int cap_pairs, int *__restrict__ counters /*[0] cross, [1] bad*/) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  bool bad = false;
  for (int j = c0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < c1; j += warps) {
    const int s = Mp[j], e = Mp[j + 1];
    for (int b = s; b < e; b += 32) {
      const int p = b + lane;
      bool remote = false;
      int i = 0;
      if (p < e) {
        i = Mi[p];
        if (i >= c0 && i < c1) bad |= !row_has(Mi, Mp[i], Mp[i + 1], j);
        else remote = true;
      }
      const unsigned m = __ballot_sync(0xffffffffu, remote);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[0], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (remote) {
          const int at = base + __popc(m & ((1u << lane) - 1u));
          if (at < cap_pairs) { cross[2 * at] = i; cross[2 * at + 1] = j; }
        }
      }
    }
  }
  if (bad) counters[1] = 1;
}

__global__ void blk_sym_cross_kernel(const int *__restrict__ pairs, long long n_pairs, const int *__restrict__ Mp,
                                     const int *__restrict__ Mi, int c0, int c1, int *__restrict__ bad) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_pairs; k += (long long)gridDim.x * blockDim.x) {
    const int i = pairs[2 * k], j = pairs[2 * k + 1];
    if (i < c0 || i >= c1) continue;
    if (!row_has(Mi, Mp[i], Mp[i + 1], j)) *bad = 1;
  }
}

int slot_ints(int nn, int capE, int capD) { return XH + 2 * nn + 2 * capE + 2 * capD; }

}  // namespace

// ------------------------------------------------------------------ partition
void block_plan(parth_b200 &h, int M, int nnz) {
  if (h.blk_TC > 0) {
    if (h.blk_tiles != (M + h.blk_TC - 1) / h.blk_TC)
      throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "row-block mode: the number of columns changed (the partition is fixed per handle; clear it first)");
    return;
  }
  BuildArgs a{};
  a.M = M; a.Q = nnz; a.dim = 1; a.pipe_cfg = 0;
  const int TC = build_pipe_tile_cols(a);
  const int tiles = (M + TC - 1) / TC;
  const int world = h.comm ? h.comm->world : 1, rank = h.comm ? h.comm->rank : 0;
  h.blk_TC = TC;
  h.blk_tiles = tiles;
  h.blk_t0 = (int)((long long)tiles * rank / world);
  h.blk_t1 = (int)((long long)tiles * (rank + 1) / world);
  h.blk_c0 = std::min(M, h.blk_t0 * TC);
  h.blk_c1 = std::min(M, h.blk_t1 * TC);
}

// first build: prove the symmetry of the whole graph across the blocks
static bool block_full_symmetry(parth_b200 &h) {
  cudaStream_t s = h.stream;
  Comm &c = *h.comm;
  const int c0 = h.blk_c0, c1 = h.blk_c1;
  const int own_nnz = h.own_q1 - h.own_q0;
  h.w0.reserve((size_t)2 * std::max(own_nnz, 1) + 2, s);  // cross pairs: at most every own entry
  h.w1.reserve(16, s);
  PB_CUDA(cudaMemsetAsync(h.w1.p, 0, sizeof(int) * 16, s));
  if (c1 > c0) {
    blk_sym_local_kernel<<<2 * h.num_sms, 256, 0, s>>>(h.Mp, h.Mi, c0, c1, h.w0.p, std::max(own_nnz, 1), h.w1.p);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched++;
  }
  // counts of all ranks
  h.w2.reserve((size_t)c.world * 2 + 2, s);
  c.all_gather(h.w1.p, h.w2.p, 2, s);
  std::vector<int> cnt((size_t)c.world * 2);
  read_ints(h, h.w2.p, cnt.data(), cnt.size());
  bool bad = false;
  std::vector<size_t> off((size_t)c.world), len((size_t)c.world);
  size_t total = 0;
  for (int r = 0; r < c.world; r++) {
    bad |= cnt[2 * r + 1] != 0;
    off[r] = total;
    len[r] = (size_t)cnt[2 * r] * 2;
    total += len[r];
  }
  if (c.world > 1 && total > 0) {
    h.w3.reserve(total + 2, s);
    if (len[c.rank])
      PB_CUDA(cudaMemcpyAsync(h.w3.p + off[c.rank], h.w0.p, sizeof(int) * len[c.rank], cudaMemcpyDeviceToDevice, s));
    c.all_gather_v(h.w3.p, off.data(), len.data(), s);
    PB_CUDA(cudaMemsetAsync(h.w1.p, 0, sizeof(int) * 4, s));
    blk_sym_cross_kernel<<<2 * h.num_sms, 256, 0, s>>>(h.w3.p, (long long)(total / 2), h.Mp, h.Mi, c0, c1, h.w1.p);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched++;
    c.all_reduce_sum(h.w1.p, 1, s);
    int b = 0;
    read_ints(h, h.w1.p, &b, 1);
    bad |= b != 0;
  } else if (total > 0) {
    bad = true;  // a single rank owns every row: an entry outside [c0, c1) is out of range (the build flags it first)
  }
  return !bad;
}

// ------------------------------------------------------------------ stage (a) + (b), one block per rank
// dAp / dAi: device arrays in GLOBAL indexing of which only Ap[c0..c1] and Ai[q0..q1) are dereferenced
// (q0 = Ap[c0], q1 = Ap[c1], passed by the caller: no read-back to learn them)
void stage_build_block(parth_b200 &h, int N, int nnz, const int *dAp, const int *dAi, int q0, int q1) {
  cudaStream_t s = h.stream;
  if (!h.comm) throw StatusError(PARTH_B200_ERR_ARG, "row-block mode needs a communicator (parth_b200_comm_init*)");
  Comm &comm = *h.comm;
  const int M = N;
  block_plan(h, M, nnz);
  const int c0 = h.blk_c0, c1 = h.blk_c1;
  if (q0 < 0 || q1 < q0 || q1 > nnz || q1 - q0 < c1 - c0)
    throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix (row-block mode): Ap[c0] / Ap[c1] are inconsistent");
  StageTimer total(h, &h.stats.build_ms, false, nullptr, h.ev2);
  h.stats.proof_kernel_ms = h.stats.diff_kernel_ms = 0;
  h.block_mode = true;
  h.diff_cached = false;
  h.cur_sym_proven = false;
  h.spec_valid = false;
  h.node_flags_valid = false;
  h.pend.active = false;

  // own rows of the new graph: global offsets Mp[c0] = q0 - c0 .. Mp[c1] = q1 - c1
  h.own_q0 = q0 - c0;
  h.own_q1 = q1 - c1;
  h.own_off = h.own_q0 & 3;
  h.Mp_own.reserve((size_t)(c1 - c0) + 8, s);
  h.Mi_own.reserve((size_t)(h.own_q1 - h.own_q0) + 16, s);
  const bool incremental = !h.assume_symmetric && h.prev_sym_proven && M == h.M_n_prev && h.Mp_prev.p && h.Mi_prev.p;
  const bool have_prev = (incremental || (h.assume_symmetric && M == h.M_n_prev && h.Mp_prev.p && h.Mi_prev.p));
  const bool steady = have_prev && h.tree_init && h.tree_M == M && h.map_len == 0;

  BuildArgs a{};
  a.Ap = dAp; a.Ai = dAi; a.dim = 1;
  a.M = c1; a.Q = q1; a.M_rows = M;
  a.blk_TC = h.blk_TC; a.blk_t0 = h.blk_t0; a.blk_t1 = h.blk_t1;
  a.Mp = const_cast<int *>(h.Mp_own_b());
  a.Mi = const_cast<int *>(h.Mi_own_b());
  a.status = h.bstatus_p();
  a.check_mirror = 0;  // mirrors live on other ranks: symmetry comes from the diffs (steady) or block_full_symmetry
  a.ev_a = h.ev2; a.ev_b = h.ev3;
  h.mail_clean = false;
  h.tile_dirty.reserve((size_t)h.blk_tiles + 2, s);
  h.chg_bits.reserve((size_t)(M + 31) / 32 + 2, s);
  if (steady) {
    a.Mp_prev = h.Mp_prev_b(); a.Mi_prev = h.Mi_prev_b(); a.nnz_prev = h.prev_q1;
    a.tile_dirty = h.tile_dirty.p;
    a.row_bits = reinterpret_cast<unsigned *>(h.chg_bits.p);
    a.dcnt = h.dcnt_p();
  }
  const int nn = h.num_nodes;
  int capE = h.x_cap, capD = capE;
  BuildStatus st{};
  std::vector<int> red;
  for (int attempt = 0;; attempt++) {
    a.pipe_cfg = h.pipe_cfg;
    if (!steady) PB_CUDA(cudaMemsetAsync(h.mail.p, 0, sizeof(int) * 12, s));
    if (steady && !h.blk_tiles_zeroed) {
      // tile counts in front of the own tiles are never written: zero them once (rows_from_tiles sums them)
      PB_CUDA(cudaMemsetAsync(h.tile_dirty.p, 0, sizeof(int) * (size_t)h.blk_tiles, s));
      h.blk_tiles_zeroed = true;
    }
    if (c1 > c0) h.stats.kernels_launched += build_pipe_launch(a, h.num_sms, s);
    else {  // a rank without tiles still takes part in the collectives
      PB_CUDA(cudaMemsetAsync(h.mail.p, 0, sizeof(int) * 12, s));
      PB_CUDA(cudaEventRecord(h.ev2, s));
      PB_CUDA(cudaEventRecord(h.ev3, s));
    }
    if (!steady) {
      int m[12];
      read_ints(h, h.mail.p, m, 12);
      std::memcpy(&st, m, sizeof(st));
      // every rank must take the same decision: OR the flags
      h.w1.reserve(16, s);
      PB_CUDA(cudaMemcpyAsync(h.w1.p, &h.bstatus_p()->flags, sizeof(int), cudaMemcpyDeviceToDevice, s));
      // (bitwise OR through a sum of per-bit indicators would need a kernel; the flags are few: gather and OR on the host)
      h.w2.reserve((size_t)comm.world + 2, s);
      comm.all_gather(h.w1.p, h.w2.p, 1, s);
      std::vector<int> fl((size_t)comm.world);
      read_ints(h, h.w2.p, fl.data(), fl.size());
      unsigned f = 0;
      for (int v : fl) f |= (unsigned)v;
      st.flags = f;
      if ((f & BUILD_WIDE) && !(f & ~(BUILD_WIDE | BUILD_NODIAG)) && h.pipe_cfg == 0) { h.pipe_cfg = 1; continue; }
      break;
    }
    // ---- steady state: rows -> counts -> scan -> emit -> all-gather -> reduce -> ONE read-back
    const int S = slot_ints(nn, capE, capD);
    h.chg_rows.reserve((size_t)kBlkRows + 1, s);
    h.w3.reserve((size_t)2 * kBlkRows + 2, s);
    h.w4.reserve((size_t)2 * kBlkRows + 2, s);
    h.scan_ws.reserve(scan_ws_words((size_t)2 * kBlkRows + 1), s);
    h.x_hdr.reserve((size_t)S, s);                      // own slot
    h.x_buf.reserve((size_t)S * comm.world, s);         // gathered slots
    h.x_off.reserve((size_t)XH + 2 * nn + 2, s);        // reduced mailbox
    if (h.d_changed.cap < 2 * 65536) h.d_changed.reserve(2 * 65536, s);
    if (h.blk_t1 > 0)
      h.stats.kernels_launched += launch_rows_from_tiles(h.tile_dirty.p, nullptr, nullptr, h.blk_t1, h.blk_TC,
                                                         reinterpret_cast<unsigned *>(h.chg_bits.p), c1, h.chg_rows.p,
                                                         kBlkRows, s);
    const int *n_rows_dev = &h.dcnt_p()->changed_rows;
    blk_count_kernel<<<h.num_sms, 256, 0, s>>>(h.chg_rows.p, n_rows_dev, kBlkRows, h.Mp_own_b(), h.Mi_own_b(), h.Mp_prev_b(),
                                               h.Mi_prev_b(), h.d_dof2node.p, h.w3.p);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched += 1 + exclusive_scan<0>(h.w3.p, h.w4.p, 2 * kBlkRows, h.scan_ws.p, s);
    PB_CUDA(cudaMemsetAsync(h.x_hdr.p, 0, sizeof(int) * (size_t)(XH + 2 * nn), s));
    blk_emit_kernel<<<h.num_sms, 256, 0, s>>>(h.chg_rows.p, n_rows_dev, kBlkRows, h.Mp_own_b(), h.Mi_own_b(), h.Mp_prev_b(),
                                              h.Mi_prev_b(), h.d_dof2node.p, h.d_meta.p, h.w4.p, h.bstatus_p(), h.dcnt_p(), c0,
                                              c1, nn, capE, capD, h.x_hdr.p);
    PB_CUDA(cudaGetLastError());
    comm.all_gather(h.x_hdr.p, h.x_buf.p, (size_t)S, s);
    PB_CUDA(cudaMemsetAsync(h.x_off.p, 0, sizeof(int) * XH, s));
    blk_reduce_kernel<<<64, 256, 0, s>>>(h.x_buf.p, comm.world, S, nn, capE, capD, h.x_off.p, h.d_changed.p,
                                         (int)(h.d_changed.cap / 2));
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched += 2;
    red.resize((size_t)XH + 2 * nn);
    read_ints(h, h.x_off.p, red.data(), red.size());
    st.flags = (unsigned)red[0];
    if ((st.flags & BUILD_WIDE) && !(st.flags & ~(BUILD_WIDE | BUILD_NODIAG)) && h.pipe_cfg == 0) { h.pipe_cfg = 1; continue; }
    if (st.flags == 0 && (red[6] || red[7]) && attempt < 8) {
      // an edge or diff list did not fit its slot: every rank sees the same headers, so every rank grows alike
      capE = capD = h.x_cap = std::max(capE * 4, std::max(red[2], red[3]) + 1024);
      const int total_e = red[2];
      if ((size_t)total_e * 2 + 64 > h.d_changed.cap) h.d_changed.reserve((size_t)total_e * 2 + 64, s);
      continue;
    }
    break;
  }
  {
    float t = 0;
    cudaEventElapsedTime(&t, h.ev2, h.ev3);
    h.stats.build_kernel_ms = t;
  }
  if (st.flags & BUILD_RANGE) throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: row index out of range");
  if (st.flags != 0)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                      "set_matrix (row-block mode): the matrix needs the general build path (unsorted / duplicate entries, a "
                      "column without diagonal or over-long columns), which is not sharded");
  // publish the block (biased pointers: global row ids / entry offsets index them)
  h.Mp = h.Mp_own_b();
  h.Mi = h.Mi_own_b();
  h.M_n = M;
  h.nnzM = nnz - M;
  h.mesh_owned = true;
  h.map_len = 0;
  h.stats.build_path = 1;
  h.stats.build_deferred = 0;
  h.stats.build_fused_detect = steady ? 1 : 0;
  if (steady) {
    if (red[4] > 0 || red[8] > 0)
      throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                        "set_matrix (row-block mode): the update is outside the incremental path (tiles the build kernel "
                        "could not compare, or more than 16384 changed rows on a rank)");
    if (red[5] != 0)
      throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix (row-block mode): the matrix is not structurally symmetric");
    h.cur_sym_proven = !h.assume_symmetric;
    h.n_chg_rows = h.n_chg_rows_total = red[1];
    h.diff_cached = true;
    h.spec_valid = true;
    h.spec_n_changed = red[2];
    h.spec_flags.assign(red.begin() + XH, red.end());
  } else {
    if (h.assume_symmetric) h.cur_sym_proven = false;
    else if (!block_full_symmetry(h))
      throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix (row-block mode): the matrix is not structurally symmetric");
    else h.cur_sym_proven = true;
  }
  total.stop();
}

}  // namespace pb
