// parth_b200/csrc/stage_symbolic.cu -- stage (d): elimination tree, column counts and nnz(L) of
// P*A*P' for the current mesh graph (full symmetric pattern, no diagonal) and perm[new] = old.
//
// Where the reference gets them: src/Parth never computes them; its ecosystem calls CHOLMOD
// (cholmod_analyze_p with a given permutation, L_NNZ = cm.lnz*2 - N; reference
// src/utils/ParthTestUtils.cpp:15-59, src/LinSysSolver/CHOLMODSolver.cpp:108-151) and carries an
// in-tree Liu etree (sym_lib::compute_etree, src/utils/etree.cpp:39-115).  Both results are
// mathematical functions of (pattern, permutation), so any exact algorithm is bit-identical.
//
// Device formulation
//   etree     Liu's algorithm with path compression is sequential along a root path, but the
//             separator tree makes sub-trees independent: in the post-order layout a column of node
//             X only sees smaller columns inside X's own sub-tree.  One warp per separator-tree node
//             (lanes walk the neighbours of one column concurrently; concurrent walks only ever
//             write the same value), nodes of one wave in one launch, waves bottom-up.  Edges that
//             break the separator property (the reference keeps them when lagging drops a dirty
//             node) become extra wave dependencies, so the result stays exact.
//   colcount  Gilbert-Ng-Peyton skeleton counting without the sequential union-find: an Euler tour
//             of the etree ranked by pointer jumping gives pre-order intervals; a neighbour k < i
//             of row i is a leaf of the row sub-tree iff no other neighbour lies in its interval;
//             consecutive leaves meet at an LCA found by binary lifting on the interval test;
//             column counts are interval sums of the per-vertex deltas.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "scan.cuh"
#include "stages.cuh"

namespace pb {

namespace {

inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

__global__ void invert_perm_kernel(const int *__restrict__ perm, int n, int *__restrict__ inv, int *__restrict__ bad) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int o = perm[k];
  if ((unsigned)o >= (unsigned)n) { *bad = 1; return; }
  inv[o] = k;
}

__global__ void fill_i32(int *__restrict__ a, long long n, int v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__device__ __forceinline__ bool heap_is_ancestor_or_self(int anc, int node) {
  while (node > anc) node = (node - 1) >> 1;
  return node == anc;
}

// columns whose smaller neighbours leave the sub-tree of their node: (node of neighbour, node of column)
__global__ void find_violations_kernel(const int *__restrict__ perm, const int *__restrict__ inv,
                                       const int *__restrict__ Mp, const int *__restrict__ Mi,
                                       const int *__restrict__ dof2node, int n, int *__restrict__ pairs, int cap,
                                       int *__restrict__ count) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int old = perm[j];
  const int X = __ldg(dof2node + old);
  for (int p = Mp[old]; p < Mp[old + 1]; p++) {
    const int nb = Mi[p];
    if (__ldg(inv + nb) >= j) continue;
    const int A = __ldg(dof2node + nb);
    if (A == X || heap_is_ancestor_or_self(X, A)) continue;
    const int at = atomicAdd(count, 1);
    if (at < cap) { pairs[2 * at] = A; pairs[2 * at + 1] = X; }
  }
}

// Liu's etree over the column ranges of one wave; one warp per range
__global__ void etree_wave_kernel(const int *__restrict__ ranges, int n_ranges, const int *__restrict__ perm,
                                  const int *__restrict__ inv, const int *__restrict__ Mp, const int *__restrict__ Mi,
                                  volatile int *anc, int *parent, int n, int *__restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_ranges) return;
  const int b = ranges[2 * w], e = ranges[2 * w + 1];
  for (int j = b; j < e; j++) {
    const int old = __ldg(perm + j);
    const int s = __ldg(Mp + old), t = __ldg(Mp + old + 1);
    for (int p = s + lane; p < t; p += 32) {
      int k = __ldg(inv + __ldg(Mi + p));
      if (k >= j) continue;
      // a walk visits every vertex at most once; anything longer means the schedule was violated
      for (int steps = 0;; steps++) {
        const int a = anc[k];
        if (a == j) break;
        anc[k] = j;
        if (a == -1) { parent[k] = j; break; }
        if (steps > n) { *err = 1; break; }
        k = a;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- fast path (valid separator tree)
// One CTA per separator-tree node.  The node's own slice of anc[] lives in shared memory; a
// neighbour in a child sub-tree enters through flat[k], the root of its already finished
// component (kept up to date by the two flatten kernels after every wave), so a walk costs one
// global load plus a few shared-memory steps.  Adjacency is prefetched 128 columns at a time by
// the whole CTA; warp 0 then walks the columns in order, lanes over the neighbours of one column.
constexpr int E2_THREADS = 128;
constexpr int E2_CHUNK = 128;
constexpr int E2_DEG = 32;

// first[r] = the smallest column of the wave that is adjacent to the finished component with root r
// (a component of a child sub-tree).  It is where Liu's walk attaches r (parent[r] = first[r]), and
// every later column that touches the component may continue from first[r] instead of from r: the
// walks of a node then never leave its own slice (shared memory) -- the dependent global loads of
// anc[r] were the critical path of the top separators.
__global__ void first_touch_kernel(const int *__restrict__ ranges, const int *__restrict__ perm,
                                   const int *__restrict__ inv, const int *__restrict__ Mp,
                                   const int *__restrict__ Mi, const int *__restrict__ flat, int *__restrict__ first) {
  const int b = ranges[2 * blockIdx.y], e = ranges[2 * blockIdx.y + 1];
  const int j = b + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= e) return;
  const int old = __ldg(perm + j);
  for (int p = __ldg(Mp + old); p < __ldg(Mp + old + 1); p++) {
    const int k = __ldg(inv + __ldg(Mi + p));
    if (k < b) atomicMin(first + __ldg(flat + k), j);
  }
}

// SLOT = unsigned short: slices of up to 65 534 columns keep their ancestor links as 16-bit offsets
// (0 = none, x - b + 1 otherwise), which halves the shared memory per CTA -- two leaves of a 200^3
// grid share an SM instead of taking turns.  SLOT = int: absolute column ids, -1 = none, also the
// layout of the global fallback for slices that do not fit.
template <typename SLOT, bool SMEM>
__global__ void __launch_bounds__(E2_THREADS)
etree_node_kernel(const int *__restrict__ ranges, const int *__restrict__ perm, const int *__restrict__ inv,
                  const int *__restrict__ Mp, const int *__restrict__ Mi, const int *__restrict__ flat,
                  const int *__restrict__ first, volatile int *anc, int *parent, int smem_cap) {
  extern __shared__ __align__(16) unsigned char e2_raw[];
  __shared__ int s_nbr[2][E2_CHUNK][E2_DEG];
  __shared__ int s_cnt[2][E2_CHUNK];
  constexpr bool SHORT = sizeof(SLOT) == 2;
  const int b = ranges[2 * blockIdx.x], e = ranges[2 * blockIdx.x + 1];
  const int tid = threadIdx.x, lane = tid & 31;
  // SMEM: every slice of the wave fits the shared memory of the launch (the host checks); else the
  // links live in the global anc[] array
  constexpr bool in_smem = SMEM;
  (void)smem_cap;
  volatile SLOT *sm = reinterpret_cast<volatile SLOT *>(e2_raw);
  if (in_smem) for (int i = tid; i < e - b; i += E2_THREADS) sm[i] = SHORT ? (SLOT)0 : (SLOT)-1;
  volatile int *gl = anc + b;
  auto own_get = [&](int x) -> int {  // ancestor link of own column x (absolute id, -1 = none)
    if constexpr (!SMEM) return gl[x - b];
    else if constexpr (SHORT) { const int v = (int)sm[x - b]; return v == 0 ? -1 : v - 1 + b; }
    else return (int)sm[x - b];
  };
  auto own_set = [&](int x, int j) {
    if constexpr (!SMEM) gl[x - b] = j;
    else if constexpr (SHORT) sm[x - b] = (SLOT)(j - b + 1);
    else sm[x - b] = (SLOT)j;
  };

  // adjacency of column c0 + t, resolved to columns of this node's own slice, into buffer `buf`.
  // Loads are issued eight neighbours at a time (independent chains Mi -> inv -> flat -> first).
  auto prefetch = [&](int c0, int buf, int t) {
    const int j = c0 + t;
    int cnt = 0;
    if (j < e) {
      const int old = __ldg(perm + j);
      const int s = __ldg(Mp + old), deg = __ldg(Mp + old + 1) - s;
      if (deg > E2_DEG) cnt = -1;  // long row: walked straight from global memory
      else
        for (int q0 = 0; q0 < deg; q0 += 8) {
          int k[8], r[8], f[8];
#pragma unroll
          for (int u = 0; u < 8; u++) k[u] = q0 + u < deg ? __ldg(inv + __ldg(Mi + s + q0 + u)) : 0x7fffffff;
#pragma unroll
          for (int u = 0; u < 8; u++) r[u] = k[u] < b ? __ldg(flat + k[u]) : -1;
#pragma unroll
          for (int u = 0; u < 8; u++) f[u] = r[u] >= 0 ? __ldg(first + r[u]) : -1;
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (k[u] >= j) continue;
            if (k[u] >= b) { s_nbr[buf][t][cnt++] = k[u]; continue; }
            // a finished component below this node: attach it at its first toucher, else enter
            // the node's own forest through that column
            if (f[u] == j) { parent[r[u]] = j; anc[r[u]] = j; }
            else if (f[u] >= b && f[u] < j) s_nbr[buf][t][cnt++] = f[u];  // always true on a valid separator tree
          }
        }
    }
    s_cnt[buf][t] = cnt;
  };

  prefetch(b, 0, tid);
  __syncthreads();
  int ci = 0;
  for (int c0 = b; c0 < e; c0 += E2_CHUNK, ci++) {
    const int buf = ci & 1;
    if (tid >= 32) {
      // three warps fetch the next chunk while warp 0 walks this one
      if (c0 + E2_CHUNK < e)
        for (int t = tid - 32; t < E2_CHUNK; t += E2_THREADS - 32) prefetch(c0 + E2_CHUNK, buf ^ 1, t);
    } else {
      const int lim = min(E2_CHUNK, e - c0);
      for (int t = 0; t < lim; t++) {
        const int j = c0 + t;
        const int cnt = s_cnt[buf][t];
        auto walk = [&](int k) {
          for (;;) {
            const int a = own_get(k);
            if (a == j) break;
            own_set(k, j);
            if (a == -1) { parent[k] = j; break; }
            k = a;
          }
        };
        if (cnt >= 0) {
          if (lane < cnt) walk(s_nbr[buf][t][lane]);
        } else {
          const int old = __ldg(perm + j);
          for (int p = __ldg(Mp + old) + lane; p < __ldg(Mp + old + 1); p += 32) {
            const int k = __ldg(inv + __ldg(Mi + p));
            if (k >= j) continue;
            if (k >= b) { walk(k); continue; }
            const int r = __ldg(flat + k), f = __ldg(first + r);
            if (f == j) { parent[r] = j; anc[r] = j; }
            else if (f >= b && f < j) walk(f);
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (in_smem) for (int i = tid; i < e - b; i += E2_THREADS) anc[b + i] = own_get(b + i);
}

// ---------------------------------------------------------------- batch-parallel node kernel
// Liu's algorithm visits the columns of a node one after the other (0.5 us per column on one warp: the top separator
// of a 200^3 grid alone took 19.8 ms, the 256 leaves 16.9 ms).  This kernel takes the columns EB_CHUNK at a time:
//   phase 1  every lower neighbour that lies before the batch is replaced by the root of its component in the forest
//            F0 of the earlier columns (all threads, full path compression: links only ever point at true ancestors);
//   phase 2  the batch is a small etree problem on {F0 roots met, batch columns} with edges (u, j), u < j.  Every edge
//            climbs from u through the provisional parents smaller than j and proposes j to the vertex it stops at
//            with an atomicMin; rounds repeat until nothing changes.  Proposals are true ancestors (a climb only
//            follows true ancestors, and j is adjacent to the sub-tree it started in), values only decrease, and at
//            the fixed point every edge (u, j) has j above u with parent[x] = the smallest such j: the elimination
//            tree, which is unique -- the result is bit-identical to the sequential walk (tests/test_gpu_symbolic.py).
// Ancestor links and provisional parents share the node's shared-memory array (-1 = none = the largest unsigned).
// A batch that holds a row longer than EB_DEG is walked sequentially by warp 0 (same state, same result).
// batch kernel: 256 columns per batch (the per-batch costs -- the dependent global loads of the adjacency, the
// barriers, the lifting table -- are latency, so a wider batch halves them per column), rows of up to 16 entries
constexpr int EB_THREADS = 256, EB_CHUNK = 256, EB_DEG = 16, EB_LEVELS = 8;  // 2^EB_LEVELS >= EB_CHUNK
__global__ void __launch_bounds__(EB_THREADS)
etree_node_batch_kernel(const int *__restrict__ ranges, const int *__restrict__ perm, const int *__restrict__ inv,
                        const int *__restrict__ Mp, const int *__restrict__ Mi, const int *__restrict__ flat,
                        const int *__restrict__ first, int *anc, int *parent, int links_in_global) {
  extern __shared__ __align__(16) unsigned char e2_raw[];
  __shared__ int s_nbr[EB_CHUNK][EB_DEG];  // sources of the batch's edges (phase 1: replaced by their F0 roots)
  __shared__ int s_pos[EB_CHUNK][EB_DEG];  // where the climb of each edge stopped last (it resumes there)
  __shared__ int s_up[EB_LEVELS][EB_CHUNK];        // binary lifting over the batch columns: 2^k-th provisional ancestor (-1: leaves the batch)
  __shared__ int s_cnt[EB_CHUNK];
  __shared__ int s_flag[2];  // [0] a round changed something, [1] the batch holds a long row
  const int b = ranges[2 * blockIdx.x], e = ranges[2 * blockIdx.x + 1];
  const int tid = threadIdx.x, lane = tid & 31;
  // the node's links: shared memory, or -- a slice that does not fit (the top separators of a 512^3 mesh) -- the node's
  // own range of the global anc[] array (initialised to -1 by the host); same algorithm, L2 latency instead of LDS
  volatile int *sm = links_in_global ? reinterpret_cast<volatile int *>(anc + b) : reinterpret_cast<volatile int *>(e2_raw);
  if (!links_in_global) for (int i = tid; i < e - b; i += EB_THREADS) sm[i] = -1;
  __syncthreads();
  for (int j0 = b; j0 < e; j0 += EB_CHUNK) {
    const int j = j0 + tid;
    // ---- adjacency of column j, resolved to columns of this node's own slice (as etree_node_kernel::prefetch)
    int cnt = 0;
    if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; }
    if (j < e) {
      const int old = __ldg(perm + j);
      const int s = __ldg(Mp + old), deg = __ldg(Mp + old + 1) - s;
      if (deg > EB_DEG) cnt = -1;
      else
        for (int q0 = 0; q0 < deg; q0 += 8) {
          int k[8], r[8], f[8];
#pragma unroll
          for (int u = 0; u < 8; u++) k[u] = q0 + u < deg ? __ldg(inv + __ldg(Mi + s + q0 + u)) : 0x7fffffff;
#pragma unroll
          for (int u = 0; u < 8; u++) r[u] = k[u] < b ? __ldg(flat + k[u]) : -1;
#pragma unroll
          for (int u = 0; u < 8; u++) f[u] = r[u] >= 0 ? __ldg(first + r[u]) : -1;
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (k[u] >= j) continue;
            if (k[u] >= b) { s_nbr[tid][cnt++] = k[u]; continue; }
            if (f[u] == j) { parent[r[u]] = j; anc[r[u]] = j; }
            else if (f[u] >= b && f[u] < j) s_nbr[tid][cnt++] = f[u];
          }
        }
    }
    s_cnt[tid] = cnt;
    __syncthreads();
    if (cnt < 0) s_flag[1] = 1;
    __syncthreads();
    if (s_flag[1]) {
      // sequential walk of this batch (a row longer than EB_DEG): warp 0, lanes over the neighbours
      if (tid < 32) {
        const int lim = min(EB_CHUNK, e - j0);
        for (int t = 0; t < lim; t++) {
          const int jj = j0 + t, c = s_cnt[t];
          auto walk = [&](int k) {
            for (;;) {
              const int a = sm[k - b];
              if (a == jj) break;
              sm[k - b] = jj;
              if (a == -1) { parent[k] = jj; break; }
              k = a;
            }
          };
          if (c >= 0) {
            if (lane < c) walk(s_nbr[t][lane]);
          } else {
            const int old = __ldg(perm + jj);
            for (int p = __ldg(Mp + old) + lane; p < __ldg(Mp + old + 1); p += 32) {
              const int k = __ldg(inv + __ldg(Mi + p));
              if (k >= jj) continue;
              if (k >= b) { walk(k); continue; }
              const int r = __ldg(flat + k), f = __ldg(first + r);
              if (f == jj) { parent[r] = jj; anc[r] = jj; }
              else if (f >= b && f < jj) walk(f);
            }
          }
          __syncwarp();
        }
      }
      __syncthreads();
      continue;
    }
    // ---- phase 1: neighbours before the batch -> roots of F0 (path halving)
    for (int q = 0; q < cnt; q++) {
      int x = s_nbr[tid][q];
      if (x >= j0) continue;
      int r = x;
      for (int v = sm[r - b]; v != -1; v = sm[r - b]) r = v;
      while (x != r) {  // second pass: everything on the path points at the root (a true ancestor: safe concurrently)
        const int v = sm[x - b];
        sm[x - b] = r;
        x = v;
      }
      s_nbr[tid][q] = r;
    }
    for (int q = 0; q < cnt; q++) s_pos[tid][q] = s_nbr[tid][q];
    __syncthreads();
    // ---- phase 2: rounds of climb + atomicMin until nothing changes.  Provisional parents grow along a chain, so
    // "the last ancestor below j" is found by binary lifting over the batch columns (a natural ordering makes the
    // whole batch one chain: walking it link by link cost 128 dependent shared-memory loads for the last columns)
    const int jend = j0 + EB_CHUNK;
    for (;;) {
      {
        int v = j < e ? sm[j - b] : -1;
        s_up[0][tid] = (v != -1 && v < jend) ? v : -1;
        __syncthreads();
#pragma unroll
        for (int k = 1; k < EB_LEVELS; k++) {
          const int y = s_up[k - 1][tid];
          s_up[k][tid] = y != -1 ? s_up[k - 1][y - j0] : -1;
          __syncthreads();
        }
      }
      bool changed = false;
      for (int q = 0; q < cnt; q++) {
        // resume where this edge stopped: that vertex stays a true ancestor of the source below j whatever the
        // parents below it have become since (they only move to closer ancestors)
        int x = s_pos[tid][q];
        if (x < j0) {  // a root of F0: one step takes it into the batch (or it stops here)
          const unsigned v0 = (unsigned)sm[x - b];
          if (v0 < (unsigned)j) x = (int)v0;
        }
        if (x >= j0) {
#pragma unroll
          for (int k = EB_LEVELS - 1; k >= 0; k--) {
            const int y = s_up[k][x - j0];
            if (y != -1 && y < j) x = y;
          }
        }
        s_pos[tid][q] = x;
        const unsigned v = (unsigned)sm[x - b];
        if (v > (unsigned)j) {  // (v < j cannot happen: x is the last ancestor below j)
          const unsigned was = atomicMin(reinterpret_cast<unsigned *>(const_cast<int *>(sm)) + (x - b), (unsigned)j);
          changed |= was > (unsigned)j;
        }
      }
      if (changed) s_flag[0] = 1;
      __syncthreads();
      const bool again = s_flag[0] != 0;
      __syncthreads();
      if (!again) break;
      if (tid == 0) s_flag[0] = 0;
      __syncthreads();
    }
    // ---- parents decided in this batch: the roots of F0 that were met and the batch columns
    for (int q = 0; q < cnt; q++) {
      const int x = s_nbr[tid][q];
      parent[x] = sm[x - b];  // an edge (x, j) exists, so x has a parent <= j (same value from every writer)
    }
    // a batch column can also receive its parent from a climb that only passed through it
    if (j < e) { const int v = sm[j - b]; if (v != -1) parent[j] = v; }
    __syncthreads();
    // the parents are in global memory: the links of the batch columns now become shortcuts (pointer doubling inside
    // the batch), so that the root searches of later batches cross a batch in one or two steps instead of walking its
    // chain -- natural orderings make every batch one long chain
    for (int round = 0; round < EB_LEVELS; round++) {
      int w = -1;
      if (j < e) {
        const int v = sm[j - b];
        if (v != -1 && v < j0 + EB_CHUNK) w = sm[v - b];
      }
      __syncthreads();
      if (w != -1) sm[j - b] = w;
      __syncthreads();
    }
  }
  if (!links_in_global) for (int i = tid; i < e - b; i += EB_THREADS) anc[b + i] = sm[i];
}

// after wave w: (i) every column of a wave-w node learns the root of its component ...
__global__ void flatten_own_kernel(const int *__restrict__ perm, const int *__restrict__ dof2node,
                                   const int *__restrict__ nodewave, int w, int n, const int *__restrict__ anc,
                                   int *__restrict__ flat) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || __ldg(nodewave + __ldg(dof2node + __ldg(perm + j))) != w) return;
  int r = j;
  for (int a = anc[r]; a != -1; a = anc[r]) r = a;
  flat[j] = r;
}
// ... (ii) and every column of an earlier wave follows its old root one hop further
__global__ void flatten_below_kernel(const int *__restrict__ perm, const int *__restrict__ dof2node,
                                     const int *__restrict__ nodewave, int w, int n, const int *__restrict__ anc,
                                     int *__restrict__ flat) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || __ldg(nodewave + __ldg(dof2node + __ldg(perm + j))) >= w) return;
  const int a = anc[flat[j]];
  if (a != -1) flat[j] = flat[a];
}

// ---------------------------------------------------------------- Euler tour of the etree
// children CSR by parent; roots hang under the virtual vertex n
__global__ void child_count_kernel(const int *__restrict__ parent, int n, int *__restrict__ cnt) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int p = parent[j];
  atomicAdd(cnt + (p == -1 ? n : p), 1);
}
__global__ void child_fill_kernel(const int *__restrict__ parent, int n, const int *__restrict__ cptr,
                                  int *__restrict__ fill, int *__restrict__ child, int *__restrict__ slot) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int p = parent[j];
  if (p == -1) p = n;
  const int at = atomicAdd(fill + p, 1);
  child[cptr[p] + at] = j;
  slot[j] = at;  // position among the siblings
}
// tour elements: 2j = "enter j", 2j+1 = "leave j"; succ links them in DFS order; element weights
// count the "enter" events, so rank-before(enter j) = pre-order number
__global__ void tour_succ_kernel(const int *__restrict__ parent, int n, const int *__restrict__ cptr,
                                 const int *__restrict__ child, const int *__restrict__ slot,
                                 unsigned long long *__restrict__ link) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int nc = cptr[j + 1] - cptr[j];
  // enter j -> enter first child, or leave j
  const unsigned succ_enter = nc > 0 ? 2u * (unsigned)child[cptr[j]] : 2u * (unsigned)j + 1u;
  // leave j -> enter next sibling, or leave parent (virtual root: end of list)
  int p = parent[j];
  if (p == -1) p = n;
  const int sl = slot[j];
  unsigned succ_leave;
  if (sl + 1 < cptr[p + 1] - cptr[p]) succ_leave = 2u * (unsigned)child[cptr[p] + sl + 1];
  else succ_leave = p == n ? 0xffffffffu : 2u * (unsigned)p + 1u;
  // link = (successor << 32) | number of "enter" events from this element to the end of its hop
  link[2 * (size_t)j] = ((unsigned long long)succ_enter << 32) | 1ull;
  link[2 * (size_t)j + 1] = ((unsigned long long)succ_leave << 32) | 0ull;
}
// ---- sampled list ranking (Helman-JaJa): Wyllie's jumping moves the whole 2n-element list log2(2n) times (25 rounds,
// 6.3 ms and 25 read-backs at 8 M DOF).  The "enter" and "leave" elements of a pseudo-random 1/32 of the vertices (a
// multiplicative hash of the vertex id: index-based sampling never meets the odd-numbered "leave" chain of a path-like
// tree) and the head are splitters; a splitter walks to the next splitter summing the weights of its segment, the
// splitters are ranked by jumping (a fixed number of rounds, no read-back), and a second walk writes the suffix sums of
// the segment's elements.  vslot = exclusive scan of the vertex flags; slot of element i = 2 vslot[i / 2] + (i & 1).
__device__ __forceinline__ bool tour_vertex_sampled(unsigned j) { return ((j * 2654435761u) >> 27) == 0u; }
__global__ void tour_flag_kernel(int n, int *__restrict__ flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) flag[j] = tour_vertex_sampled((unsigned)j);
}
__device__ __forceinline__ bool tour_is_splitter(unsigned i, unsigned head) { return tour_vertex_sampled(i >> 1) || i == head; }
__device__ __forceinline__ unsigned tour_slot(unsigned i, unsigned head, unsigned n_reg, const int *__restrict__ vslot) {
  return tour_vertex_sampled(i >> 1) ? 2u * (unsigned)__ldg(vslot + (i >> 1)) + (i & 1u) : n_reg;  // head: the extra slot
}
// spl[slot] = (next slot << 32) | segment weight
__global__ void tour_segment_kernel(const unsigned long long *__restrict__ link, long long m, unsigned head, unsigned n_reg,
                                    const int *__restrict__ vslot, unsigned long long *__restrict__ spl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == m) {  // the head's extra slot stays an empty tail when the head is a sampled element
    if (tour_vertex_sampled(head >> 1)) spl[n_reg] = 0xffffffffull << 32;
    return;
  }
  if (i > m || !tour_is_splitter((unsigned)i, head)) return;
  unsigned long long a = link[i];
  unsigned acc = (unsigned)a, cur = (unsigned)(a >> 32);
  while (cur != 0xffffffffu && !tour_is_splitter(cur, head)) {
    a = link[cur];
    acc += (unsigned)a;
    cur = (unsigned)(a >> 32);
  }
  const unsigned nxt = cur == 0xffffffffu ? 0xffffffffu : tour_slot(cur, head, n_reg, vslot);
  spl[tour_slot((unsigned)i, head, n_reg, vslot)] = ((unsigned long long)nxt << 32) | acc;
}
// suffix sums of the segment's elements from the ranked splitter (low word of ranked[slot] = weight from the
// splitter, inclusive, to the end of the tour); out keeps the link format tour_extract_kernel reads
__global__ void tour_spread_kernel(const unsigned long long *__restrict__ link, long long m, unsigned head, unsigned n_reg,
                                   const int *__restrict__ vslot, const unsigned long long *__restrict__ ranked,
                                   unsigned long long *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !tour_is_splitter((unsigned)i, head)) return;
  unsigned running = (unsigned)ranked[tour_slot((unsigned)i, head, n_reg, vslot)];
  unsigned cur = (unsigned)i;
  do {
    const unsigned long long a = link[cur];
    out[cur] = (a & 0xffffffff00000000ull) | running;
    running -= (unsigned)a;
    cur = (unsigned)(a >> 32);
  } while (cur != 0xffffffffu && !tour_is_splitter(cur, head));
}
// Wyllie jumping without the activity flag (fixed number of rounds on the short splitter list)
__global__ void tour_jump_fixed_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
                                       unsigned m) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long a = in[i];
  const unsigned s = (unsigned)(a >> 32);
  if (s == 0xffffffffu) { out[i] = a; return; }
  const unsigned long long b = in[s];
  out[i] = (b & 0xffffffff00000000ull) | (unsigned long long)((unsigned)a + (unsigned)b);
}
__global__ void tour_extract_kernel(const unsigned long long *__restrict__ link, int n, int *__restrict__ tin,
                                    int *__restrict__ size) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int after_enter = (int)(unsigned)link[2 * (size_t)j];      // enters from "enter j" on, inclusive
  const int after_leave = (int)(unsigned)link[2 * (size_t)j + 1];  // enters after "leave j"
  tin[j] = n - after_enter;
  size[j] = after_enter - after_leave;
}

// ---------------------------------------------------------------- binary lifting
__global__ void lift_init_kernel(const int *__restrict__ parent, int n, int *__restrict__ up0) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) up0[j] = parent[j];
}
__global__ void lift_step_kernel(const int *__restrict__ prev, int n, int *__restrict__ next, int *__restrict__ any) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int a = prev[j];
  const int b = a == -1 ? -1 : prev[a];
  next[j] = b;
  if (b != -1) *any = 1;
}

// ---------------------------------------------------------------- skeleton deltas
// delta[j] = [j is an etree leaf] - #children(j) + sum over rows of the leaf / LCA corrections
__global__ void delta_base_kernel(const int *__restrict__ cptr, int n, int *__restrict__ delta) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int nc = cptr[j + 1] - cptr[j];
  delta[j] = (nc == 0 ? 1 : 0) - nc;
}

__global__ void delta_rows_kernel(const int *__restrict__ perm, const int *__restrict__ inv,
                                  const int *__restrict__ Mp, const int *__restrict__ Mi, int n,
                                  const int *__restrict__ tin, const int *__restrict__ size,
                                  const int *__restrict__ up, int levels, int *__restrict__ delta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int old = perm[i];
  const int s = Mp[old], e = Mp[old + 1];
  for (int p = s; p < e; p++) {
    const int k = __ldg(inv + Mi[p]);
    if (k >= i) continue;
    const int tk = __ldg(tin + k), ek = tk + __ldg(size + k);
    // leaf of the row sub-tree <=> no other lower neighbour inside k's interval; previous leaf =
    // the lower neighbour with the largest interval that ends at or before tin[k]
    bool leaf = true;
    int best = -1, best_t = -1;
    for (int q = s; q < e; q++) {
      if (q == p) continue;
      const int k2 = __ldg(inv + Mi[q]);
      if (k2 >= i) continue;
      const int t2 = __ldg(tin + k2);
      if (t2 > tk && t2 < ek) { leaf = false; break; }
      if (t2 + __ldg(size + k2) <= tk && t2 > best_t) { best_t = t2; best = k2; }
    }
    if (!leaf) continue;
    atomicAdd(delta + k, 1);
    if (best == -1) continue;
    // LCA(best, k): lowest ancestor of `best` whose interval contains tin[k]
    int x = best;
    for (int b = levels - 1; b >= 0; b--) {
      const int a = __ldg(up + (size_t)b * n + x);
      if (a == -1) continue;
      const int ta = __ldg(tin + a);
      if (!(ta <= tk && tk < ta + __ldg(size + a))) x = a;
    }
    atomicSub(delta + __ldg(up + x), 1);  // parent of x
  }
}

__global__ void scatter_by_tin_kernel(const int *__restrict__ delta, const int *__restrict__ tin, int n,
                                      int *__restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[tin[j]] = delta[j];
}
__global__ void colcount_kernel(const int *__restrict__ pre, const int *__restrict__ tin, const int *__restrict__ size,
                                int n, int *__restrict__ colcount, unsigned long long *__restrict__ lnz) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  long long c = 0;
  if (j < n) {
    const int t = tin[j];
    c = pre[t + size[j]] - pre[t];
    colcount[j] = (int)c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(lnz, (unsigned long long)c);
}

}  // namespace

void stage_symbolic(parth_b200 &h, const int *perm_in, int *parent_out, int *colcount_out, long long *lnz_out,
                    bool on_device) {
  const int *perm_host = perm_in;
  const cudaMemcpyKind in_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  cudaStream_t s = h.stream;
  const int n = h.M_n;
  StageTimer tm(h, &h.stats.symbolic_ms);
  const bool tree_aware = perm_host == nullptr && h.tree_init && h.tree_M == n;
  if (perm_host == nullptr && !tree_aware)
    throw StatusError(PARTH_B200_ERR_ARG, "symbolic: no permutation given and no current mesh_perm");

  // scratch: all grow-only on the handle would pin ~1 GB at 8 M DOF; this stage owns its buffers
  DevBuf<int> perm, inv, anc, parent, flag;
  perm.reserve((size_t)n, s); inv.reserve((size_t)n, s); anc.reserve((size_t)n, s); parent.reserve((size_t)n, s);
  flag.reserve(8, s);
  PB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int) * 8, s));
  if (perm_host) {
    PB_CUDA(cudaMemcpyAsync(perm.p, perm_host, sizeof(int) * (size_t)n, in_kind, s));
    if (!on_device) h.stats.h2d_bytes += (long long)sizeof(int) * n;
  } else {
    PB_CUDA(cudaMemcpyAsync(perm.p, h.d_mesh_perm.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  }
  fill_i32<<<nblk(n), 256, 0, s>>>(inv.p, n, -1);
  invert_perm_kernel<<<nblk(n), 256, 0, s>>>(perm.p, n, inv.p, flag.p);
  fill_i32<<<nblk(n), 256, 0, s>>>(anc.p, n, -1);
  fill_i32<<<nblk(n), 256, 0, s>>>(parent.p, n, -1);
  h.stats.kernels_launched += 4;
  {
    int bad = 0;
    read_ints(h, flag.p, &bad, 1);
    if (bad) throw StatusError(PARTH_B200_ERR_INPUT, "symbolic: perm is not a permutation of 0..M_n-1");
  }

  // ---- waves of column ranges
  StageTimer t_etree(h, &h.stats.symbolic_etree_ms);
  std::vector<std::vector<int>> waves;  // flattened (begin, end) pairs per wave
  std::vector<int> node_wave;            // per separator-tree node (fast path only)
  bool fast = false;
  if (!tree_aware) {
    waves.push_back({0, n});
  } else {
    const int nn = h.num_nodes;
    const int cap = 1 << 16;
    DevBuf<int> pairs;
    pairs.reserve((size_t)cap * 2, s);
    find_violations_kernel<<<nblk(n), 256, 0, s>>>(perm.p, inv.p, h.Mp, h.Mi, h.d_dof2node.p, n, pairs.p, cap, flag.p + 1);
    h.stats.kernels_launched++;
    int nv = 0;
    read_ints(h, flag.p + 1, &nv, 1);
    if (nv > cap) {
      waves.push_back({0, n});  // too broken to schedule around: one sequential range
    } else {
      std::vector<int> pv((size_t)nv * 2);
      if (nv > 0) {
        PB_CUDA(cudaMemcpyAsync(pv.data(), pairs.p, sizeof(int) * 2 * (size_t)nv, cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaStreamSynchronize(s));
      }
      // dependency X -> C: C = the child of LCA(A, X) on A's side (heap numbering)
      std::vector<std::vector<int>> deps((size_t)nn);
      for (int v = 0; v < nv; v++) {
        int A = pv[2 * v], X = pv[2 * v + 1];
        int a = A, x = X;
        auto depth = [](int y) { int d = 0; while (y > 0) { y = (y - 1) >> 1; d++; } return d; };
        int da = depth(a), dx = depth(x);
        while (da > dx) { a = (a - 1) >> 1; da--; }
        while (dx > da) { x = (x - 1) >> 1; dx--; }
        if (a == x) continue;  // ancestor-related: cannot precede X in a consistent post-order layout
        while (((a - 1) >> 1) != ((x - 1) >> 1)) { a = (a - 1) >> 1; x = (x - 1) >> 1; }
        deps[X].push_back(a);
      }
      // post-order = ascending offset; wave(X) = 1 + max(children, dependencies)
      std::vector<int> order;
      for (int x = 0; x < nn; x++) if (h.meta[x * 7 + 2] != -1) order.push_back(x);
      std::sort(order.begin(), order.end(), [&](int a, int b) {
        const int oa = h.meta[a * 7 + 4], ob = h.meta[b * 7 + 4];
        return oa != ob ? oa < ob : a > b;  // empty nodes share an offset: deeper (larger id) first
      });
      // A node whose columns reach outside its sub-tree links foreign components under its own
      // columns, so any later walk may continue into it: it must run after EVERYTHING that
      // precedes it in post-order (which also orders two such nodes with respect to each other).
      std::vector<int> wave((size_t)nn, 0), subw((size_t)nn, 0);  // subw: max wave inside the sub-tree
      int max_before = -1;
      for (int x : order) {
        int w = 0;
        for (int c : {h.meta[x * 7], h.meta[x * 7 + 1]}) if (c != -1) w = std::max(w, subw[c] + 1);
        if (!deps[x].empty()) w = std::max(w, max_before + 1);
        wave[x] = w;
        max_before = std::max(max_before, w);
        subw[x] = w;
        for (int c : {h.meta[x * 7], h.meta[x * 7 + 1]}) if (c != -1) subw[x] = std::max(subw[x], subw[c]);
        const int sz = h.node_ptr[x + 1] - h.node_ptr[x];
        if (sz == 0) continue;
        if ((int)waves.size() <= w) waves.resize((size_t)w + 1);
        waves[w].push_back(h.meta[x * 7 + 4]);
        waves[w].push_back(h.meta[x * 7 + 4] + sz);
      }
      if (nv == 0) { fast = true; node_wave = wave; }
    }
  }
  {
    size_t total = 0;
    for (auto &w : waves) total += w.size();
    DevBuf<int> ranges;
    ranges.reserve(std::max<size_t>(total, 2), s);
    std::vector<int> flat;
    flat.reserve(total);
    for (auto &w : waves) flat.insert(flat.end(), w.begin(), w.end());
    if (total) {
      PB_CUDA(cudaMemcpyAsync(ranges.p, flat.data(), sizeof(int) * total, cudaMemcpyHostToDevice, s));
      PB_CUDA(cudaStreamSynchronize(s));
    }
    size_t at = 0;
    if (fast) {
      // valid separator tree: shared-memory node kernel + component flattening between the waves
      DevBuf<int> flatroot, d_wave, first;
      flatroot.reserve((size_t)n, s);
      first.reserve((size_t)n, s);
      fill_i32<<<nblk(n), 256, 0, s>>>(first.p, n, 0x7fffffff);
      h.stats.kernels_launched++;
      d_wave.reserve(node_wave.size(), s);
      PB_CUDA(cudaMemcpyAsync(d_wave.p, node_wave.data(), sizeof(int) * node_wave.size(), cudaMemcpyHostToDevice, s));
      PB_CUDA(cudaStreamSynchronize(s));
      int dev = 0, max_smem = 0;
      PB_CUDA(cudaGetDevice(&dev));
      PB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      const int dyn = max_smem - 36 * 1024;  // static part: two buffers of prefetched adjacency lists
      PB_CUDA(cudaFuncSetAttribute(etree_node_kernel<int, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
      PB_CUDA(cudaFuncSetAttribute(etree_node_kernel<unsigned short, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
      const int dyn_batch = max_smem - 44 * 1024;  // static part: the batch's edge sources, climb positions, lifting table
      PB_CUDA(cudaFuncSetAttribute(etree_node_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_batch));
      int wi = 0;
      for (auto &w : waves) {
        const int nr = (int)w.size() / 2;
        if (nr > 0) {
          if (wi > 0) {  // leaves have no finished component below them
            int longest = 0;
            for (int r = 0; r < nr; r++) longest = std::max(longest, w[2 * r + 1] - w[2 * r]);
            first_touch_kernel<<<dim3(nblk(longest), (unsigned)nr), 256, 0, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi,
                                                                                flatroot.p, first.p);
            h.stats.kernels_launched++;
          }
          // shared memory sized by the wave's largest slice (the smaller, the more CTAs per SM)
          int longest = 0;
          for (int r = 0; r < nr; r++) longest = std::max(longest, w[2 * r + 1] - w[2 * r]);
          // 16-bit slots cost ~30 % per column walk: they pay only when the wave has more nodes than the GPU has SMs
          const bool seq_only = std::getenv("PARTH_B200_ETREE_SEQ") != nullptr;
          // PARTH_B200_ETREE_SMEM_KB caps the shared-memory budget (tests: the global-links variant on small meshes)
          const char *cap_kb = std::getenv("PARTH_B200_ETREE_SMEM_KB");
          const int budget = cap_kb ? std::min(dyn_batch, std::atoi(cap_kb) * 1024) : dyn_batch;
          if (!seq_only && 4 * longest <= budget) {
            const int bytes = ((4 * longest + 1023) / 1024) * 1024;
            etree_node_batch_kernel<<<nr, EB_THREADS, bytes, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi, flatroot.p, first.p,
                                                                 anc.p, parent.p, 0);
          } else if (!seq_only) {  // a slice beyond shared memory: the same batches with the links in global memory
            etree_node_batch_kernel<<<nr, EB_THREADS, 0, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi, flatroot.p, first.p, anc.p,
                                                             parent.p, 1);
          } else if (nr > h.num_sms && longest <= 65534 && 2 * longest <= dyn) {
            const int bytes = ((2 * longest + 1023) / 1024) * 1024;
            etree_node_kernel<unsigned short, true><<<nr, E2_THREADS, bytes, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi,
                                                                                flatroot.p, first.p, anc.p, parent.p, bytes / 2);
          } else if (4 * longest <= dyn) {
            const int bytes = ((4 * longest + 1023) / 1024) * 1024;
            etree_node_kernel<int, true><<<nr, E2_THREADS, bytes, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi, flatroot.p,
                                                                     first.p, anc.p, parent.p, bytes / 4);
          } else {  // a slice larger than shared memory: links in global memory for this wave
            etree_node_kernel<int, false><<<nr, E2_THREADS, 0, s>>>(ranges.p + at, perm.p, inv.p, h.Mp, h.Mi, flatroot.p,
                                                                  first.p, anc.p, parent.p, 0);
          }
          flatten_own_kernel<<<nblk(n), 256, 0, s>>>(perm.p, h.d_dof2node.p, d_wave.p, wi, n, anc.p, flatroot.p);
          flatten_below_kernel<<<nblk(n), 256, 0, s>>>(perm.p, h.d_dof2node.p, d_wave.p, wi, n, anc.p, flatroot.p);
          h.stats.kernels_launched += 3;
        }
        at += w.size();
        wi++;
      }
      PB_CUDA(cudaStreamSynchronize(s));
    } else
    for (auto &w : waves) {
      const int nr = (int)w.size() / 2;
      if (nr > 0) {
        etree_wave_kernel<<<nblk((long long)nr * 32), 256, 0, s>>>(ranges.p + at, nr, perm.p, inv.p, h.Mp, h.Mi, anc.p,
                                                                   parent.p, n, flag.p + 4);
        h.stats.kernels_launched++;
      }
      at += w.size();
    }
    h.stats.symbolic_path = fast ? 2 : (tree_aware ? 1 : 0);
    t_etree.stop();
    int err = 0;
    read_ints(h, flag.p + 4, &err, 1);  // also keeps `ranges` alive until the waves are done
    if (err) throw StatusError(PARTH_B200_ERR_INTERNAL, "symbolic: etree walk did not terminate");
  }
  if (parent_out) {
    PB_CUDA(cudaMemcpyAsync(parent_out, parent.p, sizeof(int) * (size_t)n, out_kind, s));
    h.stats.d2h_bytes += (long long)sizeof(int) * n;
  }

  // ---- Euler tour -> pre-order intervals
  DevBuf<int> cptr, ccnt, child, slot, tin, size;
  cptr.reserve((size_t)n + 2, s); ccnt.reserve((size_t)n + 2, s); child.reserve((size_t)n + 1, s);
  slot.reserve((size_t)n + 1, s); tin.reserve((size_t)n + 1, s); size.reserve((size_t)n + 1, s);
  h.scan_ws.reserve(scan_ws_words((size_t)n + 2), s);
  PB_CUDA(cudaMemsetAsync(ccnt.p, 0, sizeof(int) * ((size_t)n + 1), s));
  child_count_kernel<<<nblk(n), 256, 0, s>>>(parent.p, n, ccnt.p);
  h.stats.kernels_launched += 1 + exclusive_scan<0>(ccnt.p, cptr.p, n + 1, h.scan_ws.p, s);
  PB_CUDA(cudaMemsetAsync(ccnt.p, 0, sizeof(int) * ((size_t)n + 1), s));
  child_fill_kernel<<<nblk(n), 256, 0, s>>>(parent.p, n, cptr.p, ccnt.p, child.p, slot.p);
  h.stats.kernels_launched++;
  {
    DevBuf<unsigned long long> la, lb;
    const long long m = 2 * (long long)n;
    la.reserve((size_t)m, s); lb.reserve((size_t)m, s);
    tour_succ_kernel<<<nblk(n), 256, 0, s>>>(parent.p, n, cptr.p, child.p, slot.p, la.p);
    h.stats.kernels_launched++;
    // the head of the tour: "enter" of the first root (the children of the virtual vertex n)
    int first_root = 0;
    {
      int cp[1];
      read_ints(h, cptr.p + n, cp, 1);
      read_ints(h, child.p + cp[0], &first_root, 1);
    }
    const unsigned head = 2u * (unsigned)first_root;
    DevBuf<int> vflag, vslot;
    vflag.reserve((size_t)n + 1, s); vslot.reserve((size_t)n + 2, s);
    tour_flag_kernel<<<nblk(n), 256, 0, s>>>(n, vflag.p);
    h.stats.kernels_launched += 1 + exclusive_scan<0>(vflag.p, vslot.p, n, h.scan_ws.p, s);
    int n_sampled = 0;
    read_ints(h, vslot.p + n, &n_sampled, 1);
    const unsigned n_reg = 2u * (unsigned)n_sampled, n_spl = n_reg + 1;
    DevBuf<unsigned long long> sa, sb;
    sa.reserve(n_spl, s); sb.reserve(n_spl, s);
    tour_segment_kernel<<<nblk(m + 1), 256, 0, s>>>(la.p, m, head, n_reg, vslot.p, sa.p);
    unsigned long long *pc = sa.p, *pn = sb.p;
    int rounds = 1;
    while ((1u << rounds) < n_spl) rounds++;
    for (int r = 0; r < rounds; r++) {
      tour_jump_fixed_kernel<<<nblk(n_spl), 256, 0, s>>>(pc, pn, n_spl);
      std::swap(pc, pn);
    }
    tour_spread_kernel<<<nblk(m), 256, 0, s>>>(la.p, m, head, n_reg, vslot.p, pc, lb.p);
    h.stats.kernels_launched += 2 + rounds;
    unsigned long long *cur = lb.p;
    tour_extract_kernel<<<nblk(n), 256, 0, s>>>(cur, n, tin.p, size.p);
    h.stats.kernels_launched++;
    PB_CUDA(cudaStreamSynchronize(s));
  }

  // ---- binary lifting table (levels x n)
  int levels = 1;
  DevBuf<int> up;
  {
    // grow level by level until no 2^b-th ancestor exists; capacity for the worst case is only
    // reserved as needed (n ints per level)
    std::vector<int *> lv;
    size_t cap_levels = 24;
    up.reserve((size_t)cap_levels * n, s);
    lift_init_kernel<<<nblk(n), 256, 0, s>>>(parent.p, n, up.p);
    h.stats.kernels_launched++;
    for (;;) {
      if ((size_t)levels == cap_levels) {
        DevBuf<int> bigger;
        bigger.reserve((size_t)(cap_levels + 8) * n, s);
        PB_CUDA(cudaMemcpyAsync(bigger.p, up.p, sizeof(int) * (size_t)levels * n, cudaMemcpyDeviceToDevice, s));
        PB_CUDA(cudaStreamSynchronize(s));
        up.swap(bigger);
        cap_levels += 8;
      }
      PB_CUDA(cudaMemsetAsync(flag.p + 3, 0, sizeof(int), s));
      lift_step_kernel<<<nblk(n), 256, 0, s>>>(up.p + (size_t)(levels - 1) * n, n, up.p + (size_t)levels * n, flag.p + 3);
      h.stats.kernels_launched++;
      int any = 0;
      read_ints(h, flag.p + 3, &any, 1);
      if (!any) break;
      levels++;
    }
  }

  // ---- deltas, interval sums
  DevBuf<int> delta, bytin, pre, cc;
  DevBuf<unsigned long long> lnz;
  delta.reserve((size_t)n + 1, s); bytin.reserve((size_t)n + 1, s); pre.reserve((size_t)n + 2, s); cc.reserve((size_t)n + 1, s);
  lnz.reserve(1, s);
  PB_CUDA(cudaMemsetAsync(lnz.p, 0, sizeof(unsigned long long), s));
  delta_base_kernel<<<nblk(n), 256, 0, s>>>(cptr.p, n, delta.p);
  delta_rows_kernel<<<nblk(n), 256, 0, s>>>(perm.p, inv.p, h.Mp, h.Mi, n, tin.p, size.p, up.p, levels, delta.p);
  scatter_by_tin_kernel<<<nblk(n), 256, 0, s>>>(delta.p, tin.p, n, bytin.p);
  h.stats.kernels_launched += 3 + exclusive_scan<0>(bytin.p, pre.p, n, h.scan_ws.p, s);
  colcount_kernel<<<nblk(n), 256, 0, s>>>(pre.p, tin.p, size.p, n, cc.p, lnz.p);
  h.stats.kernels_launched++;
  unsigned long long total = 0;
  PB_CUDA(cudaMemcpyAsync(&total, lnz.p, sizeof(total), cudaMemcpyDeviceToHost, s));
  if (colcount_out) {
    PB_CUDA(cudaMemcpyAsync(colcount_out, cc.p, sizeof(int) * (size_t)n, out_kind, s));
    h.stats.d2h_bytes += (long long)sizeof(int) * n;
  }
  PB_CUDA(cudaStreamSynchronize(s));
  *lnz_out = (long long)total;
  tm.stop();
}

}  // namespace pb
