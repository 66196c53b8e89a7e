// parth_b200/csrc/stage_maps.cu -- DOF-map stages on the flat separator tree:
//   (a3) ParthAPI::setNewToOldDOFMap / computeOldToNewDOFMap   (reference ParthAPI.cpp:126-189)
//   (a6) Integrator::integrateDOFChanges                         (Integrator.cpp:44-257)
//   (f2) Integrator::relocateDOFsForAggressiveReuse              (Integrator.cpp:377-544)
//
// Both tree edits are "some entries leave their node, some DOFs join a node" and share one
// device formulation ("regroup"):
//   * label compaction inside a node (the reference's O(deleted x node) loop, Integrator.cpp:86-91)
//     = exclusive scan of "deleted" flags laid out in LABEL space: new = old - #deleted labels below
//   * surviving entries keep their relative order -> stream compaction by an exclusive scan
//   * joining DOFs are appended in ascending id with labels size, size+1, ..., then the
//     reference sorts DOFs but not the labels (Integrator.cpp:241-244, 536-539) -> dofs[] is the
//     sorted MERGE of survivors and joiners while labels[] stays positional
//   * offsets: one post-order walk over <= 8191 nodes on the host (HMD.cpp:681-697)
// Added-DOF placement (Integrator.cpp:118-202) is an order-dependent FIFO; it is reproduced
// exactly by a wavefront schedule: within one pass a DOF waits for its still-unplaced added
// neighbours with a smaller id (they precede it in the queue), sees the placements of earlier
// passes, and ignores later-queue neighbours even if they were placed in the same pass.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

#include "scan.cuh"
#include "stages.cuh"
#include "tree.cuh"

namespace cg = cooperative_groups;

namespace pb {

namespace {

constexpr int kNoInit = 1 << 30;  // "no relocation target" (reference: num_HMD_nodes + 1)

__device__ __forceinline__ int node_of_entry_d(const int *__restrict__ node_ptr, int nodes, int k) {
  int lo = 0, hi = nodes;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(node_ptr + mid) <= k) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

__device__ __forceinline__ bool is_sep_d(int sep, int node) {
  if (sep > node) return false;
  int a = (node - 1) / 2;
  while (a > sep && a > 0) a = (a - 1) / 2;
  return a == sep;
}

// heap-consistent tree (parent == (i-1)/2 and level == depth for every node, which is what
// HMD::Decompose produces, HMD.cpp:303-337): findCommonSeparator in pure integer arithmetic
__device__ __forceinline__ int lca_heap_d(int a, int b) {
  int da = 31 - __clz(a + 1), db = 31 - __clz(b + 1);
  while (da > db) { a = (a - 1) >> 1; da--; }
  while (db > da) { b = (b - 1) >> 1; db--; }
  while (a != b) { a = (a - 1) >> 1; b = (b - 1) >> 1; }
  return a;
}

__device__ __forceinline__ int lca_d(const int *__restrict__ meta, int first, int second) {
  if (meta == nullptr) return lca_heap_d(first, second);
  int hi = first, lo = second;
  if (__ldg(meta + hi * 7 + 5) < __ldg(meta + lo * 7 + 5)) { hi = second; lo = first; }
  while (__ldg(meta + hi * 7 + 5) != __ldg(meta + lo * 7 + 5)) hi = __ldg(meta + hi * 7 + 3);
  while (hi != lo) { hi = __ldg(meta + hi * 7 + 3); lo = __ldg(meta + lo * 7 + 3); }
  return hi;
}

// ---------------------------------------------------------------- (a3) map inversion
__global__ void fill_kernel(int *__restrict__ a, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__global__ void invert_map_kernel(const int *__restrict__ n2o, int M, int M_prev, int *__restrict__ o2n,
                                  int *__restrict__ bad) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= M) return;
  int o = n2o[n];
  if (o == -1) return;
  if ((unsigned)o >= (unsigned)M_prev) { *bad = 1; return; }
  o2n[o] = n;
}

// ascending list of the positions i with a[i] == -1, given the exclusive scan of that predicate
__global__ void list_minus_one_kernel(const int *__restrict__ a, const int *__restrict__ pre, int n,
                                      int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i] == -1) out[pre[i]] = i;
}

// ---------------------------------------------------------------- regroup, step 1
// MODE 0 (integrate): v = old_to_new[dof], stays iff v != -1;   MODE 1 (relocate): v = dof, stays
// iff target[dof] == kNoInit.  keep[k] in entry space, del[] in label space; integrate also
// seeds nd2n[v] = node for the entries that stay.
template <int MODE>
__global__ void regroup_flags_kernel(const int *__restrict__ node_ptr, int nodes, const int *__restrict__ dofs,
                                     const int *__restrict__ labels, int total, const int *__restrict__ key,
                                     int *__restrict__ keep, int *__restrict__ del, int *__restrict__ nd2n) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  const int nd = node_of_entry_d(node_ptr, nodes, k);
  const int d = dofs[k];
  int stays;
  if (MODE == 0) {
    const int v = __ldg(key + d);
    stays = v != -1;
    if (stays) nd2n[v] = nd;
  } else {
    stays = __ldg(key + d) == kNoInit;
  }
  keep[k] = stays;
  del[__ldg(node_ptr + nd) + labels[k]] = !stays;
}

__global__ void gather_kernel(const int *__restrict__ src, const int *__restrict__ idx, int n, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

// step 2: compacted survivors (ids and compacted labels) in entry order
template <int MODE>
__global__ void regroup_compact_kernel(const int *__restrict__ node_ptr, int nodes, const int *__restrict__ dofs,
                                       const int *__restrict__ labels, int total, const int *__restrict__ key,
                                       const int *__restrict__ S_keep, const int *__restrict__ S_del,
                                       int *__restrict__ survC, int *__restrict__ labC) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  if (S_keep[k + 1] == S_keep[k]) return;
  const int nd = node_of_entry_d(node_ptr, nodes, k);
  const int a = __ldg(node_ptr + nd);
  const int lab = labels[k];
  const int t = S_keep[k];
  survC[t] = MODE == 0 ? __ldg(key + dofs[k]) : dofs[k];
  labC[t] = lab - (S_del[a + lab] - S_del[a]);
}

// ---------------------------------------------------------------- joiners: per-node rank
// rank = position among the joiners of the same node in the (ascending) joiner list; joinS[aptr[node] + rank] = dof.
// The list is cut into segments: CTA (t, g) counts the joiners of touched node t in segment g, a tiny scan turns the
// counts into segment bases, and the same grid walks its segments again with a block scan per 256 entries.  (One CTA
// per node over the whole list took 0.27 ms for the 97 000 joiners of config 3.)
__global__ void __launch_bounds__(256)
rank_joiners_count_kernel(const int *__restrict__ touched, const int *__restrict__ join_node, int n_join, int seg,
                          int *__restrict__ segcnt) {
  const int node = touched[blockIdx.x];
  const int lo = blockIdx.y * seg, hi = min(n_join, lo + seg);
  int c = 0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) c += __ldg(join_node + i) == node;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int s_w[8];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; k++) t += s_w[k];
    segcnt[blockIdx.x * gridDim.y + blockIdx.y] = t;
  }
}

__global__ void rank_joiners_scan_kernel(int *__restrict__ segcnt, int n_touched, int n_seg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_touched) return;
  int run = 0;
  for (int g = 0; g < n_seg; g++) {
    const int c = segcnt[t * n_seg + g];
    segcnt[t * n_seg + g] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(256)
rank_joiners_kernel(const int *__restrict__ touched, const int *__restrict__ join, const int *__restrict__ join_node,
                    int n_join, int seg, const int *__restrict__ segbase, const int *__restrict__ aptr,
                    int *__restrict__ joinS) {
  __shared__ int s_scan[33];
  const int node = touched[blockIdx.x];
  const int base = aptr[node] + (segbase ? segbase[blockIdx.x * gridDim.y + blockIdx.y] : 0);
  const int lo = blockIdx.y * seg, hi = min(n_join, lo + seg);
  int running = 0;
  for (int i0 = lo; i0 < hi; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const int f = i < hi && __ldg(join_node + i) == node;
    int total;
    const int ex = block_excl_scan<256>(f, s_scan, &total);
    if (f) joinS[base + running + ex] = join[i];
    running += total;
  }
}

__global__ void count_joiners_kernel(const int *__restrict__ join_node, int n_join, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_join) atomicAdd(cnt + join_node[i], 1);
}

__device__ __forceinline__ int lower_bound_d(const int *__restrict__ a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// step 3a: survivors into the merged node slices
__global__ void regroup_place_surv_kernel(const int *__restrict__ sptr, const int *__restrict__ aptr,
                                          const int *__restrict__ nptr, int nodes, int total_surv,
                                          const int *__restrict__ survC, const int *__restrict__ labC,
                                          const int *__restrict__ joinS, int *__restrict__ dofs2,
                                          int *__restrict__ labels2, int *__restrict__ unsorted) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_surv) return;
  const int nd = node_of_entry_d(sptr, nodes, t);
  const int j = t - __ldg(sptr + nd);
  const int v = survC[t];
  const int na = __ldg(aptr + nd + 1) - __ldg(aptr + nd);
  int shift = 0;
  if (na > 0) {
    shift = lower_bound_d(joinS + __ldg(aptr + nd), na, v);
    if (j > 0 && survC[t - 1] > v) *unsorted = 1;  // the merge needs ascending survivors
  }
  const int base = __ldg(nptr + nd);
  dofs2[base + j + shift] = v;
  labels2[base + j] = labC[t];
}

// step 3b: joiners
__global__ void regroup_place_join_kernel(const int *__restrict__ sptr, const int *__restrict__ aptr,
                                          const int *__restrict__ nptr, int nodes, int n_join,
                                          const int *__restrict__ survC, const int *__restrict__ joinS,
                                          int *__restrict__ dofs2, int *__restrict__ labels2) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_join) return;
  const int nd = node_of_entry_d(aptr, nodes, u);
  const int r = u - __ldg(aptr + nd);
  const int v = joinS[u];
  const int s0 = __ldg(sptr + nd), ns = __ldg(sptr + nd + 1) - s0;
  const int shift = lower_bound_d(survC + s0, ns, v);
  const int base = __ldg(nptr + nd);
  dofs2[base + r + shift] = v;
  labels2[base + ns + r] = ns + r;
}

// ---------------------------------------------------------------- (a6) added-DOF placement
// Exact wavefront schedule of the reference's FIFO placement (Integrator.cpp:118-202).  Inside one
// pass an added DOF d depends on the still-unplaced added neighbours with a smaller id (they precede
// it in the queue); this is a DAG.  pass_init counts those predecessors and seeds the frontier with
// the DOFs that have none; every sweep decides the DOFs of the current frontier (they only look at
// decisions of earlier sweeps, i.e. earlier launches), releases their successors and appends the
// ones that became ready to the next frontier.  Work per pass = edges among the unplaced DOFs.
// state: placed_pass[d] (0 = not placed), indeg[d]; cnt[s] = size of the frontier of sweep s.
struct PlaceCounters { int unplaced; int placed_in_pass; int pad0, pad1; };

__global__ void place_pass_init_kernel(const int *__restrict__ added, int n_added, const int *__restrict__ Mp,
                                       const int *__restrict__ Mi, const int *__restrict__ n2o,
                                       const int *__restrict__ placed_pass, int *__restrict__ indeg,
                                       int *__restrict__ frontier, int *__restrict__ fcnt, PlaceCounters *cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_added) return;
  const int d = added[q];
  if (placed_pass[d] != 0) return;
  atomicAdd(&cnt->unplaced, 1);
  int deg = 0;
  for (int p = Mp[d]; p < Mp[d + 1]; p++) {
    const int nb = Mi[p];
    deg += nb < d && __ldg(n2o + nb) == -1 && placed_pass[nb] == 0;
  }
  indeg[d] = deg;
  if (deg == 0) frontier[atomicAdd(fcnt, 1)] = d;
}

// Decision of ONE ready DOF d by ONE thread: the node it joins, or -2 when it has no visible neighbour (it stays
// unplaced: next ring).  The arrays other CTAs write during the pass (nd2n, placed_pass) are read through L2
// (ld.global.cg): the dataflow kernel keeps running while they change, so a line cached in L1 earlier must not be
// served again.
__device__ __forceinline__ int place_decide_one(int d, const int *__restrict__ Mp, const int *__restrict__ Mi,
                                                const int *__restrict__ n2o, const int *__restrict__ meta,
                                                const int *nd2n, const int *placed_pass, int pass) {
  const int s = Mp[d], e = Mp[d + 1];
  // visible node of a neighbour, -1 if none: survivors always; added DOFs placed in an earlier
  // pass, or in this pass if they precede d in the queue (then they were decided before d became ready)
  auto visible = [&](int nb) -> int {
    if (__ldg(n2o + nb) != -1) return __ldcg(nd2n + nb);
    const int pp = __ldcg(placed_pass + nb);
    if (pp != 0 && (pp < pass || nb < d)) return __ldcg(nd2n + nb);
    return -1;
  };
  // distinct visible neighbour nodes (a mesh vertex touches a handful of nodes at most)
  constexpr int kMaxDistinct = 16;
  int uniq[kMaxDistinct];
  int nu = 0;
  bool overflow = false;
  for (int p = s; p < e; p++) {
    const int r = visible(Mi[p]);
    if (r == -1) continue;
    bool seen = false;
    for (int u = 0; u < nu; u++) seen |= uniq[u] == r;
    if (seen) continue;
    if (nu < kMaxDistinct) uniq[nu++] = r; else overflow = true;
  }
  if (nu == 0) return -2;
  if (nu == 1 && !overflow) return uniq[0];
  // minimum over all pairs of distinct neighbour nodes of {descendant if ancestor-related, else LCA}
  int mn = 1 << 30;
  auto pair_node = [&](int ra, int rb) {
    const int big = ra > rb ? ra : rb, small = ra > rb ? rb : ra;
    const int c = is_sep_d(small, big) ? big : lca_d(meta, small, big);
    mn = c < mn ? c : mn;
  };
  if (!overflow) {
    for (int a = 0; a < nu; a++)
      for (int b2 = a + 1; b2 < nu; b2++) pair_node(uniq[a], uniq[b2]);
  } else {
    for (int p = s; p < e; p++) {
      const int ra = visible(Mi[p]);
      if (ra == -1) continue;
      for (int p2 = p + 1; p2 < e; p2++) {
        const int rb = visible(Mi[p2]);
        if (rb != -1 && rb != ra) pair_node(ra, rb);
      }
    }
  }
  return mn;
}

// the successors of d (added neighbours with a larger id that were unplaced at pass start) lose one predecessor;
// the ones that became ready are appended to the next frontier (launch-per-sweep path)
__device__ __forceinline__ void place_release_one(int d, int *__restrict__ next, int *__restrict__ fcnt_out,
                                                  const int *__restrict__ Mp, const int *__restrict__ Mi,
                                                  const int *__restrict__ n2o, const int *placed_pass, int *indeg,
                                                  int pass) {
  for (int p = Mp[d]; p < Mp[d + 1]; p++) {
    const int nb = Mi[p];
    if (nb <= d || __ldg(n2o + nb) != -1) continue;
    const int pp = __ldcg(placed_pass + nb);
    if (pp != 0 && pp != pass) continue;            // placed in an earlier pass: not in this DAG
    if (atomicSub(indeg + nb, 1) == 1) next[atomicAdd(fcnt_out, 1)] = nb;
  }
}

// launch-per-sweep path (kept for wavefronts of millions of DOFs and as the cross-check of the dataflow kernel)
__global__ void place_sweep_kernel(const int *__restrict__ frontier, const int *__restrict__ fcnt_in,
                                   int *__restrict__ next, int *__restrict__ fcnt_out, const int *__restrict__ Mp,
                                   const int *__restrict__ Mi, const int *__restrict__ n2o,
                                   const int *__restrict__ meta, int *nd2n, int *placed_pass, int *indeg, int pass,
                                   PlaceCounters *cnt) {
  const int nf = *fcnt_in;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nf; q += gridDim.x * blockDim.x) {
    const int d = frontier[q];
    const int r = place_decide_one(d, Mp, Mi, n2o, meta, nd2n, placed_pass, pass);
    // the successors' test "unplaced at pass start" sees d as unplaced or placed in THIS pass: both pass it
    place_release_one(d, next, fcnt_out, Mp, Mi, n2o, placed_pass, indeg, pass);
    if (r != -2) {
      __stcg(nd2n + d, r);
      __stcg(placed_pass + d, pass);
      atomicAdd(&cnt->placed_in_pass, 1);
    }
  }
}

// ---- pull placement: ONE launch per pass, no barrier between wavefronts, no fences.
// The wavefront schedule is a DAG (a DOF waits for its unplaced added neighbours with a smaller id).  Walking it sweep
// by sweep costs a barrier (or a launch) per wavefront and leaves the GPU idle: a remeshed patch has ~3 x patch-edge
// wavefronts of a few hundred DOFs (config 3: 136 sweeps).  History on that config (profiles/r08_placement.md): launch
// per sweep 4.6 ms; thread-block cluster + hardware barrier per sweep 1.6 ms; push queue with fenced hand-over 1.7 ms
// (three device-scope fences per hop); this kernel 0.3 ms.
//
// The whole state of a DOF during placement lives in ONE 32-bit word,
//   st[v] = pass << 16 | (node + 1)   placed in `pass` (survivors: pass 0);   pass << 16   decided in `pass`: stays
//   unplaced;   0   added, never decided,
// so a reader needs no ordering between locations: it polls the words of the predecessors it waits for and the value it
// finally reads IS the decision (relaxed device-scope loads and stores, measured 430 ns store -> visible on B200,
// scripts/gmem_latency_probe.cu).  One warp per DOF, one lane per neighbour (the neighbours' words are fetched by one
// load instruction: a dependent load costs 155 ns at L2).  The ORDER in which the warps take the DOFs comes from a
// ready queue: a DOF is pushed when its last predecessor has STARTED (relaxed atomic countdown), tickets are dealt
// statically (warp w takes slots w, w + W, ...), so the warps follow the wavefronts whatever the numbering of the mesh
// (owning the DOFs in id order serialises the walk plane by plane: 3.4 ms) and a successor has its adjacency and its
// other predecessors' words in registers by the time the last decision it waits for lands.  The queue is only a
// schedule -- correctness rests on the polled words alone.  Progress: a popped DOF always has a warp; among the popped,
// undecided DOFs the one with the smallest id waits for nobody that is not decided or being decided, and the smallest
// unfilled slot gets filled (DAG) -- provided every warp is resident, hence grid <= SM count.  One warp per DOF, not
// two half-warps: two groups in one warp time-share its issue slot while either of them spins (measured 2x slower).
constexpr int kPullThreads = 512;

// relaxed device-scope accesses for the words other CTAs poll (volatile would make them system-scope)
__device__ __forceinline__ int ld_poll(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_poll(int *p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// visible node of neighbour nb for the DOF d being decided in `pass` (-1: none), given a first read v of its word;
// waits for an undecided predecessor
__device__ __forceinline__ int pull_visible(int v, int nb, int d, const int *st, int pass) {
  for (;;) {
    const int code = v & 0xffff, p = v >> 16;
    if (code) return (p < pass || nb < d) ? code - 1 : -1;
    if (p == pass || nb > d) return -1;  // decided "stays unplaced" in this pass / later in the queue: invisible
    v = ld_poll(st + nb);
  }
}

__device__ __forceinline__ int pair_node_d(const int *__restrict__ meta, int ra, int rb) {
  const int big = ra > rb ? ra : rb, small = ra > rb ? rb : ra;
  return is_sep_d(small, big) ? big : lca_d(meta, small, big);
}

// kPullQueues ready queues (a DOF goes to queue id % kPullQueues; its segment of the slot array starts at qoff[q] and
// holds qoff[q + 1] - qoff[q] = the number of unplaced DOFs of that residue, known at pass start): the push counter
// is the one same-address atomic on the path of every DOF, and same-address atomics serialise at L2.
constexpr int kPullQueues = 16, kTailStride = 32;  // one tail per 128-byte line

__device__ __forceinline__ void pull_push(int *queue, int *tails, const int *qoff, int x) {
  const int q = x & (kPullQueues - 1);
  st_poll(queue + __ldg(qoff + q) + atomicAdd(tails + q * kTailStride, 1), x);
}

__global__ void __launch_bounds__(kPullThreads)
place_pass_pull_kernel(int *queue, int *tails, const int *__restrict__ qoff, const int *__restrict__ Mp,
                       const int *__restrict__ Mi, const int *__restrict__ meta, int *st, int *indeg, int pass,
                       PlaceCounters *cnt) {
  const int lane = threadIdx.x & 31;
  // warp w serves queue w % kPullQueues, slots w / kPullQueues, + (warps per queue), ... (CTA-interleaved: consecutive
  // slots of a queue go to different SMs)
  const int w = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nwarps = gridDim.x * (kPullThreads / 32);
  const int q = w & (kPullQueues - 1), per_q = nwarps / kPullQueues;
  const int q0 = __ldg(qoff + q), total = __ldg(qoff + q + 1) - q0;
  int placed = 0;
  for (int t = w / kPullQueues; t < total && w < per_q * kPullQueues; t += per_q) {
    int d = 0;
    if (lane == 0)
      while ((d = ld_poll(queue + q0 + t)) < 0) {}
    d = __shfl_sync(0xffffffffu, d, 0);
    const int s = __ldg(Mp + d), deg = __ldg(Mp + d + 1) - s;
    int code = 0;
    if (deg > 32) {  // wide row: one thread walks it, neighbour by neighbour (same rule)
      if (lane == 0) {
        for (int p = s; p < s + deg; p++) {  // successors first: they fetch their rows while this one is decided
          const int x = __ldg(Mi + p);
          if (x > d && (ld_poll(st + x) & 0xffff) == 0 && atomicSub(indeg + x, 1) == 1) pull_push(queue, tails, qoff, x);
        }
        int mn = 1 << 30, first = -1;
        bool two = false;
        for (int p = s; p < s + deg; p++) {
          const int a = __ldg(Mi + p), ra = pull_visible(ld_poll(st + a), a, d, st, pass);
          if (ra == -1) continue;
          if (first == -1) first = ra;
          for (int p2 = p + 1; p2 < s + deg; p2++) {
            const int b2 = __ldg(Mi + p2), rb = pull_visible(ld_poll(st + b2), b2, d, st, pass);
            if (rb == -1 || rb == ra) continue;
            const int c = pair_node_d(meta, ra, rb);
            mn = c < mn ? c : mn;
            two = true;
          }
        }
        code = first == -1 ? 0 : (two ? mn : first) + 1;
      }
    } else {
      const int nb = lane < deg ? __ldg(Mi + s + lane) : -1;
      const int v = nb >= 0 ? ld_poll(st + nb) : (1 << 16);
      // release the successors (added neighbours with a larger id, unplaced at pass start) BEFORE deciding
      const bool ready = nb > d && (v & 0xffff) == 0 && atomicSub(indeg + nb, 1) == 1;
      if (ready) pull_push(queue, tails, qoff, nb);
      const int r = nb >= 0 ? pull_visible(v, nb, d, st, pass) : -1;
      // distinct visible nodes by match_any; minimum over all pairs of distinct nodes by a shuffle loop over the leaders
      const unsigned peers = __match_any_sync(0xffffffffu, r);
      const bool leader = r != -1 && (__ffs(peers) - 1) == lane;
      const unsigned um = __ballot_sync(0xffffffffu, leader);
      const int nu = __popc(um);
      if (nu == 1) code = __shfl_sync(0xffffffffu, r, __ffs(um) - 1) + 1;
      else if (nu > 1) {
        int mn = 1 << 30;
        for (unsigned m = um; m; m &= m - 1) {
          const int k = __ffs(m) - 1;
          const int rk = __shfl_sync(0xffffffffu, r, k);
          if (leader && k > lane) {
            const int c = pair_node_d(meta, r, rk);
            mn = c < mn ? c : mn;
          }
        }
        code = __reduce_min_sync(0xffffffffu, mn) + 1;
      }
    }
    if (lane == 0) {  // code 0: no visible neighbour, stays unplaced (next ring)
      st_poll(st + d, (pass << 16) | code);
      placed += code != 0;
    }
    __syncwarp();
  }
  placed = __reduce_add_sync(0xffffffffu, placed);
  if (lane == 0 && placed) atomicAdd(&cnt->placed_in_pass, placed);
}

// st = node + 1 for the survivors (pass 0), 0 for the added DOFs (nd2n holds -1 for them)
__global__ void place_state_init_kernel(const int *__restrict__ nd2n, int n, int *__restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st[i] = nd2n[i] + 1;
}

// pass start of the pull path, three small launches: (1) predecessor counts + unplaced DOFs per queue, (2) queue
// segment offsets, (3) the DOFs without predecessor enter their queues
__global__ void place_pull_count_kernel(const int *__restrict__ added, int n_added, const int *__restrict__ Mp,
                                        const int *__restrict__ Mi, const int *__restrict__ st, int *__restrict__ indeg,
                                        int *__restrict__ qcount) {
  __shared__ int hist[kPullQueues];
  if (threadIdx.x < kPullQueues) hist[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_added && (st[added[i]] & 0xffff) == 0) {
    const int d = added[i];
    int deg = 0;
    for (int p = Mp[d]; p < Mp[d + 1]; p++) {
      const int nb = Mi[p];
      deg += nb < d && (st[nb] & 0xffff) == 0;
    }
    indeg[d] = deg;
    atomicAdd(hist + (d & (kPullQueues - 1)), 1);
  }
  __syncthreads();
  if (threadIdx.x < kPullQueues && hist[threadIdx.x]) atomicAdd(qcount + threadIdx.x, hist[threadIdx.x]);
}

__global__ void place_pull_offsets_kernel(const int *__restrict__ qcount, int *__restrict__ qoff, PlaceCounters *cnt) {
  if (threadIdx.x == 0) {
    int run = 0;
    for (int q = 0; q < kPullQueues; q++) { qoff[q] = run; run += qcount[q]; }
    qoff[kPullQueues] = run;
    cnt->unplaced = run;
  }
}

__global__ void place_pull_roots_kernel(const int *__restrict__ added, int n_added, const int *__restrict__ st,
                                        const int *__restrict__ indeg, int *__restrict__ queue, int *__restrict__ tails,
                                        const int *__restrict__ qoff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_added) return;
  const int d = added[i];
  if ((st[d] & 0xffff) == 0 && indeg[d] == 0) pull_push(queue, tails, qoff, d);
}

// pull path epilogue: the node of every added DOF out of its state word
__global__ void place_decode_kernel(const int *__restrict__ added, int n_added, const int *__restrict__ st,
                                    int *__restrict__ nd2n) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_added) return;
  const int d = added[q], code = st[d] & 0xffff;
  nd2n[d] = code ? code - 1 : -1;
}

__global__ void finish_added_kernel(const int *__restrict__ added, int n_added, int spare, int *__restrict__ nd2n,
                                    int *__restrict__ join_node) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_added) return;
  const int d = added[q];
  int nd = nd2n[d];
  if (nd == -1) { nd = spare; nd2n[d] = spare; }
  join_node[q] = nd;
}

// ---------------------------------------------------------------- (f2) relocation targets
__global__ void occ_kernel(const int *__restrict__ edges, int n_edges, int *__restrict__ occ) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  atomicAdd(occ + edges[2 * e], 1);
  atomicAdd(occ + edges[2 * e + 1], 1);
}

__global__ void relocate_targets_kernel(const int *__restrict__ edges, int n_edges, const int *__restrict__ dof2node,
                                        const int *__restrict__ meta, const int *__restrict__ occ, int threshold,
                                        int *__restrict__ target, int *__restrict__ keep_edge) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int a = edges[2 * e], b = edges[2 * e + 1];
  const int na = __ldg(dof2node + a), nb = __ldg(dof2node + b);
  const int big = na > nb ? na : nb, small = na > nb ? nb : na;
  int keep = 0;
  if (small == big) keep = 1;
  else if (!is_sep_d(small, big)) {
    const int cs = lca_d(meta, na, nb);
    if (cs < threshold) atomicMin(target + (occ[a] > occ[b] ? a : b), cs);
    else keep = 1;
  }
  keep_edge[e] = keep;
}

__global__ void compact_edges_kernel(const int *__restrict__ edges, int n_edges, const int *__restrict__ keep_edge,
                                     const int *__restrict__ pre, int *__restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges || !keep_edge[e]) return;
  out[2 * pre[e]] = edges[2 * e];
  out[2 * pre[e] + 1] = edges[2 * e + 1];
}

__global__ void moved_flags_kernel(const int *__restrict__ target, int M, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) flag[i] = target[i] != kNoInit;
}

__global__ void list_moved_kernel(const int *__restrict__ flag, const int *__restrict__ pre, const int *__restrict__ target,
                                  int M, int *__restrict__ moved, int *__restrict__ moved_node, int *__restrict__ dof2node) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M || !flag[i]) return;
  moved[pre[i]] = i;
  moved_node[pre[i]] = target[i];
  dof2node[i] = target[i];
}

inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

// ---------------------------------------------------------------- (f2) relocation, sparse formulation
// A reuse update moves a few hundred DOFs; the dense formulation above spends a dozen passes over M ints on them.
// Here the per-DOF scratch (occurrence counts, relocation targets) lives in two PERSISTENT arrays that are clean
// between calls and are reset entry by entry afterwards, the moved DOFs are collected from the edge endpoints and
// sorted by one CTA, and only the nodes that lose or receive DOFs are regrouped (one CTA each); the slices of all
// other nodes are copied to their new place in one streaming pass.  Same arrays as regroup_finish<1>, bit for bit.
constexpr int kSparseEdges = 8192;            // changed edges the sparse path takes (2 * kSparseEdges candidates)
constexpr int kSparseCand = 2 * kSparseEdges;

// The whole front half of the sparse relocation in ONE launch of ONE CTA (<= 8192 changed edges): occurrence counts,
// relocation targets, ordered compaction of the edges that stay, collection + ascending sort of the moved DOFs, per-node
// leaver / joiner counts, redirect of dof2node, reset of the occurrence scratch, and (filter) the kept edges written
// back over the edge list.  As eleven stream operations of 3-5 us each the same work was launch-bound (~45 us of the
// headline update); the phases only need CTA barriers.  out: [0] n_moved, [1] unsorted, [2] size mismatch, [3] kept,
// [4..5] a copy of sticky_in (the regroup kernels' error flags).
// Dynamic shared memory: kSparseCand ints (the moved list / sort buffer) + kSparseEdges bytes (keep flags).
__global__ void apply_layout_kernel(const int *__restrict__ nptr, const int *__restrict__ off, int nn,
                                    int *__restrict__ node_ptr, int *__restrict__ meta) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x <= nn) node_ptr[x] = nptr[x];
  if (x < nn) meta[(size_t)x * 7 + 4] = off[x];
}

__global__ void __launch_bounds__(1024)
reloc_front_kernel(int *__restrict__ edges, int ne, int *__restrict__ dof2node, const int *__restrict__ meta, int threshold,
                   int *__restrict__ occ, int *__restrict__ target, int *__restrict__ kept_edges, int *__restrict__ moved,
                   int *__restrict__ moved_node, int *__restrict__ cnt, int nn, int *__restrict__ out, int filter,
                   const int *__restrict__ sticky_in) {
  extern __shared__ int s[];
  unsigned char *keep = reinterpret_cast<unsigned char *>(s + kSparseCand);
  __shared__ int s_scan[33];
  __shared__ int s_n;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * nn; i += 1024) cnt[i] = 0;
  if (tid < 4) out[tid] = 0;
  else if (tid < 6) out[tid] = sticky_in[tid - 4];  // error flags of earlier stream-ordered updates ride along
  if (tid == 0) s_n = 0;
  for (int q = tid; q < 2 * ne; q += 1024) atomicAdd(occ + edges[q], 1);
  __syncthreads();
  for (int e = tid; e < ne; e += 1024) {
    const int a = edges[2 * e], b = edges[2 * e + 1];
    const int na = dof2node[a], nb = dof2node[b];
    const int big = na > nb ? na : nb, small = na > nb ? nb : na;
    int k = 0;
    if (small == big) k = 1;
    else if (!is_sep_d(small, big)) {
      const int cs = lca_d(meta, na, nb);
      if (cs < threshold) atomicMin(target + (occ[a] > occ[b] ? a : b), cs);
      else k = 1;
    }
    keep[e] = (unsigned char)k;
  }
  __syncthreads();
  // kept edges, in order (a thread owns 8 consecutive edges of a 8192-edge tile: one tile)
  {
    int c = 0;
    const int e0 = tid * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) c += e0 + i < ne ? keep[e0 + i] : 0;
    int total;
    int pos = block_excl_scan<1024>(c, s_scan, &total);
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (e0 + i < ne && keep[e0 + i]) {
        kept_edges[2 * pos] = edges[2 * (e0 + i)];
        kept_edges[2 * pos + 1] = edges[2 * (e0 + i) + 1];
        pos++;
      }
    if (tid == 0) out[3] = total;
  }
  // moved DOFs: every end point with a target, listed by its first visitor
  for (int q = tid; q < 2 * ne; q += 1024) {
    const int d = edges[q];
    if (target[d] == kNoInit) continue;
    if (atomicExch(occ + d, (int)0x80000000) >= 0) s[atomicAdd(&s_n, 1)] = d;
  }
  __syncthreads();
  const int n = s_n;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + tid; i < P; i += 1024) s[i] = 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const int x = s[i], y = s[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { s[i] = y; s[l] = x; }
        }
      }
      __syncthreads();
    }
  // (before dof2node is redirected: the nodes that lose DOFs are the old ones)
  for (int i = tid; i < n; i += 1024) {
    const int d = s[i], t = target[d];
    atomicAdd(cnt + dof2node[d], 1);
    atomicAdd(cnt + nn + t, 1);
    moved[i] = d;
    moved_node[i] = t;
    dof2node[d] = t;
  }
  for (int q = tid; q < 2 * ne; q += 1024) occ[edges[q]] = 0;
  if (tid == 0) out[0] = n;
  __syncthreads();
  if (filter) {
    const int kept = out[3];
    for (int q = tid; q < 2 * kept; q += 1024) edges[q] = kept_edges[q];
  }
}

// the slices of the nodes nothing happens to move to their new offsets (8 consecutive entries per thread: one node
// look-up for the run unless it crosses a node boundary)
constexpr int RC_RUN = 8;
__global__ void regroup_copy_kernel(const int *__restrict__ optr, const int *__restrict__ nptr,
                                    const int *__restrict__ changed, int nodes, int new_total,
                                    const int *__restrict__ dofs, const int *__restrict__ labels,
                                    int *__restrict__ dofs2, int *__restrict__ labels2) {
  const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * RC_RUN;
  if (p0 >= new_total) return;
  int nd = node_of_entry_d(nptr, nodes, (int)p0);
  int end = __ldg(nptr + nd + 1), delta = __ldg(optr + nd) - __ldg(nptr + nd), skip = __ldg(changed + nd);
  const int p1 = (int)min((long long)new_total, p0 + RC_RUN);
  if (p1 <= end && !skip && ((p0 | (long long)delta) & 3) == 0 && p1 - p0 == RC_RUN) {
    // the whole run lies in one unchanged node and source and destination are 16-byte aligned
    const int4 *sd = reinterpret_cast<const int4 *>(dofs + p0 + delta), *sl = reinterpret_cast<const int4 *>(labels + p0 + delta);
    int4 *dd = reinterpret_cast<int4 *>(dofs2 + p0), *dl = reinterpret_cast<int4 *>(labels2 + p0);
    const int4 a0 = sd[0], a1 = sd[1], b0 = sl[0], b1 = sl[1];
    dd[0] = a0; dd[1] = a1; dl[0] = b0; dl[1] = b1;
    return;
  }
  for (int p = (int)p0; p < p1; p++) {
    while (p >= end) {  // next non-empty node
      nd++;
      end = __ldg(nptr + nd + 1);
      delta = __ldg(optr + nd) - __ldg(nptr + nd);
      skip = __ldg(changed + nd);
    }
    if (skip) continue;
    dofs2[p] = dofs[p + delta];
    labels2[p] = labels[p + delta];
  }
}

// exclusive prefix MAXIMUM over the threads of a 512-thread block (smem: int[17]); *total = block maximum
__device__ __forceinline__ int block_excl_max_scan_512(int v, int *smem, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc = max(inc, t);
  }
  int ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) ex = INT_MIN;
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 16 ? smem[lane] : INT_MIN;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi = max(wi, t);
    }
    int wex = __shfl_up_sync(0xffffffffu, wi, 1);
    if (lane == 0) wex = INT_MIN;
    if (lane < 16) smem[lane] = wex;
    if (lane == 15) smem[16] = wi;
  }
  __syncthreads();
  const int res = max(ex, smem[warp]);
  *total = smem[16];
  __syncthreads();
  return res;
}

// one CTA per node that loses or receives DOFs.  Shared memory: bitmap of the deleted labels + its word prefix
// (label compaction), histogram of the survivors over "number of joiners below" (joiner positions), scan scratch.
// A thread owns RG_ITEMS consecutive entries of a chunk, so the entry order is the thread order.
constexpr int RG_THREADS = 512;
constexpr int RG_ITEMS = 8;
__global__ void __launch_bounds__(RG_THREADS)
regroup_changed_kernel(const int *__restrict__ list, const int *__restrict__ optr, const int *__restrict__ nptr,
                       const int *__restrict__ aptr, const int *__restrict__ target, const int *__restrict__ joinS,
                       const int *__restrict__ dofs, const int *__restrict__ labels, int *__restrict__ dofs2,
                       int *__restrict__ labels2, int *__restrict__ flags /*[0] unsorted, [1] size mismatch*/) {
  extern __shared__ int rg_smem[];
  __shared__ int s_scan[33];
  const int x = list[blockIdx.x];
  const int o0 = optr[x], size = optr[x + 1] - o0;
  const int base = nptr[x], new_size = nptr[x + 1] - base;
  const int a0 = aptr[x], na = aptr[x + 1] - a0;
  const int words = (size + 31) >> 5;
  int *bits = rg_smem, *wpre = rg_smem + words, *hist = rg_smem + 2 * words;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * words + na + 1; i += RG_THREADS) rg_smem[i] = 0;
  __syncthreads();
  for (int k = tid; k < size; k += RG_THREADS)
    if (__ldg(target + dofs[o0 + k]) != kNoInit) {
      const int lab = labels[o0 + k];
      atomicOr(bits + (lab >> 5), 1 << (lab & 31));
    }
  __syncthreads();
  int running = 0;
  for (int w0 = 0; w0 < words; w0 += RG_THREADS) {
    const int w = w0 + tid;
    const int c = w < words ? __popc((unsigned)bits[w]) : 0;
    int total;
    const int ex = block_excl_scan<RG_THREADS>(c, s_scan, &total);
    if (w < words) wpre[w] = running + ex;
    running += total;
  }
  __syncthreads();
  // survivors, in entry order
  int ns = 0, last = INT_MIN;  // survivors so far, largest survivor so far (ascending <=> each exceeds it)
  bool unsorted = false;
  for (int k0 = 0; k0 < size; k0 += RG_THREADS * RG_ITEMS) {
    int v[RG_ITEMS], lab[RG_ITEMS];
    unsigned keep = 0;
    int lmax = INT_MIN;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++) {
      const int k = k0 + tid * RG_ITEMS + i;
      v[i] = 0; lab[i] = 0;
      if (k < size) {
        v[i] = dofs[o0 + k];
        lab[i] = labels[o0 + k];
        if (__ldg(target + v[i]) == kNoInit) { keep |= 1u << i; lmax = max(lmax, v[i]); }
      }
    }
    int total, tmax;
    const int ex = block_excl_scan<RG_THREADS>(__popc(keep), s_scan, &total);
    int prev = max(last, block_excl_max_scan_512(lmax, s_scan, &tmax));
    int j = ns + ex;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++) {
      if (!((keep >> i) & 1u)) continue;
      int shift = 0;
      if (na > 0) {
        shift = lower_bound_d(joinS + a0, na, v[i]);
        unsorted |= prev > v[i];  // the merge with the joiners needs ascending survivors
        prev = max(prev, v[i]);
        atomicAdd(hist + shift, 1);
      }
      dofs2[base + j + shift] = v[i];
      labels2[base + j] =
          lab[i] - (wpre[lab[i] >> 5] + __popc((unsigned)bits[lab[i] >> 5] & ((1u << (lab[i] & 31)) - 1u)));
      j++;
    }
    ns += total;
    last = max(last, tmax);
  }
  if (unsorted) flags[0] = 1;
  if (ns + na != new_size && tid == 0) flags[1] = 1;
  __syncthreads();  // the histogram is complete
  // joiner r sits behind the survivors with at most r joiners below them
  running = 0;
  for (int r0 = 0; r0 < na; r0 += RG_THREADS) {
    const int r = r0 + tid;
    const int c = r < na ? hist[r] : 0;
    int total;
    const int ex = block_excl_scan<RG_THREADS>(c, s_scan, &total);
    if (r < na) {
      const int below = running + ex + c;  // survivors with shift <= r
      dofs2[base + r + below] = joinS[a0 + r];
      labels2[base + ns + r] = ns + r;
    }
    running += total;
  }
}

// The same regrouping by a thread-block CLUSTER per node (RGC CTAs, one GPC): a leaf that loses a handful of DOFs
// still has ~30 000 entries to walk (two random 4-byte gathers each), which one CTA did in 85 us at 200^3.  Every CTA
// takes one chunk of the node's entries; the leaving-label bitmap and its word prefix are spread over the CTAs'
// shared memory by label-word range (remote atomicOr / loads through distributed shared memory), the survivor counts
// and the per-range totals are exchanged the same way, and four cluster barriers separate the phases.  The joiner
// histogram lives in rank 0's shared memory.
constexpr int RGC = 8;

__global__ void __cluster_dims__(RGC, 1, 1) __launch_bounds__(RG_THREADS)
regroup_changed_cluster_kernel(const int *__restrict__ list, const int *__restrict__ optr,
                               const int *__restrict__ nptr, const int *__restrict__ aptr,
                               const int *__restrict__ target, const int *__restrict__ joinS,
                               const int *__restrict__ dofs, const int *__restrict__ labels, int *__restrict__ dofs2,
                               int *__restrict__ labels2, int *__restrict__ flags /*[0] unsorted, [1] size mismatch*/,
                               int wper_cap, int hist_cap) {
  extern __shared__ int rg_smem[];
  __shared__ int s_scan[33];
  __shared__ int s_meta[4];   // [0] survivors of my chunk, [1] leaving labels in my word range, [2] largest survivor of my chunk
  __shared__ int s_pre[2 * RGC];  // survivors before rank r | leaving labels in the word ranges before rank r
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), tid = threadIdx.x;
  const int x = list[blockIdx.x / RGC];
  const int o0 = optr[x], size = optr[x + 1] - o0;
  const int base = nptr[x], new_size = nptr[x + 1] - base;
  const int a0 = aptr[x], na = aptr[x + 1] - a0;
  const int words = (size + 31) >> 5, wper = max(1, (words + RGC - 1) / RGC);  // label words owned per CTA (<= wper_cap)
  int *bits = rg_smem, *wpre = rg_smem + wper_cap, *hist = rg_smem + 2 * wper_cap;  // hist: rank 0 only
  for (int i = tid; i < 2 * wper_cap + (rank == 0 ? hist_cap : 0); i += RG_THREADS) rg_smem[i] = 0;
  cl.sync();
  // my chunk of the entries: whole 4096-entry tiles, so that the tile loop below is the original one
  constexpr int TILE = RG_THREADS * RG_ITEMS;
  const int tiles = (size + TILE - 1) / TILE, tper = (tiles + RGC - 1) / RGC;
  const int k_lo = min(size, rank * tper * TILE), k_hi = min(size, (rank + 1) * tper * TILE);
  // ---- phase 1: leaving labels into the distributed bitmap, survivors of my chunk counted
  int mine = 0, mymax = INT_MIN;
  for (int k = k_lo + tid; k < k_hi; k += RG_THREADS) {
    const int v = dofs[o0 + k];
    if (__ldg(target + v) != kNoInit) {
      const int lab = labels[o0 + k], w = lab >> 5, owner = w / wper;
      atomicOr(cl.map_shared_rank(bits, owner) + (w - owner * wper), 1 << (lab & 31));
    } else {
      mine++;
      mymax = max(mymax, v);
    }
  }
  {
    int total, tmax;
    block_excl_scan<RG_THREADS>(mine, s_scan, &total);
    block_excl_max_scan_512(mymax, s_scan, &tmax);
    if (tid == 0) { s_meta[0] = total; s_meta[2] = tmax; }
  }
  cl.sync();
  // ---- phase 2: word prefix inside my range
  {
    int running = 0;
    for (int w0 = 0; w0 < wper; w0 += RG_THREADS) {
      const int w = w0 + tid;
      const int c = w < wper ? __popc((unsigned)bits[w]) : 0;
      int total;
      const int ex = block_excl_scan<RG_THREADS>(c, s_scan, &total);
      if (w < wper) wpre[w] = running + ex;
      running += total;
    }
    if (tid == 0) s_meta[1] = running;
  }
  cl.sync();
  if (tid == 0) {
    int sv = 0, lv = 0;
    for (int r = 0; r < RGC; r++) {
      const int *m = cl.map_shared_rank(s_meta, r);
      s_pre[r] = sv; s_pre[RGC + r] = lv;
      sv += m[0]; lv += m[1];
    }
  }
  __syncthreads();
  // ---- phase 3: survivors of my chunk, in entry order
  int ns = s_pre[rank], last = INT_MIN;
  for (int r = 0; r < rank; r++) last = max(last, cl.map_shared_rank(s_meta, r)[2]);
  bool unsorted = false;
  int *hist0 = cl.map_shared_rank(hist, 0);
  for (int k0 = k_lo; k0 < k_hi; k0 += TILE) {
    int v[RG_ITEMS], lab[RG_ITEMS];
    unsigned keep = 0;
    int lmax = INT_MIN;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++) {
      const int k = k0 + tid * RG_ITEMS + i;
      v[i] = 0; lab[i] = 0;
      if (k < k_hi) {
        v[i] = dofs[o0 + k];
        lab[i] = labels[o0 + k];
        if (__ldg(target + v[i]) == kNoInit) { keep |= 1u << i; lmax = max(lmax, v[i]); }
      }
    }
    int total, tmax;
    const int ex = block_excl_scan<RG_THREADS>(__popc(keep), s_scan, &total);
    int prev = max(last, block_excl_max_scan_512(lmax, s_scan, &tmax));
    int j = ns + ex;
#pragma unroll
    for (int i = 0; i < RG_ITEMS; i++) {
      if (!((keep >> i) & 1u)) continue;
      int shift = 0;
      if (na > 0) {
        shift = lower_bound_d(joinS + a0, na, v[i]);
        unsorted |= prev > v[i];  // the merge with the joiners needs ascending survivors
        prev = max(prev, v[i]);
        atomicAdd(hist0 + shift, 1);
      }
      dofs2[base + j + shift] = v[i];
      const int w = lab[i] >> 5, owner = w / wper, wl = w - owner * wper;
      const int below = s_pre[RGC + owner] + cl.map_shared_rank(wpre, owner)[wl] +
                        __popc((unsigned)cl.map_shared_rank(bits, owner)[wl] & ((1u << (lab[i] & 31)) - 1u));
      labels2[base + j] = lab[i] - below;
      j++;
    }
    ns += total;
    last = max(last, tmax);
  }
  if (unsorted) flags[0] = 1;
  if (rank == RGC - 1 && tid == 0 && ns + na != new_size) flags[1] = 1;
  cl.sync();  // the histogram is complete; nobody reads remote shared memory after this point except rank 0 its own
  // ---- phase 4 (rank 0): joiner r sits behind the survivors with at most r joiners below them
  if (rank == 0) {
    const int ns_all = s_pre[RGC - 1] + cl.map_shared_rank(s_meta, RGC - 1)[0];
    int running = 0;
    for (int r0 = 0; r0 < na; r0 += RG_THREADS) {
      const int r = r0 + tid;
      const int c = r < na ? hist[r] : 0;
      int total;
      const int ex = block_excl_scan<RG_THREADS>(c, s_scan, &total);
      if (r < na) {
        const int below = running + ex + c;  // survivors with shift <= r
        dofs2[base + r + below] = joinS[a0 + r];
        labels2[base + ns_all + r] = ns_all + r;
      }
      running += total;
    }
  }
  cl.sync();  // remote shared memory (s_meta of the last rank) stays alive until rank 0 is done
}

// mesh_perm of the listed nodes (HMD::assignCoarseLocalPermToGlobal restricted to the nodes whose slice or offset changed)
__global__ void assemble_listed_kernel(const int *__restrict__ list, const int *__restrict__ node_ptr,
                                       const int *__restrict__ offs, const int *__restrict__ dofs,
                                       const int *__restrict__ labels, int *__restrict__ mesh_perm) {
  const int nd = list[blockIdx.y];
  const int a = node_ptr[nd], b = node_ptr[nd + 1], off = offs[blockIdx.y];
  for (int k = a + blockIdx.x * blockDim.x + threadIdx.x; k < b; k += gridDim.x * blockDim.x)
    mesh_perm[labels[k] + off] = dofs[k];
}

// the persistent target scratch goes back to its clean state, entry by entry
__global__ void reloc_reset_target_kernel(const int *__restrict__ moved, int n_moved, int *__restrict__ target) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_moved) target[moved[q]] = kNoInit;
}

// HMD::getCoarseHMDNodeOffset (HMD.cpp:525-536) on the host mirror
int coarse_offset(const parth_b200 &h, int node) {
  int cur = h.meta[node * 7 + 2];
  while (h.meta[cur * 7] != -1) cur = h.meta[cur * 7];
  if (h.meta[cur * 7 + 1] != -1) return coarse_offset(h, h.meta[cur * 7 + 1]);
  return h.meta[cur * 7 + 4];
}

// HMD::updateCoarseHMDNodeOffset(0) (HMD.cpp:681-697): post-order prefix of the node sizes
void update_offsets(parth_b200 &h) {
  if (h.meta[2] == -1) return;
  double acc = coarse_offset(h, 0);
  // iterative post-order through the stored links
  std::vector<std::pair<int, int>> stack{{0, 0}};
  while (!stack.empty()) {
    auto &top = stack.back();
    const int x = top.first;
    if (top.second == 0) {
      top.second = 1;
      if (h.meta[x * 7] != -1) { stack.push_back({h.meta[x * 7], 0}); continue; }
    }
    if (top.second == 1) {
      top.second = 2;
      if (h.meta[x * 7 + 1] != -1) { stack.push_back({h.meta[x * 7 + 1], 0}); continue; }
    }
    h.meta[x * 7 + 4] = (int)acc;
    acc += h.node_ptr[x + 1] - h.node_ptr[x];
    stack.pop_back();
  }
}

// Shared tail of integrate / relocate.  On entry: keep[] (w0) and del[] (w1) are filled for the
// `total` old entries; join (ascending ids) / join_node hold the n_join DOFs that enter nodes.
// Rebuilds dofs/labels (double buffered), node_ptr, offsets.  Returns the touched-node list.
void rank_joiners(parth_b200 &h, int n_touched, const int *d_touched, const int *d_join, const int *d_join_node, int n_join,
                  const int *d_aptr, int *d_out) {
  if (n_touched == 0 || n_join == 0) return;
  cudaStream_t s = h.stream;
  const int n_seg = std::max(1, std::min(128, n_join / 2048));
  const int seg = ((n_join + n_seg - 1) / n_seg + 255) / 256 * 256;
  if (n_seg == 1) {  // a short joiner list: one CTA per node walks it, no segment bases needed
    rank_joiners_kernel<<<dim3((unsigned)n_touched, 1), 256, 0, s>>>(d_touched, d_join, d_join_node, n_join, seg, nullptr,
                                                                     d_aptr, d_out);
    h.stats.kernels_launched++;
    return;
  }
  h.rank_seg.reserve((size_t)n_touched * n_seg + 1, s);
  const dim3 grid((unsigned)n_touched, (unsigned)n_seg);
  rank_joiners_count_kernel<<<grid, 256, 0, s>>>(d_touched, d_join_node, n_join, seg, h.rank_seg.p);
  rank_joiners_scan_kernel<<<nblk(n_touched), 256, 0, s>>>(h.rank_seg.p, n_touched, n_seg);
  rank_joiners_kernel<<<grid, 256, 0, s>>>(d_touched, d_join, d_join_node, n_join, seg, h.rank_seg.p, d_aptr, d_out);
  h.stats.kernels_launched += 3;
}

template <int MODE>
std::vector<int> regroup_finish(parth_b200 &h, int total, const int *key, const int *d_join, const int *d_join_node,
                                int n_join, int new_total) {
  cudaStream_t s = h.stream;
  const int nn = h.num_nodes;
  h.scan_ws.reserve(scan_ws_words((size_t)std::max(total, n_join) + 1), s);
  h.w2.reserve((size_t)total + 2, s);  // S_keep
  h.w3.reserve((size_t)total + 2, s);  // S_del
  h.stats.kernels_launched += exclusive_scan<0>(h.w0.p, h.w2.p, total, h.scan_ws.p, s);
  h.stats.kernels_launched += exclusive_scan<0>(h.w1.p, h.w3.p, total, h.scan_ws.p, s);
  // survivors per node
  h.w4.reserve((size_t)nn * 3 + 8, s);
  gather_kernel<<<nblk(nn + 1), 256, 0, s>>>(h.w2.p, h.d_node_ptr.p, nn + 1, h.w4.p);
  h.stats.kernels_launched++;
  std::vector<int> sptr((size_t)nn + 1), aptr((size_t)nn + 1, 0), nptr((size_t)nn + 1, 0), cnt((size_t)nn, 0);
  read_ints(h, h.w4.p, sptr.data(), (size_t)nn + 1);
  const int total_surv = sptr[nn];
  // joiners per node
  if (n_join > 0) {
    PB_CUDA(cudaMemsetAsync(h.w4.p, 0, sizeof(int) * (size_t)nn, s));
    count_joiners_kernel<<<nblk(n_join), 256, 0, s>>>(d_join_node, n_join, h.w4.p);
    h.stats.kernels_launched++;
    read_ints(h, h.w4.p, cnt.data(), (size_t)nn);
  }
  std::vector<int> touched;
  for (int x = 0; x < nn; x++) {
    aptr[x + 1] = aptr[x] + cnt[x];
    nptr[x + 1] = nptr[x] + (sptr[x + 1] - sptr[x]) + cnt[x];
    if (cnt[x]) touched.push_back(x);
  }
  if (nptr[nn] != new_total) throw StatusError(PARTH_B200_ERR_INTERNAL, "regroup: entry count mismatch");
  // device copies of the three pointer arrays + touched list
  h.w5.reserve((size_t)3 * (nn + 1) + touched.size() + 8, s);
  int *d_sptr = h.w5.p, *d_aptr = h.w5.p + (nn + 1), *d_nptr = h.w5.p + 2 * (nn + 1), *d_touched = h.w5.p + 3 * (nn + 1);
  PB_CUDA(cudaMemcpyAsync(d_sptr, sptr.data(), sizeof(int) * ((size_t)nn + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(d_aptr, aptr.data(), sizeof(int) * ((size_t)nn + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(d_nptr, nptr.data(), sizeof(int) * ((size_t)nn + 1), cudaMemcpyHostToDevice, s));
  if (!touched.empty())
    PB_CUDA(cudaMemcpyAsync(d_touched, touched.data(), sizeof(int) * touched.size(), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));
  // compacted survivors
  h.r_cnt.reserve((size_t)std::max(total_surv, 1), s);  // survC
  h.r_G.reserve((size_t)std::max(total_surv, 1), s);    // labC
  h.r_sub.reserve((size_t)std::max(n_join, 1), s);      // joinS
  if (total > 0) {
    regroup_compact_kernel<MODE><<<nblk(total), 256, 0, s>>>(h.d_node_ptr.p, nn, h.d_dofs.p, h.d_labels.p, total, key,
                                                            h.w2.p, h.w3.p, h.r_cnt.p, h.r_G.p);
    h.stats.kernels_launched++;
  }
  rank_joiners(h, (int)touched.size(), d_touched, d_join, d_join_node, n_join, d_aptr, h.r_sub.p);
  h.d_dofs2.reserve((size_t)std::max(new_total, 1), s);
  h.d_labels2.reserve((size_t)std::max(new_total, 1), s);
  int *d_flag = h.mail.p + 16;
  PB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
  if (total_surv > 0) {
    regroup_place_surv_kernel<<<nblk(total_surv), 256, 0, s>>>(d_sptr, d_aptr, d_nptr, nn, total_surv, h.r_cnt.p, h.r_G.p,
                                                              h.r_sub.p, h.d_dofs2.p, h.d_labels2.p, d_flag);
    h.stats.kernels_launched++;
  }
  if (n_join > 0) {
    regroup_place_join_kernel<<<nblk(n_join), 256, 0, s>>>(d_sptr, d_aptr, d_nptr, nn, n_join, h.r_cnt.p, h.r_sub.p,
                                                          h.d_dofs2.p, h.d_labels2.p);
    h.stats.kernels_launched++;
  }
  int unsorted = 0;
  read_ints(h, d_flag, &unsorted, 1);
  if (unsorted)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                      "DOF map: surviving DOFs of a node that receives DOFs are not in ascending order "
                      "(the reference expects survivors to keep their order, remeshPatch.cpp:160-176)");
  h.d_dofs.swap(h.d_dofs2);
  h.d_labels.swap(h.d_labels2);
  h.node_ptr = nptr;
  update_offsets(h);
  upload_tree_meta(h);
  return touched;
}

}  // namespace

// ------------------------------------------------------------------ (a3)
void stage_set_map(parth_b200 &h, const int *map, int len, bool on_device) {
  cudaStream_t s = h.stream;
  h.map_len = 0;
  h.n_added = h.n_removed = 0;
  if (len == 0 || h.M_n_prev == 0) return;  // ParthAPI.cpp:138-142: no previous mesh -> the map is dropped
  if (len != h.M_n) throw StatusError(PARTH_B200_ERR_ARG, "setNewToOldDOFMap: map length must equal M_n");
  h.morton_n = 0;  // the DOF set changes: coordinates (MORTON_CODE) must be supplied again, after this call
  const int M = h.M_n, Mp = h.M_n_prev;
  h.new_to_old.reserve((size_t)M, s);
  h.old_to_new.reserve((size_t)Mp, s);
  PB_CUDA(cudaMemcpyAsync(h.new_to_old.p, map, sizeof(int) * (size_t)M, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (!on_device) h.stats.h2d_bytes += (long long)sizeof(int) * M;
  int *d_bad = h.mail.p + 16;
  PB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
  fill_kernel<<<nblk(Mp), 256, 0, s>>>(h.old_to_new.p, Mp, -1);
  invert_map_kernel<<<nblk(M), 256, 0, s>>>(h.new_to_old.p, M, Mp, h.old_to_new.p, d_bad);
  h.stats.kernels_launched += 2;
  h.scan_ws.reserve(scan_ws_words((size_t)std::max(M, Mp) + 1), s);
  h.w0.reserve((size_t)std::max(M, Mp) + 2, s);
  h.stats.kernels_launched += exclusive_scan<3>(h.new_to_old.p, h.w0.p, M, h.scan_ws.p, s);
  int n_added = 0, bad = 0;
  read_ints(h, h.w0.p + M, &n_added, 1);
  read_ints(h, d_bad, &bad, 1);
  if (bad) throw StatusError(PARTH_B200_ERR_INPUT, "setNewToOldDOFMap: old DOF id out of range");
  h.added.reserve((size_t)std::max(n_added, 1), s);
  list_minus_one_kernel<<<nblk(M), 256, 0, s>>>(h.new_to_old.p, h.w0.p, M, h.added.p);
  h.stats.kernels_launched += 1 + exclusive_scan<3>(h.old_to_new.p, h.w0.p, Mp, h.scan_ws.p, s);
  int n_removed = 0;
  read_ints(h, h.w0.p + Mp, &n_removed, 1);
  h.removed.reserve((size_t)std::max(n_removed, 1), s);
  list_minus_one_kernel<<<nblk(Mp), 256, 0, s>>>(h.old_to_new.p, h.w0.p, Mp, h.removed.p);
  h.stats.kernels_launched++;
  if ((double)n_added > 0.4 * M) {  // ParthAPI.cpp:181-188: too many new DOFs -> forget everything, cold start
    h.tree_init = false;
    return;
  }
  h.map_len = len;
  h.n_added = n_added;
  h.n_removed = n_removed;
}

// ------------------------------------------------------------------ (a6)
void stage_integrate(parth_b200 &h) {
  cudaStream_t s = h.stream;
  StageTimer tm(h, &h.stats.integrate_ms, false, &h.timer_ms[1]);
  const int nn = h.num_nodes, M = h.M_n;
  const int total = h.node_ptr[nn];
  h.w0.reserve((size_t)std::max(total, M) + 2, s);  // keep
  h.w1.reserve((size_t)std::max(total, M) + 2, s);  // del (label space)
  h.r_chosen.reserve((size_t)M + 1, s);             // nd2n -> becomes dof2node
  fill_kernel<<<nblk(M), 256, 0, s>>>(h.r_chosen.p, M, -1);
  h.stats.kernels_launched++;
  if (total > 0) {
    regroup_flags_kernel<0><<<nblk(total), 256, 0, s>>>(h.d_node_ptr.p, nn, h.d_dofs.p, h.d_labels.p, total,
                                                       h.old_to_new.p, h.w0.p, h.w1.p, h.r_chosen.p);
    h.stats.kernels_launched++;
  }
  int *nd2n = h.r_chosen.p;
  const int n_added = h.n_added;
  h.r_perm.reserve((size_t)std::max(n_added, 1), s);  // node of each added DOF
  if (n_added > 0) {
    // ---- ring placement (Integrator.cpp:118-202)
    h.r_g2l.reserve((size_t)M + 1, s);  // placed_pass
    h.r_l2g.reserve((size_t)M + 1, s);  // indeg
    // one launch per pass (place_pass_pull_kernel: r_g2l holds the state words); PARTH_B200_PLACE_PER_SWEEP forces
    // the launch-per-sweep path (cross-check in tests/test_gpu_scenarios.py; r_g2l = placed_pass, r_l2g = indeg)
    // (the pull kernel's warps poll each other: it needs a cooperative launch of one CTA per SM; a context that cannot
    // give that -- no cooperative launch, SMs taken away by a partition -- takes the launch-per-sweep path)
    static int pull_ok[64] = {};  // per device: 0 unknown, 1 yes, -1 no
    {
      const int dev = h.device;
      if (dev >= 0 && dev < 64 && pull_ok[dev] == 0) {
        int coop = 0, per_sm = 0;
        PB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_pass_pull_kernel, kPullThreads, 0));
        pull_ok[dev] = (coop && per_sm >= 1) ? 1 : -1;
      }
    }
    const bool pull_path = !std::getenv("PARTH_B200_PLACE_PER_SWEEP") && nn < 65535 &&
                           (h.device < 0 || h.device >= 64 || pull_ok[h.device] == 1);
    if (!pull_path) PB_CUDA(cudaMemsetAsync(h.r_g2l.p, 0, sizeof(int) * (size_t)M, s));
    constexpr int kBatch = 16;          // launch-per-sweep path: sweeps enqueued per read-back
    const int max_sweeps = n_added + kBatch + 2;
    pb::DevBuf<int> cbuf, fr, fcnt;
    cbuf.reserve(8, s);
    fr.reserve((size_t)2 * n_added + 2, s);
    fcnt.reserve(pull_path ? 1024 : (size_t)max_sweeps + kBatch + 4, s);
    PlaceCounters *d_cnt = reinterpret_cast<PlaceCounters *>(cbuf.p);
    int *F[2] = {fr.p, fr.p + n_added};
    std::vector<int> fc((size_t)kBatch);
    const unsigned grid = std::min<unsigned>(nblk(n_added), 4u * (unsigned)h.num_sms);
    h.stats.integrate_passes = h.stats.integrate_sweeps = 0;
    for (int pass = 1;; pass++) {
      h.stats.integrate_passes = pass;
      PB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(PlaceCounters), s));
      int c[4] = {0, 0, 0, 0};
      if (pull_path) {
        if (pass > 30000) throw StatusError(PARTH_B200_ERR_INTERNAL, "integrate: more than 30000 placement passes");
        if (pass == 1) {
          place_state_init_kernel<<<nblk(M), 256, 0, s>>>(nd2n, M, h.r_g2l.p);
          h.stats.kernels_launched++;
        }
        // fcnt: [0, Q*32) queue tails (one per 128-byte line), then Q counts, then Q + 1 segment offsets
        int *tails = fcnt.p, *qcount = fcnt.p + kPullQueues * kTailStride, *qoff = qcount + kPullQueues;
        PB_CUDA(cudaMemsetAsync(fcnt.p, 0, sizeof(int) * (kPullQueues * kTailStride + 2 * kPullQueues + 1), s));
        PB_CUDA(cudaMemsetAsync(fr.p, 0xff, sizeof(int) * (size_t)n_added, s));        // queue slots: -1 = empty
        place_pull_count_kernel<<<nblk(n_added), 256, 0, s>>>(h.added.p, n_added, h.Mp, h.Mi, h.r_g2l.p, h.r_l2g.p, qcount);
        place_pull_offsets_kernel<<<1, 32, 0, s>>>(qcount, qoff, d_cnt);
        place_pull_roots_kernel<<<nblk(n_added), 256, 0, s>>>(h.added.p, n_added, h.r_g2l.p, h.r_l2g.p, fr.p, tails, qoff);
        {
          // the warps poll each other: a COOPERATIVE launch, so that the runtime refuses the grid instead of letting
          // resident CTAs spin on tickets of CTAs that cannot be scheduled (a GPU shared with another context)
          int *a_q = fr.p, *a_t = tails, *a_st = h.r_g2l.p, *a_in = h.r_l2g.p, a_pass = pass;
          const int *a_off = qoff, *a_mp = h.Mp, *a_mi = h.Mi, *a_meta = h.meta_heap ? nullptr : h.d_meta.p;
          PlaceCounters *a_cnt = d_cnt;
          void *args[] = {&a_q, &a_t, &a_off, &a_mp, &a_mi, &a_meta, &a_st, &a_in, &a_pass, &a_cnt};
          PB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(place_pass_pull_kernel), dim3((unsigned)h.num_sms),
                                              dim3(kPullThreads), args, 0, s));
        }
        h.stats.kernels_launched += 4;
        read_ints(h, reinterpret_cast<const int *>(d_cnt), c, 4);
      } else {
        PB_CUDA(cudaMemsetAsync(fcnt.p, 0, sizeof(int) * ((size_t)max_sweeps + kBatch + 4), s));
        place_pass_init_kernel<<<nblk(n_added), 256, 0, s>>>(h.added.p, n_added, h.Mp, h.Mi, h.new_to_old.p, h.r_g2l.p,
                                                            h.r_l2g.p, F[0], fcnt.p, d_cnt);
        h.stats.kernels_launched++;
        int sweep = 0;
        bool done = false;
        while (!done) {
          for (int rep = 0; rep < kBatch; rep++, sweep++) {
            place_sweep_kernel<<<grid, 256, 0, s>>>(F[sweep & 1], fcnt.p + sweep, F[(sweep + 1) & 1], fcnt.p + sweep + 1,
                                                    h.Mp, h.Mi, h.new_to_old.p, h.meta_heap ? nullptr : h.d_meta.p, nd2n,
                                                    h.r_g2l.p, h.r_l2g.p, pass, d_cnt);
            h.stats.kernels_launched++;
          }
          // an empty frontier ends the pass (the dependency graph is acyclic)
          read_ints(h, fcnt.p + sweep - kBatch + 1, fc.data(), (size_t)kBatch);
          for (int v : fc) if (v == 0) done = true;
          if (sweep > max_sweeps) throw StatusError(PARTH_B200_ERR_INTERNAL, "integrate: placement did not terminate");
        }
        h.stats.integrate_sweeps += sweep;
        read_ints(h, reinterpret_cast<const int *>(d_cnt), c, 4);
      }
      if (c[0] == 0 || c[1] == 0) break;  // nothing left, or a pass that placed nothing ends the loop
      if (c[1] == c[0]) break;            // everything that was left has been placed
    }
    if (pull_path) {
      place_decode_kernel<<<nblk(n_added), 256, 0, s>>>(h.added.p, n_added, h.r_g2l.p, nd2n);
      h.stats.kernels_launched++;
    }
    // ---- leftovers: last empty leaf, else the last node (Integrator.cpp:204-217)
    h.w2.reserve((size_t)total + 2, s);
    h.scan_ws.reserve(scan_ws_words((size_t)total + 1), s);
    h.stats.kernels_launched += exclusive_scan<0>(h.w0.p, h.w2.p, total, h.scan_ws.p, s);
    h.w4.reserve((size_t)nn * 3 + 8, s);
    gather_kernel<<<nblk(nn + 1), 256, 0, s>>>(h.w2.p, h.d_node_ptr.p, nn + 1, h.w4.p);
    h.stats.kernels_launched++;
    std::vector<int> sp((size_t)nn + 1);
    read_ints(h, h.w4.p, sp.data(), (size_t)nn + 1);
    int spare = -1;
    for (int x = nn - 1; x >= 0; x--)
      if (h.meta[x * 7] == -1 && h.meta[x * 7 + 1] == -1 && sp[x + 1] == sp[x]) { spare = x; break; }
    if (spare == -1) spare = nn - 1;
    finish_added_kernel<<<nblk(n_added), 256, 0, s>>>(h.added.p, n_added, spare, nd2n, h.r_perm.p);
    h.stats.kernels_launched++;
  }
  regroup_finish<0>(h, total, h.old_to_new.p, h.added.p, h.r_perm.p, n_added, M);
  // DOF_to_HMD_node = new_DOF_to_HMD_node; mesh_perm.clear(); resize(M_n); assign from the root
  h.d_dof2node.swap(h.r_chosen);
  h.tree_M = M;
  h.d_mesh_perm.reserve((size_t)std::max(M, 1), s);
  PB_CUDA(cudaMemsetAsync(h.d_mesh_perm.p, 0, sizeof(int) * (size_t)M, s));
  assemble_all(h);
  tm.stop();
}


// the persistent per-DOF scratch of the sparse relocation and its output slices, sized for M DOFs.  Called when a tree
// is imported, so that the first re-ordering update of a handle does not pay for the memory-pool growth (20 ms at
// 134 M DOF, profiles/r04_configs.md), and lazily by the relocation itself.
void relocate_prepare(parth_b200 &h, int M) {
  cudaStream_t s = h.stream;
  if (!h.reloc_sticky.p) {
    h.reloc_sticky.reserve(4, s);
    PB_CUDA(cudaMemsetAsync(h.reloc_sticky.p, 0, sizeof(int) * 4, s));
  }
  if (h.rel_M == M || M <= 0) return;
  h.rel_occ.reserve((size_t)M + 1, s);
  h.rel_target.reserve((size_t)M + 1, s);
  h.d_dofs2.reserve((size_t)M, s);
  h.d_labels2.reserve((size_t)M, s);
  PB_CUDA(cudaMemsetAsync(h.rel_occ.p, 0, sizeof(int) * (size_t)M, s));
  fill_kernel<<<nblk(M), 256, 0, s>>>(h.rel_target.p, M, kNoInit);
  h.stats.kernels_launched++;
  h.rel_M = M;
}

// the regroup kernels' sticky error flags: checked now, or left for the next synchronising call (stream-ordered API)
void check_reloc_flags(parth_b200 &h) {
  if (!h.reloc_check_pending || !h.reloc_sticky.p) return;
  h.reloc_check_pending = false;
  int fl[2] = {0, 0};
  read_ints(h, h.reloc_sticky.p, fl, 2);
  if (fl[0] || fl[1]) PB_CUDA(cudaMemsetAsync(h.reloc_sticky.p, 0, sizeof(int) * 2, h.stream));
  if (fl[1]) throw StatusError(PARTH_B200_ERR_INTERNAL, "relocate: regrouped node size mismatch");
  if (fl[0])
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                      "DOF map: surviving DOFs of a node that receives DOFs are not in ascending order "
                      "(the reference expects survivors to keep their order, remeshPatch.cpp:160-176)");
}

// Sparse relocation (kernels: "relocation, sparse formulation" above).  Same result as the dense path below.
static void relocate_sparse(parth_b200 &h, int threshold_node, std::vector<int> &dirty_fine_out, bool filter_edges) {
  cudaStream_t s = h.stream;
  const int nn = h.num_nodes, M = h.tree_M, ne = h.n_changed;
  relocate_prepare(h, M);
  h.w2.reserve((size_t)std::max(ne, 1) * 2, s);  // kept edges
  h.w4.reserve((size_t)nn * 2 + 8, s);           // cnt_out | cnt_in
  h.r_perm.reserve((size_t)kSparseCand, s);      // moved ids, ascending
  h.r_ws.reserve((size_t)kSparseCand, s);        // their target nodes
  // behind the per-node counts, so that ONE copy fetches everything:
  // [0] n_moved, [1] unsorted survivors, [2] size mismatch, [3] kept edges, [4..5] sticky flags of earlier updates
  int *d_cnt = h.w4.p + 2 * (size_t)nn;
  const size_t front_smem = sizeof(int) * kSparseCand + kSparseEdges;
  {
    static bool attr[64] = {};  // > 48 KB of dynamic shared memory: opt in once per device
    const int dev = h.device;
    if (dev < 0 || dev >= 64 || !attr[dev]) {
      PB_CUDA(cudaFuncSetAttribute(reloc_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)front_smem));
      if (dev >= 0 && dev < 64) attr[dev] = true;
    }
  }
  PB_MARK("reloc: entry");
  reloc_front_kernel<<<1, 1024, front_smem, s>>>(h.d_changed.p, ne, h.d_dof2node.p, h.d_meta.p, threshold_node, h.rel_occ.p,
                                                 h.rel_target.p, h.w2.p, h.r_perm.p, h.r_ws.p, h.w4.p, nn, d_cnt,
                                                 filter_edges ? 1 : 0, h.reloc_sticky.p);
  PB_CUDA(cudaGetLastError());
  h.stats.kernels_launched++;
  // ONE read-back: kept edges, moved DOFs, leavers / joiners per node
  h.pin.reserve((size_t)2 * nn + 8);
  // (the sticky error flags of an earlier stream-ordered update ride along: this update's regroup kernels run later)
  PB_CUDA(cudaMemcpyAsync(h.pin.p, h.w4.p, sizeof(int) * ((size_t)nn * 2 + 6), cudaMemcpyDeviceToHost, s));
  PB_MARK("reloc: front + read-back enqueued");
  PB_CUDA(cudaStreamSynchronize(s));
  PB_MARK("reloc: read-back arrived");
  h.stats.d2h_bytes += (long long)sizeof(int) * (2 * nn + 6);
  if (h.reloc_check_pending) {
    h.reloc_check_pending = false;
    const int f0 = h.pin.p[2 * (size_t)nn + 4], f1 = h.pin.p[2 * (size_t)nn + 5];
    if (f0 || f1) PB_CUDA(cudaMemsetAsync(h.reloc_sticky.p, 0, sizeof(int) * 2, s));
    if (f1) throw StatusError(PARTH_B200_ERR_INTERNAL, "relocate: regrouped node size mismatch (previous update)");
    if (f0)
      throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                        "DOF map: surviving DOFs of a node that receives DOFs are not in ascending order (previous update)");
  }
  const int kept = h.pin.p[2 * (size_t)nn + 3], n_moved = h.pin.p[2 * (size_t)nn];
  const std::vector<int> cnt_out(h.pin.p, h.pin.p + nn), cnt_in(h.pin.p + nn, h.pin.p + 2 * nn);
  if (filter_edges) h.n_changed = kept;
  if (n_moved == 0) { collect_timers(h); return; }
  // new layout on the host (<= 8191 nodes); the device idles until the kernels below are enqueued
  std::vector<int> aptr((size_t)nn + 1, 0), nptr((size_t)nn + 1, 0), changed_flag((size_t)nn, 0), changed, touched;
  int max_words = 0, max_in = 0;
  for (int x = 0; x < nn; x++) {
    const int size = h.node_ptr[x + 1] - h.node_ptr[x];
    aptr[x + 1] = aptr[x] + cnt_in[x];
    nptr[x + 1] = nptr[x] + size - cnt_out[x] + cnt_in[x];
    if (cnt_in[x]) touched.push_back(x);
    if (cnt_in[x] || cnt_out[x]) {
      changed.push_back(x);
      changed_flag[x] = 1;
      max_words = std::max(max_words, (size + 31) / 32);
      max_in = std::max(max_in, cnt_in[x]);
    }
  }
  const int new_total = nptr[nn];
  if (new_total != h.node_ptr[nn]) throw StatusError(PARTH_B200_ERR_INTERNAL, "relocate: entry count mismatch");
  // offsets of the new layout (host, <= 8191 nodes); mesh_perm is re-assembled only for the nodes whose slice or
  // offset changed (DOFs move inside the sub-tree of the receiving separator: nothing outside it shifts)
  std::vector<int> old_off((size_t)nn);
  for (int x = 0; x < nn; x++) old_off[x] = h.meta[(size_t)x * 7 + 4];
  h.node_ptr = nptr;
  update_offsets(h);
  std::vector<int> asm_list, asm_off;
  for (int x = 0; x < nn; x++)
    if (nptr[x + 1] > nptr[x] && (changed_flag[x] || h.meta[(size_t)x * 7 + 4] != old_off[x])) {
      asm_list.push_back(x);
      asm_off.push_back(h.meta[(size_t)x * 7 + 4]);
    }
  // one upload: aptr | nptr | changed flags | changed list | touched list | assemble list | its offsets | the offsets of
  // all nodes (pinned: no wait)
  const size_t up = (size_t)3 * nn + 2 + changed.size() + touched.size() + 2 * asm_list.size() + nn;
  h.pin_rel.reserve(up + 8);
  h.w5.reserve(up + 8, s);
  {
    int *q = h.pin_rel.p;
    std::copy(aptr.begin(), aptr.end(), q); q += nn + 1;
    std::copy(nptr.begin(), nptr.end(), q); q += nn + 1;
    std::copy(changed_flag.begin(), changed_flag.end(), q); q += nn;
    std::copy(changed.begin(), changed.end(), q); q += changed.size();
    std::copy(touched.begin(), touched.end(), q); q += touched.size();
    std::copy(asm_list.begin(), asm_list.end(), q); q += asm_list.size();
    std::copy(asm_off.begin(), asm_off.end(), q); q += asm_off.size();
    for (int x = 0; x < nn; x++) q[x] = h.meta[(size_t)x * 7 + 4];
  }
  PB_MARK("reloc: layout planned");
  PB_CUDA(cudaMemcpyAsync(h.w5.p, h.pin_rel.p, sizeof(int) * up, cudaMemcpyHostToDevice, s));
  const int *d_aptr = h.w5.p, *d_nptr = h.w5.p + (nn + 1), *d_flag = h.w5.p + 2 * (nn + 1);
  const int *d_changed_list = d_flag + nn, *d_touched = d_changed_list + changed.size();
  const int *d_asm = d_touched + touched.size(), *d_asm_off = d_asm + asm_list.size();
  const int *d_off_all = d_asm_off + asm_list.size();
  h.r_sub.reserve((size_t)std::max(n_moved, 1), s);  // joiners grouped by node
  h.d_dofs2.reserve((size_t)std::max(new_total, 1), s);
  h.d_labels2.reserve((size_t)std::max(new_total, 1), s);
  // fork: the streaming copy of the untouched slices (the longest kernel of the stage, it only needs the upload) runs
  // on the side stream while this one ranks the joiners and regroups the changed nodes (a few CTAs, latency-bound);
  // the two write disjoint slices of the new arrays.  PARTH_B200_RELOC_ONE_STREAM keeps everything on one stream.
  static const bool one_stream = std::getenv("PARTH_B200_RELOC_ONE_STREAM") != nullptr;
  cudaStream_t sc = one_stream ? s : h.stream2;
  if (!one_stream) {
    PB_CUDA(cudaEventRecord(h.ev_fork, s));
    PB_CUDA(cudaStreamWaitEvent(sc, h.ev_fork, 0));
  }
  regroup_copy_kernel<<<nblk((new_total + RC_RUN - 1) / RC_RUN), 256, 0, sc>>>(h.d_node_ptr.p, d_nptr, d_flag, nn, new_total,
                                                                             h.d_dofs.p, h.d_labels.p, h.d_dofs2.p, h.d_labels2.p);
  if (!one_stream) PB_CUDA(cudaEventRecord(h.ev_join, sc));
  rank_joiners(h, (int)touched.size(), d_touched, h.r_perm.p, h.r_ws.p, n_moved, d_aptr, h.r_sub.p);
  // one cluster of RGC CTAs per changed node; shared memory: its slice of the label bitmap + word prefix, and (rank 0)
  // the joiner histogram.  PARTH_B200_REGROUP_ONE_CTA keeps the one-CTA-per-node kernel (cross-check in tests)
  const int wper_cap = ((max_words + RGC - 1) / RGC + 3) & ~3, hist_cap = max_in + 2;
  const size_t smem_c = sizeof(int) * ((size_t)2 * wper_cap + hist_cap);
  if (smem_c <= 200 * 1024 && !std::getenv("PARTH_B200_REGROUP_ONE_CTA")) {
    if (smem_c > 48 * 1024)
      PB_CUDA(cudaFuncSetAttribute(regroup_changed_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    regroup_changed_cluster_kernel<<<(unsigned)changed.size() * RGC, RG_THREADS, smem_c, s>>>(
        d_changed_list, h.d_node_ptr.p, d_nptr, d_aptr, h.rel_target.p, h.r_sub.p, h.d_dofs.p, h.d_labels.p, h.d_dofs2.p,
        h.d_labels2.p, h.reloc_sticky.p, wper_cap, hist_cap);
  } else {
    const size_t smem = sizeof(int) * ((size_t)2 * max_words + max_in + 2);
    if (smem > 200 * 1024) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "relocate: a node is too large for the regroup kernel");
    if (smem > 48 * 1024)
      PB_CUDA(cudaFuncSetAttribute(regroup_changed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    regroup_changed_kernel<<<(unsigned)changed.size(), RG_THREADS, smem, s>>>(d_changed_list, h.d_node_ptr.p, d_nptr, d_aptr,
                                                                            h.rel_target.p, h.r_sub.p, h.d_dofs.p, h.d_labels.p,
                                                                            h.d_dofs2.p, h.d_labels2.p, h.reloc_sticky.p);
  }
  reloc_reset_target_kernel<<<nblk(n_moved), 256, 0, s>>>(h.r_perm.p, n_moved, h.rel_target.p);
  PB_CUDA(cudaGetLastError());
  if (!one_stream) PB_CUDA(cudaStreamWaitEvent(s, h.ev_join, 0));  // join: everything below reads the copied slices
  h.stats.kernels_launched += 3;
  h.d_dofs.swap(h.d_dofs2);
  h.d_labels.swap(h.d_labels2);
  if (!asm_list.empty()) {
    // (d_nptr: the new node_ptr, already on the device; upload_tree_meta below refreshes the handle's own copy)
    assemble_listed_kernel<<<dim3(32, (unsigned)asm_list.size()), 256, 0, s>>>(d_asm, d_nptr, d_asm_off, h.d_dofs.p,
                                                                              h.d_labels.p, h.d_mesh_perm.p);
    PB_CUDA(cudaGetLastError());
    h.stats.kernels_launched++;
  }
  PB_MARK("reloc: kernels enqueued");
  // the device mirrors of node_ptr and of the offsets (only these two change: ids, links, levels and the visit order
  // stay) are refreshed from the upload above -- upload_tree_meta's three pageable copies cost ~20 us of host time
  // right where the host has to get ahead of the device again
  apply_layout_kernel<<<nblk(nn + 1), 256, 0, s>>>(d_nptr, d_off_all, nn, h.d_node_ptr.p, h.d_meta.p);
  PB_CUDA(cudaGetLastError());
  h.stats.kernels_launched++;
  PB_MARK("reloc: tree meta uploaded");
  collect_timers(h);  // (the stage timers that completed with the read-back; off the critical path here)
  h.reloc_check_pending = true;
  if (!h.defer_reloc_check) check_reloc_flags(h);  // host-pointer API: errors are reported by the call that caused them
  dirty_fine_out.insert(dirty_fine_out.end(), touched.begin(), touched.end());
}

// ------------------------------------------------------------------ (f2)
// Moves one endpoint of every changed edge whose common separator lies above `threshold_node`
// into that separator; the edge list keeps only the edges that still need a repair.
void stage_relocate(parth_b200 &h, int threshold_node, std::vector<int> &dirty_fine_out, bool filter_edges) {
  cudaStream_t s = h.stream;
  const int nn = h.num_nodes, M = h.tree_M, ne = h.n_changed;
  if (ne <= 0) return;
  const int total = h.node_ptr[nn];
  if (ne <= kSparseEdges && total == M && !std::getenv("PARTH_B200_DENSE_RELOCATE")) {
    relocate_sparse(h, threshold_node, dirty_fine_out, filter_edges);
    return;
  }
  h.r_chosen.reserve((size_t)M + 1, s);  // occ
  h.r_g2l.reserve((size_t)M + 2, s);     // target
  h.r_l2g.reserve((size_t)std::max(M, ne) + 2, s);  // moved flag / its prefix lives in w1
  h.w0.reserve((size_t)std::max(std::max(total, M), ne) + 2, s);
  h.w1.reserve((size_t)std::max(std::max(total, M), ne) + 2, s);
  h.scan_ws.reserve(scan_ws_words((size_t)std::max(std::max(total, M), ne) + 1), s);
  PB_CUDA(cudaMemsetAsync(h.r_chosen.p, 0, sizeof(int) * (size_t)M, s));
  fill_kernel<<<nblk(M), 256, 0, s>>>(h.r_g2l.p, M, kNoInit);
  occ_kernel<<<nblk(ne), 256, 0, s>>>(h.d_changed.p, ne, h.r_chosen.p);
  relocate_targets_kernel<<<nblk(ne), 256, 0, s>>>(h.d_changed.p, ne, h.d_dof2node.p, h.d_meta.p, h.r_chosen.p,
                                                  threshold_node, h.r_g2l.p, h.w0.p);
  h.stats.kernels_launched += 3;
  // filtered edge list, order preserved
  h.stats.kernels_launched += exclusive_scan<0>(h.w0.p, h.w1.p, ne, h.scan_ws.p, s);
  int kept = 0;
  read_ints(h, h.w1.p + ne, &kept, 1);
  h.w2.reserve((size_t)std::max(kept, 1) * 2, s);
  compact_edges_kernel<<<nblk(ne), 256, 0, s>>>(h.d_changed.p, ne, h.w0.p, h.w1.p, h.w2.p);
  h.stats.kernels_launched++;
  if (filter_edges) {
    if (kept > 0)
      PB_CUDA(cudaMemcpyAsync(h.d_changed.p, h.w2.p, sizeof(int) * 2 * (size_t)kept, cudaMemcpyDeviceToDevice, s));
    h.n_changed = kept;
  }
  // moved DOFs, ascending
  moved_flags_kernel<<<nblk(M), 256, 0, s>>>(h.r_g2l.p, M, h.r_l2g.p);
  h.stats.kernels_launched += 1 + exclusive_scan<0>(h.r_l2g.p, h.w1.p, M, h.scan_ws.p, s);
  int n_moved = 0;
  read_ints(h, h.w1.p + M, &n_moved, 1);
  if (n_moved == 0) return;
  h.r_perm.reserve((size_t)n_moved, s);  // moved ids
  h.r_ws.reserve((size_t)n_moved, s);    // their target nodes
  if (total > 0) {
    // flags are computed BEFORE dof2node is redirected (the nodes that lose DOFs are the old ones)
    regroup_flags_kernel<1><<<nblk(total), 256, 0, s>>>(h.d_node_ptr.p, nn, h.d_dofs.p, h.d_labels.p, total,
                                                       h.r_g2l.p, h.w0.p, h.r_chosen.p /*del, reuses occ*/, nullptr);
    h.stats.kernels_launched++;
  }
  list_moved_kernel<<<nblk(M), 256, 0, s>>>(h.r_l2g.p, h.w1.p, h.r_g2l.p, M, h.r_perm.p, h.r_ws.p, h.d_dof2node.p);
  h.stats.kernels_launched++;
  // regroup_finish expects del[] in w1
  PB_CUDA(cudaMemcpyAsync(h.w1.p, h.r_chosen.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToDevice, s));
  std::vector<int> touched = regroup_finish<1>(h, total, h.r_g2l.p, h.r_perm.p, h.r_ws.p, n_moved, total);
  dirty_fine_out.insert(dirty_fine_out.end(), touched.begin(), touched.end());
  assemble_all(h);
}

}  // namespace pb
