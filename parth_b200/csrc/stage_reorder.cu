// parth_b200/csrc/stage_reorder.cu -- host drivers of stage (c): extraction of dirty fine nodes
// (extract.cu), batched ordering (amd.cu), label / permutation update.
// Reference: Integrator::permuteFineHMDNodes (src/Parth/Integrator.cpp:851-893),
// HMD::createMultiSubMesh (HMD.cpp:538-612), HMD::getCoarseMesh (HMD.cpp:465-481),
// HMD::AMDNodeNDPermutation (HMD.cpp:58-75).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "reorder.cuh"
#include "scan.cuh"
#include "stages.cuh"

namespace pb {

namespace {

struct Batch {
  int count = 0, total = 0, nnz = 0;
  std::vector<int> vtx;  // count + 1
};

// uploads the node list / vertex offsets of a batch of fine nodes (r_list, r_vtx)
Batch batch_of(parth_b200 &h, const std::vector<int> &nodes) {
  cudaStream_t s = h.stream;
  Batch b;
  b.count = (int)nodes.size();
  b.vtx.assign(b.count + 1, 0);
  for (int g = 0; g < b.count; g++) {
    const int nd = nodes[g];
    if (nd < 0 || nd >= h.num_nodes) throw StatusError(PARTH_B200_ERR_ARG, "permute_fine: node id out of range");
    b.vtx[g + 1] = b.vtx[g] + (h.node_ptr[nd + 1] - h.node_ptr[nd]);
  }
  b.total = b.vtx[b.count];
  h.r_list.reserve((size_t)b.count + 1, s);
  h.r_vtx.reserve((size_t)b.count + 2, s);
  PB_CUDA(cudaMemcpyAsync(h.r_list.p, nodes.data(), sizeof(int) * (size_t)b.count, cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.r_vtx.p, b.vtx.data(), sizeof(int) * ((size_t)b.count + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));  // pageable sources
  return b;
}

// ... and builds the concatenated sub-mesh CSR (r_G, r_sub)
Batch extract_fine(parth_b200 &h, const std::vector<int> &nodes) {
  cudaStream_t s = h.stream;
  Batch b = batch_of(h, nodes);
  h.r_cnt.reserve((size_t)b.total + 1, s);
  h.r_G.reserve((size_t)b.total + 2, s);
  h.scan_ws.reserve(scan_ws_words((size_t)b.total + 1), s);
  // row-block mode (stage_sharded.cu): the rows of the batch are spread over the ranks; each rank counts / fills the
  // rows it owns and two all-reduces give every rank the whole sub-mesh CSR
  const bool blk = h.block_mode && h.comm;
  const int row0 = blk ? h.blk_c0 : 0, row1 = blk ? h.blk_c1 : 0x7fffffff;
  h.stats.kernels_launched += launch_extract_count(h.r_list.p, h.r_vtx.p, b.count, b.total, h.d_node_ptr.p,
                                                   h.d_dofs.p, h.Mp, h.Mi, h.d_dof2node.p, h.r_cnt.p, s, row0, row1);
  if (blk && h.comm->world > 1 && b.total > 0) h.comm->all_reduce_sum(h.r_cnt.p, (size_t)b.total, s);
  h.stats.kernels_launched += exclusive_scan<0>(h.r_cnt.p, h.r_G.p, b.total, h.scan_ws.p, s);
  read_ints(h, h.r_G.p + b.total, &b.nnz, 1);
  h.r_sub.reserve((size_t)std::max(b.nnz, 1), s);
  if (blk && h.comm->world > 1 && b.nnz > 0) PB_CUDA(cudaMemsetAsync(h.r_sub.p, 0, sizeof(int) * (size_t)b.nnz, s));
  h.stats.kernels_launched += launch_extract_fill(h.r_list.p, h.r_vtx.p, b.count, b.total, h.d_node_ptr.p,
                                                  h.d_dofs.p, h.Mp, h.Mi, h.d_dof2node.p, h.r_G.p, h.r_sub.p, s, row0, row1);
  if (blk && h.comm->world > 1 && b.nnz > 0) h.comm->all_reduce_sum(h.r_sub.p, (size_t)b.nnz, s);
  return b;
}

// orders every graph of the batch held in r_vtx / r_G / r_sub into r_perm
void order_batch(parth_b200 &h, const Batch &b, const std::vector<int> &graph_nnz) {
  cudaStream_t s = h.stream;
  std::vector<long long> wsptr((size_t)b.count + 1, 0);
  // graphs that fit are ordered out of shared memory (their slab is only the fallback for rows that
  // turn out not to be symmetric); the staging area is sized by the largest of them
  const long long cap = amd_stage_capacity_ints();
  long long max_stage = 0;
  bool any_unstaged = false;
  for (int g = 0; g < b.count; g++) {
    const int n = b.vtx[g + 1] - b.vtx[g];
    wsptr[g + 1] = wsptr[g] + (graph_nnz[g] > 0 ? amd_slab_ints(n, graph_nnz[g]) : 0);
    if (n <= 0) continue;
    const long long need = amd_stage_need_ints(n, graph_nnz[g]);
    if (need <= cap) max_stage = std::max(max_stage, need); else any_unstaged = true;
  }
  const bool trace = std::getenv("PARTH_B200_TRACE") != nullptr;
  auto tick = [&]() { if (trace) cudaStreamSynchronize(s); return std::chrono::steady_clock::now(); };
  const auto ta = tick();
  const size_t cap_before = h.r_ws.cap;
  h.r_wsptr.reserve((size_t)b.count + 1, s);
  h.r_ws.reserve((size_t)std::max<long long>(wsptr[b.count], 1), s);
  h.r_perm.reserve((size_t)std::max(b.total, 1), s);
  PB_CUDA(cudaMemcpyAsync(h.r_wsptr.p, wsptr.data(), sizeof(long long) * ((size_t)b.count + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));
  const auto tb = tick();
  if (h.block_mode && h.comm && h.comm->world > 1 && b.count > 1) {
    // row-block mode: every rank holds every graph of the batch; they are dealt to the ranks largest first onto the
    // least loaded rank (same plan everywhere), ordered there, and the local permutations all-reduced (the other
    // ranks' segments are zero).  A single graph is ordered redundantly by every rank instead: no collective.
    const int world = h.comm->world;
    std::vector<int> order((size_t)b.count), mine;
    for (int g = 0; g < b.count; g++) order[g] = g;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      return b.vtx[x + 1] - b.vtx[x] > b.vtx[y + 1] - b.vtx[y];
    });
    std::vector<long long> load((size_t)world, 0);
    for (int g : order) {
      int best = 0;
      for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
      const long long n = b.vtx[g + 1] - b.vtx[g];
      load[best] += n * n / 64 + n;  // ordering cost grows faster than the size
      if (best == h.comm->rank) mine.push_back(g);
    }
    std::sort(mine.begin(), mine.end());
    h.x_glist.reserve(mine.size() + 1, s);
    if (!mine.empty())
      PB_CUDA(cudaMemcpyAsync(h.x_glist.p, mine.data(), sizeof(int) * mine.size(), cudaMemcpyHostToDevice, s));
    PB_CUDA(cudaMemsetAsync(h.r_perm.p, 0, sizeof(int) * (size_t)std::max(b.total, 1), s));
    PB_CUDA(cudaStreamSynchronize(s));  // pageable source
    h.stats.kernels_launched += launch_amd_batch(b.count, h.r_vtx.p, h.r_G.p, h.r_sub.p, h.r_wsptr.p, h.r_ws.p,
                                                 h.r_perm.p, max_stage, any_unstaged, s, h.x_glist.p, (int)mine.size());
    h.comm->all_reduce_sum(h.r_perm.p, (size_t)b.total, s);
  } else
  h.stats.kernels_launched += launch_amd_batch(b.count, h.r_vtx.p, h.r_G.p, h.r_sub.p, h.r_wsptr.p, h.r_ws.p,
                                               h.r_perm.p, max_stage, any_unstaged, s);
  if (trace) {
    const auto tc = tick();
    std::fprintf(stderr, "[parth_b200] order_batch: work-space %zu -> %zu ints (%.3f ms), kernels %.3f ms, staged need %lld ints%s\n",
                 cap_before, h.r_ws.cap, std::chrono::duration<double, std::milli>(tb - ta).count(),
                 std::chrono::duration<double, std::milli>(tc - tb).count(), max_stage, any_unstaged ? " + slab graphs" : "");
  }
}

std::vector<int> batch_graph_nnz(parth_b200 &h, const Batch &b) {
  // G at the graph boundaries -> entries per graph
  std::vector<int> at((size_t)b.count + 1), out((size_t)b.count);
  h.w5.reserve((size_t)b.count + 1, h.stream);
  // small strided read: gather on the host side from individual copies would cost count syncs,
  // so copy the whole prefix array when it is small, else gather with cudaMemcpy2D-free loop
  std::vector<int> G((size_t)b.total + 1);
  if (b.total + 1 > 0) {
    PB_CUDA(cudaMemcpyAsync(G.data(), h.r_G.p, sizeof(int) * ((size_t)b.total + 1), cudaMemcpyDeviceToHost, h.stream));
    PB_CUDA(cudaStreamSynchronize(h.stream));
    h.stats.d2h_bytes += (long long)sizeof(int) * (b.total + 1);
  }
  for (int g = 0; g < b.count; g++) out[g] = G[b.vtx[g + 1]] - G[b.vtx[g]];
  return out;
}

bool device_orders_with_amd(const parth_b200 &h) {
  // the device has one ordering kernel (AMD).  METIS_NodeND leaf ordering (ReorderingType METIS,
  // HMD.cpp:34-56) is third-party and not reproduced; under the REPAIR policy AMD stands in for
  // it, otherwise the request is refused so that parity runs never silently diverge.
  return h.force_amd || h.reorder_type == PARTH_B200_AMD || h.coarse_policy == PARTH_B200_COARSE_REPAIR ||
         h.coarse_policy == PARTH_B200_COARSE_RELOCATE;
}

}  // namespace

// sharded mode: deal the nodes to the ranks, largest first onto the least loaded rank (ties: lower
// node id, lower rank) -- a pure function of the replicated tree, so every rank derives the same plan
static std::vector<int> my_share(parth_b200 &h, const std::vector<int> &nodes) {
  std::vector<int> order(nodes.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  auto size_of = [&](int nd) { return h.node_ptr[nd + 1] - h.node_ptr[nd]; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const int sa = size_of(nodes[a]), sb = size_of(nodes[b]);
    return sa != sb ? sa > sb : nodes[a] < nodes[b];
  });
  std::vector<long long> load((size_t)h.world, 0);
  h.xchg_nodes = nodes;
  h.xchg_owner.assign(nodes.size(), 0);
  for (int i : order) {
    int best = 0;
    for (int r = 1; r < h.world; r++) if (load[r] < load[best]) best = r;
    h.xchg_owner[i] = best;
    load[best] += size_of(nodes[i]);
  }
  h.xchg_pending = true;
  std::vector<int> mine;
  for (size_t i = 0; i < nodes.size(); i++) if (h.xchg_owner[i] == h.rank) mine.push_back(nodes[i]);
  return mine;
}

// Integrator::permuteFineHMDNodes
void stage_permute_fine(parth_b200 &h, const std::vector<int> &all_nodes) {
  if (all_nodes.empty()) return;
  const std::vector<int> nodes = (h.world > 1 && !h.force_amd) ? my_share(h, all_nodes) : all_nodes;
  if (nodes.empty()) return;
  cudaStream_t s = h.stream;
  StageTimer tm(h, &h.stats.reorder_ms, /*accumulate=*/true);
  const bool trace = std::getenv("PARTH_B200_TRACE") != nullptr;
  auto now = [&]() { cudaStreamSynchronize(s); return std::chrono::steady_clock::now(); };
  auto t0 = trace ? now() : std::chrono::steady_clock::time_point();
  Batch b;
  const bool morton = h.reorder_type == PARTH_B200_MORTON_CODE && !h.force_amd;
  auto t1 = t0;
  if (morton) {
    // The reference prints "Morton code is not implemented yet" and leaves perm unset (HMD.cpp:82-83); the
    // device orders the node by the global Morton rank of its DOFs (morton.cu) -- no sub-mesh is needed.
    b = batch_of(h, nodes);
    stage_morton_nodes(h, b.count, b.total);
  } else {
    b = extract_fine(h, nodes);
    std::vector<int> gnnz = batch_graph_nnz(h, b);
    t1 = trace ? now() : t0;
    if (!device_orders_with_amd(h))
      for (int g = 0; g < b.count; g++)
        if (gnnz[g] > 0) {
          throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                            "permute_fine: METIS_NodeND leaf ordering is not available on the device "
                            "(use ReorderingType AMD or the REPAIR coarse policy)");
        }
    order_batch(h, b, gnnz);
  }
  auto t2 = trace ? now() : t0;
  h.stats.kernels_launched += launch_apply_order(h.r_list.p, h.r_vtx.p, b.count, b.total, h.d_node_ptr.p,
                                                 h.d_meta.p, h.d_dofs.p, h.r_perm.p, h.d_labels.p,
                                                 h.d_mesh_perm.p, s);
  tm.stop();
  if (trace) {
    auto t3 = now();
    auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
    int nmax = 0;
    for (int g = 0; g < b.count; g++) nmax = std::max(nmax, b.vtx[g + 1] - b.vtx[g]);
    std::fprintf(stderr, "[parth_b200] permute_fine: %d graphs (max n %d, %s): extract %.3f ms, order %.3f ms, apply %.3f ms\n",
                 b.count, nmax, morton ? "Morton" : "AMD", ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
}

// HMD::createMultiSubMesh for one node (coarse == 0) / HMD::getCoarseMesh (coarse == 1)
void stage_submesh(parth_b200 &h, int node, int coarse, int *n_out, int *nnz_out, int *Mp, int *Mi, int *l2g) {
  cudaStream_t s = h.stream;
  if (!coarse) {
    Batch b = extract_fine(h, std::vector<int>{node});
    *n_out = b.total;
    *nnz_out = b.nnz;
    if (Mp) read_ints(h, h.r_G.p, Mp, (size_t)b.total + 1);
    if (Mi && b.nnz > 0) read_ints(h, h.r_sub.p, Mi, (size_t)b.nnz);
    if (l2g && b.total > 0) read_ints(h, h.d_dofs.p + h.node_ptr[node], l2g, (size_t)b.total);
    return;
  }
  // pre-order list of the sub-tree's nodes (HMD::getAssignedDOFsOfCoarseHMD, HMD.cpp:434-447)
  std::vector<int> list, stack{node};
  while (!stack.empty()) {
    int x = stack.back();
    stack.pop_back();
    if (h.meta[x * 7 + 2] == -1) continue;
    list.push_back(x);
    if (h.meta[x * 7 + 1] != -1) stack.push_back(h.meta[x * 7 + 1]);
    if (h.meta[x * 7] != -1) stack.push_back(h.meta[x * 7]);
  }
  const int M = h.M_n;
  h.r_list.reserve(list.size() + 1, s);
  h.r_chosen.reserve((size_t)M + 1, s);
  h.r_g2l.reserve((size_t)M + 2, s);
  h.scan_ws.reserve(scan_ws_words((size_t)M + 1), s);
  PB_CUDA(cudaMemcpyAsync(h.r_list.p, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));
  PB_CUDA(cudaMemsetAsync(h.r_chosen.p, 0, sizeof(int) * (size_t)M, s));
  h.stats.kernels_launched += launch_mark_nodes(h.r_list.p, (int)list.size(), h.d_node_ptr.p, h.d_dofs.p, h.r_chosen.p, s);
  h.stats.kernels_launched += exclusive_scan<0>(h.r_chosen.p, h.r_g2l.p, M, h.scan_ws.p, s);
  int n = 0;
  read_ints(h, h.r_g2l.p + M, &n, 1);
  h.r_l2g.reserve((size_t)std::max(n, 1), s);
  h.r_cnt.reserve((size_t)n + 1, s);
  h.r_G.reserve((size_t)n + 2, s);
  h.stats.kernels_launched += launch_compact_chosen(h.r_chosen.p, h.r_g2l.p, M, h.r_l2g.p, s);
  h.stats.kernels_launched += launch_marker_count(h.r_l2g.p, n, h.Mp, h.Mi, h.r_chosen.p, h.r_cnt.p, s);
  h.stats.kernels_launched += exclusive_scan<0>(h.r_cnt.p, h.r_G.p, n, h.scan_ws.p, s);
  int nnz = 0;
  read_ints(h, h.r_G.p + n, &nnz, 1);
  h.r_sub.reserve((size_t)std::max(nnz, 1), s);
  h.stats.kernels_launched += launch_marker_fill(h.r_l2g.p, n, h.Mp, h.Mi, h.r_chosen.p, h.r_g2l.p, h.r_G.p, h.r_sub.p, s);
  *n_out = n;
  *nnz_out = nnz;
  if (Mp) read_ints(h, h.r_G.p, Mp, (size_t)n + 1);
  if (Mi && nnz > 0) read_ints(h, h.r_sub.p, Mi, (size_t)nnz);
  if (l2g && n > 0) read_ints(h, h.r_l2g.p, l2g, (size_t)n);
}

// multi-GPU exchange helpers: label slices of the listed nodes <-> a concatenated device buffer
void stage_pack_labels(parth_b200 &h, const int *nodes_in, int n, int *dbuf, bool unpack) {
  if (n <= 0) return;
  std::vector<int> nodes(nodes_in, nodes_in + n);
  cudaStream_t s = h.stream;
  std::vector<int> vtx((size_t)n + 1, 0);
  for (int g = 0; g < n; g++) {
    if (nodes[g] < 0 || nodes[g] >= h.num_nodes) throw StatusError(PARTH_B200_ERR_ARG, "pack_labels: node id out of range");
    vtx[g + 1] = vtx[g] + (h.node_ptr[nodes[g] + 1] - h.node_ptr[nodes[g]]);
  }
  h.r_list.reserve((size_t)n + 1, s);
  h.r_vtx.reserve((size_t)n + 2, s);
  PB_CUDA(cudaMemcpyAsync(h.r_list.p, nodes.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.r_vtx.p, vtx.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));
  if (unpack)
    h.stats.kernels_launched += launch_unpack_labels(h.r_list.p, h.r_vtx.p, n, vtx[n], h.d_node_ptr.p, h.d_meta.p,
                                                     h.d_dofs.p, dbuf, h.d_labels.p, h.d_mesh_perm.p, s);
  else
    h.stats.kernels_launched += launch_pack_labels(h.r_list.p, h.r_vtx.p, n, vtx[n], h.d_node_ptr.p, h.d_labels.p, dbuf, s);
}

// Dirty COARSE nodes.  The reference re-dissects the sub-tree with METIS_ComputeVertexSeparator
// (Integrator::rePartitionCoarseHMDNode, Integrator.cpp:812-849), which is third-party and not
// reproduced on the device.  REPAIR restores the separator property instead: one endpoint of every
// broken edge moves into the common separator of its two nodes (the reference's own relocation
// rule, Integrator.cpp:377-544, applied without a level limit) and the nodes that received DOFs
// are re-ordered.  The reported dirty sets / reuse / changed-edge list stay the reference's.
int stage_coarse(parth_b200 &h) {
  if (h.coarse_policy == PARTH_B200_COARSE_CALLBACK)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "coarse policy CALLBACK: no re-dissection callback registered");
  std::vector<int> touched;
  stage_relocate(h, 1 << 29, touched, /*filter_edges=*/false);
  std::vector<int> nodes = h.dirty_fine;
  // RELOCATE: the labels the relocation left (survivors in their old order, joiners appended) stand; stage_relocate
  // has already re-assembled mesh_perm from them
  if (h.coarse_policy != PARTH_B200_COARSE_RELOCATE) nodes.insert(nodes.end(), touched.begin(), touched.end());
  std::sort(nodes.begin(), nodes.end());
  nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
  stage_permute_fine(h, nodes);
  return PARTH_B200_OK;
}

// HMD::AMDNodeNDPermutation on a batch of host graphs (sorted, duplicate-free rows)
void stage_amd_batch(parth_b200 &h, int count, const int *graph_ptr, const int *Mp, const int *Mi, int *perm) {
  cudaStream_t s = h.stream;
  if (count == 0) return;
  Batch b;
  b.count = count;
  b.vtx.assign(graph_ptr, graph_ptr + count + 1);
  b.total = b.vtx[count];
  std::vector<int> G((size_t)b.total + 1), gnnz((size_t)count);
  long long base = 0;
  const int *mp = Mp;
  for (int g = 0; g < count; g++) {
    const int n = b.vtx[g + 1] - b.vtx[g];
    for (int k = 0; k < n; k++) G[b.vtx[g] + k] = (int)(base + mp[k]);
    gnnz[g] = mp[n];
    base += mp[n];
    mp += n + 1;
  }
  G[b.total] = (int)base;
  b.nnz = (int)base;
  h.r_vtx.reserve((size_t)count + 2, s);
  h.r_G.reserve((size_t)b.total + 2, s);
  h.r_sub.reserve((size_t)std::max(b.nnz, 1), s);
  PB_CUDA(cudaMemcpyAsync(h.r_vtx.p, b.vtx.data(), sizeof(int) * ((size_t)count + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.r_G.p, G.data(), sizeof(int) * ((size_t)b.total + 1), cudaMemcpyHostToDevice, s));
  if (b.nnz > 0)
    PB_CUDA(cudaMemcpyAsync(h.r_sub.p, Mi, sizeof(int) * (size_t)b.nnz, cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaStreamSynchronize(s));
  order_batch(h, b, gnnz);
  if (b.total > 0) {
    PB_CUDA(cudaMemcpyAsync(perm, h.r_perm.p, sizeof(int) * (size_t)b.total, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
  }
}

}  // namespace pb
