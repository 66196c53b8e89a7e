// parth_b200/csrc/stage_perm.cu -- host drivers: tree import/export, change detection, dirty
// sets, permutation assembly / expansion, and the computePermutation control flow
// (reference src/Parth/ParthAPI.cpp:191-279, src/Parth/Integrator.cpp:546-683).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "scan.cuh"
#include "stages.cuh"
#include "tree.cuh"

namespace pb {

// ------------------------------------------------------------------ tree fixture in / out
static void compute_visit(parth_b200 &h, std::vector<int> &visit) {
  // nodes reached by the recursion from the root through stored left/right links
  // (HMD::assignCoarseLocalPermToGlobal, HMD.cpp:348-363)
  visit.assign(h.num_nodes, 0);
  if (h.num_nodes == 0 || h.meta[2] == -1) return;
  std::vector<int> stack{0};
  while (!stack.empty()) {
    int x = stack.back();
    stack.pop_back();
    visit[x] = 1;
    int l = h.meta[x * 7 + 0], r = h.meta[x * 7 + 1];
    if (l != -1) stack.push_back(l);
    if (r != -1) stack.push_back(r);
  }
}

void upload_tree_meta(parth_b200 &h, bool sync) {
  cudaStream_t s = h.stream;
  const int nodes = h.num_nodes;
  h.meta_heap = true;
  for (int x = 0; x < nodes && h.meta_heap; x++) {
    if (h.meta[x * 7 + 2] == -1) continue;
    int depth = 0;
    for (int y = x; y > 0; y = (y - 1) >> 1) depth++;
    if (h.meta[x * 7 + 2] != x || h.meta[x * 7 + 5] != depth || (x > 0 && h.meta[x * 7 + 3] != (x - 1) / 2))
      h.meta_heap = false;
  }
  h.d_meta.reserve((size_t)nodes * 7, s);
  h.d_node_ptr.reserve((size_t)nodes + 1, s);
  h.d_visit.reserve((size_t)nodes, s);
  std::vector<int> visit;
  compute_visit(h, visit);
  // pageable sources: cudaMemcpyAsync returns after staging them, so the vectors may change later
  PB_CUDA(cudaMemcpyAsync(h.d_meta.p, h.meta.data(), sizeof(int) * (size_t)nodes * 7, cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.d_node_ptr.p, h.node_ptr.data(), sizeof(int) * ((size_t)nodes + 1), cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.d_visit.p, visit.data(), sizeof(int) * (size_t)nodes, cudaMemcpyHostToDevice, s));
  if (sync) PB_CUDA(cudaStreamSynchronize(s));
}

void assemble_all(parth_b200 &h) {
  h.d_mesh_perm.reserve((size_t)std::max(h.tree_M, 1), h.stream);
  h.stats.kernels_launched += launch_assemble_all(h.d_node_ptr.p, h.d_meta.p, h.d_visit.p, h.num_nodes,
                                                  h.d_dofs.p, h.d_labels.p, h.node_ptr[h.num_nodes],
                                                  h.d_mesh_perm.p, h.stream);
}

void stage_import_tree(parth_b200 &h, int M_n, int levels, int nodes, const int *meta,
                       const int *node_ptr, const int *dofs, const int *labels, const int *Mp_prev,
                       const int *Mi_prev) {
  cudaStream_t s = h.stream;
  // the mailbox, the node flags and the host-side dirty-set code are sized for <= 8191 nodes (12 levels, what
  // HMD::initHMD clamps to, HMD.cpp:98-100)
  if (nodes <= 0 || nodes > 8191 || M_n <= 0 || node_ptr[0] != 0 || node_ptr[nodes] > M_n)
    throw StatusError(PARTH_B200_ERR_ARG, "import_tree: inconsistent sizes (1 <= nodes <= 8191, node_ptr[nodes] <= M_n)");
  for (int x = 0; x < nodes; x++) {
    if (node_ptr[x + 1] < node_ptr[x]) throw StatusError(PARTH_B200_ERR_ARG, "import_tree: node_ptr is not monotone");
    for (int k = 0; k < 2; k++) {  // left / right links
      const int c = meta[(size_t)x * 7 + k];
      if (c < -1 || c >= nodes) throw StatusError(PARTH_B200_ERR_ARG, "import_tree: child link out of range");
    }
    const int par = meta[(size_t)x * 7 + 3];
    if (par < -1 || par >= nodes) throw StatusError(PARTH_B200_ERR_ARG, "import_tree: parent link out of range");
  }
  {
    const int total_in = node_ptr[nodes];
    for (int k = 0; k < total_in; k++)
      if (dofs[k] < 0 || dofs[k] >= M_n) throw StatusError(PARTH_B200_ERR_ARG, "import_tree: DOF id out of range");
  }
  if (Mp_prev && Mi_prev && (Mp_prev[0] != 0 || Mp_prev[M_n] < 0))
    throw StatusError(PARTH_B200_ERR_ARG, "import_tree: snapshot row pointers are inconsistent");
  h.num_levels = levels;
  h.nd_levels = levels;
  h.num_nodes = nodes;
  h.tree_M = M_n;
  h.meta.assign(meta, meta + (size_t)nodes * 7);
  h.node_ptr.assign(node_ptr, node_ptr + nodes + 1);
  const int total = node_ptr[nodes];
  h.d_dofs.reserve((size_t)M_n, s);
  h.d_labels.reserve((size_t)M_n, s);
  h.d_dof2node.reserve((size_t)M_n, s);
  PB_CUDA(cudaMemcpyAsync(h.d_dofs.p, dofs, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, s));
  PB_CUDA(cudaMemcpyAsync(h.d_labels.p, labels, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, s));
  upload_tree_meta(h);
  PB_CUDA(cudaMemsetAsync(h.d_dof2node.p, 0, sizeof(int) * (size_t)M_n, s));
  h.stats.kernels_launched += launch_dof2node(h.d_node_ptr.p, nodes, h.d_dofs.p, total, h.d_dof2node.p, s);
  h.d_mesh_perm.reserve((size_t)M_n, s);
  PB_CUDA(cudaMemsetAsync(h.d_mesh_perm.p, 0, sizeof(int) * (size_t)M_n, s));
  assemble_all(h);
  relocate_prepare(h, M_n);  // the first re-ordering update must not pay for buffer growth
  h.tree_init = true;
  if (Mp_prev && Mi_prev) {
    const int nnz = Mp_prev[M_n];
    if (h.Mp && (h.Mp == h.Mp_prev.p || h.Mi == h.Mi_prev.p)) {
      // the current graph aliases the snapshot buffers (right after a computePermutation): they are about to
      // hold the imported snapshot, so there is no current graph until the next setMatrix / setMesh
      h.Mp = h.Mi = nullptr;
      h.M_n = h.nnzM = 0;
      h.cur_sym_proven = false;
    }
    h.Mp_prev.reserve((size_t)M_n + 1, s);
    h.Mi_prev.reserve((size_t)std::max(nnz, 1), s);
    PB_CUDA(cudaMemcpyAsync(h.Mp_prev.p, Mp_prev, sizeof(int) * ((size_t)M_n + 1), cudaMemcpyHostToDevice, s));
    if (nnz > 0)
      PB_CUDA(cudaMemcpyAsync(h.Mi_prev.p, Mi_prev, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, s));
    h.M_n_prev = M_n;
    h.nnz_prev = nnz;
    h.prev_sym_proven = false;  // caller-provided snapshot: nothing is known about it
    h.diff_cached = false;
    h.spec_valid = false;
  }
  PB_CUDA(cudaStreamSynchronize(s));
}

void stage_export_tree(parth_b200 &h, int *meta, int *node_ptr, int *dofs, int *labels, int *dof2node,
                       int *mesh_perm) {
  cudaStream_t s = h.stream;
  if (!h.tree_init) throw StatusError(PARTH_B200_ERR_NO_TREE, "export_tree: no tree");
  const int total = h.node_ptr[h.num_nodes];
  std::memcpy(meta, h.meta.data(), sizeof(int) * h.meta.size());
  std::memcpy(node_ptr, h.node_ptr.data(), sizeof(int) * h.node_ptr.size());
  if (dofs) PB_CUDA(cudaMemcpyAsync(dofs, h.d_dofs.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, s));
  if (labels) PB_CUDA(cudaMemcpyAsync(labels, h.d_labels.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, s));
  if (dof2node)
    PB_CUDA(cudaMemcpyAsync(dof2node, h.d_dof2node.p, sizeof(int) * (size_t)h.tree_M, cudaMemcpyDeviceToHost, s));
  if (mesh_perm)
    PB_CUDA(cudaMemcpyAsync(mesh_perm, h.d_mesh_perm.p, sizeof(int) * (size_t)h.tree_M, cudaMemcpyDeviceToHost, s));
  PB_CUDA(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------ (b) change detection
// Integrator::computeChangedEdges (Integrator.cpp:259-338).  Result: h.d_changed / h.n_changed.
// The node flags of findRegionConnections are computed in the same enqueue and fetched with the
// edge count in ONE read-back (h.node_flags); stage_dirty_sets consumes them.
void stage_detect(parth_b200 &h) {
  cudaStream_t s = h.stream;
  const int M = h.M_n, nn = h.num_nodes;
  h.n_changed = 0;
  h.node_flags_valid = false;
  if (h.spec_valid && h.diff_cached && h.map_len == 0 && (h.Mp == h.Mp_own.p || h.block_mode) &&
      (int)h.spec_flags.size() == nn * 2) {
    // stage (a) enqueued the edge kernels right behind its row diff and fetched their result with
    // its own mailbox: nothing to launch, nothing to wait for, nothing to time
    h.spec_valid = false;
    h.n_changed = h.spec_n_changed;
    if (h.n_chg_rows > 0) { h.node_flags.swap(h.spec_flags); h.node_flags_valid = true; }
    h.stats.detect_ms = 0;
    h.timer_ms[2] = 0;
    return;
  }
  if (h.block_mode)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                      "row-block mode: change detection runs with the build (set_matrix); there is no snapshot of this graph "
                      "to compare with (call parth_b200_accept_snapshot after the first matrix)");
  StageTimer tm(h, &h.stats.detect_ms, false, &h.timer_ms[2]);
  h.spec_valid = false;
  const int n_words = (M + 31) / 32;
  h.scan_ws.reserve(scan_ws_words((size_t)std::max(n_words, M) + 1), s);
  int n_rows = 0;
  const int *rows = nullptr;
  if (h.diff_cached && h.map_len == 0 && h.Mp == h.Mp_own.p) {
    // stage (a) already diffed this graph against the snapshot for its symmetry proof
    n_rows = h.n_chg_rows;
    rows = h.chg_rows.p;
  } else {
    h.w0.reserve((size_t)n_words + 1, s);  // row bitmap
    h.w1.reserve((size_t)n_words + 2, s);  // popc prefix
    unsigned *bits = reinterpret_cast<unsigned *>(h.w0.p);
    h.mail_clean = false;  // the diff kernels count into the mailbox
    if (h.map_len > 0)
      h.stats.kernels_launched += launch_diff_rows_mapped(h.Mp, h.Mi, h.Mp_prev.p, h.Mi_prev.p, h.new_to_old.p,
                                                          h.old_to_new.p, M, bits, h.dcnt_p(), s);
    else
      h.stats.kernels_launched += launch_diff_rows(h.Mp, h.Mi, h.Mp_prev.p, h.Mi_prev.p, M, bits, h.dcnt_p(), s);
    int cnt[4];
    read_ints(h, reinterpret_cast<const int *>(h.dcnt_p()), cnt, 4);
    n_rows = cnt[0];
    if (n_rows > 0) {
      h.stats.kernels_launched += exclusive_scan<1>(h.w0.p, h.w1.p, n_words, h.scan_ws.p, s);
      h.w2.reserve((size_t)n_rows + 1, s);  // changed rows, ascending
      h.stats.kernels_launched += launch_list_changed_rows(bits, h.w1.p, n_words, M, h.w2.p, n_rows, s);
    }
    rows = h.w2.p;
  }
  if (n_rows > 0) {
    h.w3.reserve((size_t)n_rows + 2, s);  // per-row edge counts -> offsets
    h.w4.reserve((size_t)n_rows + 2, s);
    h.w5.reserve((size_t)nn * 2 + 2, s);  // [0] edge count, [2 .. 2+nn) within-node flags, then LCA flags
    h.stats.kernels_launched += launch_count_row_edges(rows, n_rows, nullptr, h.Mp, h.Mi, h.d_dof2node.p, h.w3.p, s);
    h.stats.kernels_launched += exclusive_scan<0>(h.w3.p, h.w4.p, n_rows, h.scan_ws.p, s);
    if (h.d_changed.cap < 2 * 65536) h.d_changed.reserve(2 * 65536, s);
    std::vector<int> fl((size_t)nn * 2 + 2);
    for (int attempt = 0; attempt < 2; attempt++) {
      const int cap = (int)(h.d_changed.cap / 2);
      PB_CUDA(cudaMemsetAsync(h.w5.p, 0, sizeof(int) * ((size_t)nn * 2 + 2), s));
      h.stats.kernels_launched += launch_emit_row_edges(rows, n_rows, nullptr, h.Mp, h.Mi, h.d_dof2node.p, h.w4.p,
                                                        h.d_changed.p, cap, s);
      h.stats.kernels_launched += launch_mark_dirty_nodes_dev(h.d_changed.p, h.w4.p + n_rows, cap, h.d_dof2node.p,
                                                              h.d_meta.p, h.w5.p + 2, h.w5.p + 2 + nn, h.w5.p, s);
      read_ints(h, h.w5.p, fl.data(), fl.size());
      if (fl[0] <= cap) break;
      h.d_changed.reserve((size_t)fl[0] * 2 + 64, s);  // the list did not fit: grow and emit again
    }
    h.n_changed = fl[0];
    h.node_flags.assign(fl.begin() + 2, fl.end());
    h.node_flags_valid = true;
  }
  tm.stop();
}

// ------------------------------------------------------------------ (b) dirty sets on 2^(L+1)-1 nodes
static bool node_is_leaf(const parth_b200 &h, int x) { return h.meta[x * 7] == -1 && h.meta[x * 7 + 1] == -1; }

static void unmark_children(const parth_b200 &h, int x, std::vector<int> &marker) {
  // HMD::unMarkLeftRightSubMeshes (HMD.cpp:421-432)
  std::vector<int> stack{x};
  while (!stack.empty()) {
    int y = stack.back();
    stack.pop_back();
    int l = h.meta[y * 7], r = h.meta[y * 7 + 1];
    if (l != -1) { marker[l] = 0; stack.push_back(l); }
    if (r != -1) { marker[r] = 0; stack.push_back(r); }
  }
}

static long long subtree_size(const parth_b200 &h, int x) {
  // |HMD::getAssignedDOFsOfCoarseHMD| (HMD.cpp:434-447)
  long long tot = 0;
  std::vector<int> stack{x};
  while (!stack.empty()) {
    int y = stack.back();
    stack.pop_back();
    tot += h.node_ptr[y + 1] - h.node_ptr[y];
    int l = h.meta[y * 7], r = h.meta[y * 7 + 1];
    if (l != -1) stack.push_back(l);
    if (r != -1) stack.push_back(r);
  }
  return tot;
}

// findRegionConnections + filterChanges + lag filter + getReuseRatio
// (Integrator.cpp:685-810, 616-647).  `pre_fine` = nodes already queued by the relocation step.
void stage_dirty_sets(parth_b200 &h, const std::vector<int> &pre_fine) {
  cudaStream_t s = h.stream;
  const int nn = h.num_nodes;
  std::vector<int> flags;
  if (h.node_flags_valid && (int)h.node_flags.size() == nn * 2) {
    flags.swap(h.node_flags);  // fetched together with the edge list by stage_detect
    h.node_flags_valid = false;
    h.stats.dirty_ms = 0;  // host work only: no device span to measure
  } else {
    StageTimer tm(h, &h.stats.dirty_ms);
    // the edge list or the tree changed since detection (relocation): classify again
    h.w5.reserve((size_t)nn * 2 + 2, s);
    PB_CUDA(cudaMemsetAsync(h.w5.p, 0, sizeof(int) * (size_t)nn * 2, s));
    h.stats.kernels_launched += launch_mark_dirty_nodes(h.d_changed.p, h.n_changed, h.d_dof2node.p, h.d_meta.p,
                                                        h.w5.p, h.w5.p + nn, s);
    flags.resize((size_t)nn * 2);
    tm.stop();
    read_ints(h, h.w5.p, flags.data(), (size_t)nn * 2);
  }
  const int *within = flags.data(), *lca = flags.data() + nn;

  h.dirty_fine = pre_fine;
  for (int x = 0; x < nn; x++) if (within[x]) h.dirty_fine.push_back(x);
  if (h.aggressive) {
    std::sort(h.dirty_fine.begin(), h.dirty_fine.end());
    h.dirty_fine.erase(std::unique(h.dirty_fine.begin(), h.dirty_fine.end()), h.dirty_fine.end());
  }
  std::vector<int> coarse;
  for (int x = 0; x < nn; x++) if (lca[x]) coarse.push_back(x);
  std::vector<int> cflag(nn, 0), fflag(nn, 0);
  for (int x : coarse) { cflag[x] = 1; fflag[x] = 1; }
  for (int x : h.dirty_fine) fflag[x] = 1;
  for (int x : coarse) {
    if (cflag[x] == 1) { unmark_children(h, x, cflag); unmark_children(h, x, fflag); }
    fflag[x] = 0;
  }
  h.dirty_coarse.clear();
  for (int x = 0; x < nn; x++) if (cflag[x] == 1) h.dirty_coarse.push_back(x);
  {
    std::vector<int> keep;
    for (int x : h.dirty_fine) if (fflag[x] == 1) keep.push_back(x);
    h.dirty_fine.swap(keep);
  }
  if (h.lagging) {
    std::vector<int> keep;
    for (int x : h.dirty_fine) if (!node_is_leaf(h, x)) keep.push_back(x);
    h.dirty_fine.swap(keep);
    const int lvl = ((int)std::pow(2, h.lag_lvl + 1) - 1) - 1;  // getLastHMDNodeInLevel(lag_lvl) - 1
    keep.clear();
    for (int x : h.dirty_coarse) if (x < lvl) keep.push_back(x);
    h.dirty_coarse.swap(keep);
  }
  h.dirty_coarse_saved = h.dirty_coarse;
  h.dirty_fine_saved = h.dirty_fine;
  double nodes = 0;
  for (int x : h.dirty_coarse) nodes += (double)subtree_size(h, x);
  for (int x : h.dirty_fine) nodes += h.node_ptr[x + 1] - h.node_ptr[x];
  h.reuse = h.tree_M != 0 ? 1 - (nodes * 1.0 / h.tree_M) : 0.0;
}

// ------------------------------------------------------------------ (e) expansion
void stage_expand(parth_b200 &h, const int *dmesh_perm, int n, int dim, int *dout) {
  StageTimer tm(h, &h.stats.expand_ms);
  h.stats.kernels_launched += launch_expand(dmesh_perm, n, dim, dout, h.stream);
  tm.stop();
}

// (a19) snapshot, ParthAPI.cpp:270-278: a buffer swap when the graph is ours, a copy when borrowed
void stage_snapshot(parth_b200 &h) {
  cudaStream_t s = h.stream;
  if (h.block_mode) {
    // row-block mode: the rank's rows of the two graphs trade places, with the offsets that bias their pointers
    if (h.Mp == h.Mp_own_b()) {
      h.Mp_own.swap(h.Mp_prev);
      h.Mi_own.swap(h.Mi_prev);
      std::swap(h.own_off, h.prev_off);
      std::swap(h.own_q0, h.prev_q0);
      std::swap(h.own_q1, h.prev_q1);
    }
    h.prev_sym_proven = h.cur_sym_proven;
    h.diff_cached = false;
    h.spec_valid = false;
    h.Mp = h.Mp_prev_b();
    h.Mi = h.Mi_prev_b();
    h.M_n_prev = h.M_n;
    h.nnz_prev = h.nnzM;
    h.map_len = 0;
    h.n_added = h.n_removed = 0;
    return;
  }
  if (h.mesh_owned && h.Mp == h.Mp_own.p) {
    h.Mp_own.swap(h.Mp_prev);
    h.Mi_own.swap(h.Mi_prev);
  } else if (h.Mp != h.Mp_prev.p) {
    h.Mp_prev.reserve((size_t)h.M_n + 1, s);
    h.Mi_prev.reserve((size_t)std::max(h.nnzM, 1), s);
    PB_CUDA(cudaMemcpyAsync(h.Mp_prev.p, h.Mp, sizeof(int) * ((size_t)h.M_n + 1), cudaMemcpyDeviceToDevice, s));
    if (h.nnzM > 0)
      PB_CUDA(cudaMemcpyAsync(h.Mi_prev.p, h.Mi, sizeof(int) * (size_t)h.nnzM, cudaMemcpyDeviceToDevice, s));
  }
  h.prev_sym_proven = h.cur_sym_proven;
  h.diff_cached = false;
  h.Mp = h.Mp_prev.p;  // the current graph and the snapshot are now the same arrays
  h.Mi = h.Mi_prev.p;
  h.mesh_owned = true;
  h.M_n_prev = h.M_n;
  h.nnz_prev = h.nnzM;
  h.map_len = 0;
  h.n_added = h.n_removed = 0;
}

// cold start (HMD::initHMD, HMD.cpp:89-120): the separator tree comes from the host decomposer
static void cold_start(parth_b200 &h) {
  if (!h.decomposer && !h.builtin_decomposer)
    throw StatusError(PARTH_B200_ERR_NO_TREE,
                      "computePermutation: no separator tree (import one or set a decomposer)");
  cudaStream_t s = h.stream;
  int levels = h.nd_levels;
  if (levels > 12) levels = 10;  // HMD.cpp:98-100
  const int nodes = (int)std::pow(2, levels + 1) - 1;
  std::vector<int> Mp((size_t)h.M_n + 1), Mi((size_t)std::max(h.nnzM, 1));
  PB_CUDA(cudaMemcpyAsync(Mp.data(), h.Mp, sizeof(int) * Mp.size(), cudaMemcpyDeviceToHost, s));
  if (h.nnzM > 0)
    PB_CUDA(cudaMemcpyAsync(Mi.data(), h.Mi, sizeof(int) * (size_t)h.nnzM, cudaMemcpyDeviceToHost, s));
  PB_CUDA(cudaStreamSynchronize(s));
  std::vector<int> meta((size_t)nodes * 7, -1), node_ptr((size_t)nodes + 1, 0), dofs((size_t)h.M_n), labels((size_t)h.M_n);
  const bool builtin = !h.decomposer;
  int rc = builtin ? builtin_decomposer(nullptr, h.M_n, Mp.data(), Mi.data(), levels, nodes, meta.data(),
                                        node_ptr.data(), dofs.data(), labels.data())
                   : h.decomposer(h.decomposer_user, h.M_n, Mp.data(), Mi.data(), levels, nodes, meta.data(),
                                  node_ptr.data(), dofs.data(), labels.data());
  if (rc != 0) throw StatusError(PARTH_B200_ERR_NO_TREE, "decomposer callback failed");
  stage_import_tree(h, h.M_n, levels, nodes, meta.data(), node_ptr.data(), dofs.data(), labels.data(), nullptr, nullptr);
  if (builtin) {
    // the built-in dissection hands back identity labels: order every node on the device
    // (HMD::Decompose orders leaves and separators as it goes, HMD.cpp:248-250,297-301)
    std::vector<int> all;
    for (int x = 0; x < nodes; x++) if (node_ptr[x + 1] > node_ptr[x]) all.push_back(x);
    // (ReorderingType MORTON_CODE with coordinates for every DOF orders them along the Z curve instead)
    h.force_amd = !(h.reorder_type == PARTH_B200_MORTON_CODE && h.morton_n == h.M_n);
    // a cold start orders EVERY node on every rank (the replicated trees must agree before any label exchange):
    // the per-node sharing of the warm path is switched off for this call
    const int world = h.world;
    h.world = 1;
    try { stage_permute_fine(h, all); } catch (...) { h.force_amd = false; h.world = world; throw; }
    h.force_amd = false;
    h.world = world;
  }
}

// Integrator::getGlobalPerm (Integrator.cpp:546-683) after detection found changes
static int global_perm(parth_b200 &h) {
  std::vector<int> pre_fine;
  if (h.aggressive) {
    const int thr = (int)std::pow(2, h.relocate_lvl + 1) - 1;  // getLastHMDNodeInLevel(relocate_lvl)
    stage_relocate(h, thr, pre_fine);
    h.node_flags_valid = false;  // the edge list and the tree have changed
  }
  stage_dirty_sets(h, pre_fine);
  const bool full = h.reuse < 0.2;
  if (full) {
    h.dirty_fine.clear();
    h.dirty_coarse.assign(1, 0);
    h.reuse = 0;
  }
  if (!h.dirty_coarse.empty()) {
    if (h.coarse_policy == PARTH_B200_COARSE_REPORT) return PARTH_B200_NEEDS_REDISSECT;
    // the reference orders the fine nodes first (Integrator.cpp:659); REPAIR re-orders them again
    // together with the touched separators after the relocation, so once is enough
    return stage_coarse(h);
  }
  stage_permute_fine(h, h.dirty_fine);
  return PARTH_B200_OK;
}

// ParthAPI::computePermutation (ParthAPI.cpp:191-279); dperm: device buffer of M_n*dim ints
int stage_compute_permutation(parth_b200 &h, int *dperm, int dim) {
  // A build left in flight by the device-pointer set_matrix (steady state: tree present, same vertex
  // set, no DOF map): expand the CURRENT mesh permutation behind it before waiting for its flags.
  // Most reuse updates leave the permutation untouched (no change, edges inside a node or towards an
  // ancestor, lag-dropped nodes): the expansion then already ran while the host was reading the
  // flags; when nodes are re-ordered after all it is simply run again at the end.
  bool spec_expanded = false;
  cudaEvent_t spec_end = nullptr;
  h.stats.expand_speculative = 0;
  if (h.pend.active) {
    if (h.tree_init && h.tree_M == h.pend.M && h.map_len == 0 && dperm) {
      StageTimer tm(h, &h.stats.expand_ms, false, nullptr, h.ev_mail);  // bracket opens at the build's mailbox event
      h.stats.kernels_launched += launch_expand(h.d_mesh_perm.p, h.pend.M, dim, dperm, h.stream);
      spec_end = tm.stop();
      spec_expanded = true;
    }
    PB_MARK("cp: speculative expansion enqueued");
    stage_build_settle(h);
    PB_MARK("cp: build settled (mailbox read)");
  }
  for (double &t : h.timer_ms) t = 0;
  h.xchg_pending = false;
  h.xchg_nodes.clear();
  h.xchg_owner.clear();
  h.reuse = 0;
  h.n_changed = 0;
  h.sim_dim = dim;
  h.stats.detect_ms = h.stats.integrate_ms = h.stats.dirty_ms = h.stats.reorder_ms = 0;
  h.stats.assemble_ms = h.stats.expand_ms = 0;
  if (!h.Mp || h.M_n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "computePermutation: no mesh set");
  if (h.block_mode && (!h.tree_init || h.M_n != h.M_n_prev || h.tree_M != h.M_n))
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED,
                      "computePermutation (row-block mode): the cold start needs the whole graph on one rank; import the "
                      "separator tree and call parth_b200_accept_snapshot after the first matrix");
  if (h.M_n != h.M_n_prev && h.map_len == 0) h.tree_init = false;
  int rc = PARTH_B200_OK;
  if (!h.tree_init) {
    cold_start(h);  // host callback: init_time stays 0 (not device time)
    StageTimer tm(h, &h.stats.expand_ms, false, &h.timer_ms[4]);
    h.stats.kernels_launched += launch_expand(h.d_mesh_perm.p, h.M_n, dim, dperm, h.stream);
    tm.stop();
    h.reuse = 0;
  } else {
    if (h.map_len > 0) stage_integrate(h);
    stage_detect(h);
    PB_MARK("cp: detect done");
    // compute_permutation bracket: opens at the end of the speculative expansion when there is one
    StageTimer tc(h, &h.timer_ms[3], false, nullptr, spec_end);
    const long long launched0 = h.stats.kernels_launched;
    const bool no_change = h.n_changed == 0;  // (the relocation step filters the edge list later on)
    if (no_change) {
      h.reuse = 1;
    } else {
      rc = global_perm(h);
      PB_MARK("cp: global_perm done");
    }
    if (rc == PARTH_B200_OK && h.xchg_pending) {
      // sharded mode: the other ranks' nodes are still missing; the caller gathers them, then
      // stage_finish_permutation expands and snapshots
      tc.stop();
      return PARTH_B200_NEEDS_EXCHANGE;
    }
    // the speculative expansion stands when nothing touched the tree or the permutation
    const bool untouched = spec_expanded && h.tree_init && h.map_len == 0 &&
                           (no_change || (!h.aggressive && h.dirty_fine.empty() && h.dirty_coarse.empty()));
    if (rc == PARTH_B200_OK && untouched) {
      h.stats.expand_speculative = 1;  // its timer (pending) fills expand_ms
    } else if (rc == PARTH_B200_OK) {
      stage_expand(h, h.d_mesh_perm.p, h.M_n, dim, dperm);
    } else if (rc == PARTH_B200_NEEDS_REDISSECT) {
      // REPORT policy: nothing was re-ordered.  The caller still gets a valid permutation (the tree's current one)
      // and the snapshot does NOT advance: the unresolved edges are reported again by the next update, like a
      // reference update that never finished (the reference snapshots only after getGlobalPerm, ParthAPI.cpp:244-278).
      if (!spec_expanded && dperm) stage_expand(h, h.d_mesh_perm.p, h.M_n, dim, dperm);
      if (h.stats.kernels_launched == launched0) tc.cancel(); else tc.stop();
      h.diff_cached = false;  // the row diff was taken against a snapshot that stays; recompute next time
      h.spec_valid = false;
      return rc;
    }
    if (h.stats.kernels_launched == launched0) tc.cancel(); else tc.stop();
    PB_MARK("cp: expanded");
  }
  stage_snapshot(h);
  PB_MARK("cp: snapshot done");
  return rc;
}

void stage_finish_permutation(parth_b200 &h, int *dperm, int dim) {
  if (!h.xchg_pending) throw StatusError(PARTH_B200_ERR_ARG, "finish_permutation: no exchange pending");
  h.xchg_pending = false;
  stage_expand(h, h.d_mesh_perm.p, h.M_n, dim, dperm);
  stage_snapshot(h);
}

}  // namespace pb
