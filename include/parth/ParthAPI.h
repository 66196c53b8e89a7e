/*
 * include/parth/ParthAPI.h -- PARTH::ParthAPI, drop-in for the reference class
 * (reference src/Parth/include/ParthAPI.h:16-145) on the dynamic-reuse ordering path.
 *
 * Same method names, argument meaning, defaults and public option / timer fields; every method is
 * a thin forward to the C ABI in include/parth_b200.h (libparth_b200.so, CUDA sm_100a).  There is
 * no CPU path: constructing the object without a usable GPU throws std::runtime_error, and any
 * non-zero status of the library is thrown too (the reference itself reports errors only through
 * assert / std::cerr, ParthAPI.cpp:191-279).
 *
 * Differences a maintainer should know (all documented in INTEGRATION.md):
 *   - `hmd` and `integrator` are lazy HOST MIRRORS of device state (tools read hmd.HMD_tree[i].DOFs,
 *     hmd.DOF_to_HMD_node, integrator.reuse_ratio, integrator.dirty_*_saved; reference
 *     tests/example_creator.cpp:66-67, src/LinSysSolver/LinSysSolver.hpp:176,216-238): call
 *     syncTreeMirror() before reading the tree vectors.
 *   - cold start: the reference dissects with METIS inside HMD::Decompose (HMD.cpp:122-346).  Here a
 *     host callback (setDecomposer) or the built-in BFS level-structure dissection builds the tree,
 *     and every node is ordered by the device AMD kernel.
 *   - dirty COARSE nodes are repaired on the device instead of re-dissected (setCoarsePolicy).
 *   - ReorderingType::MORTON_CODE works (the reference prints "Morton code is not implemented yet",
 *     HMD.cpp:82-83) once setCoordinates has supplied one point per DOF.
 */
#ifndef PARTH_B200_PARTHAPI_H
#define PARTH_B200_PARTHAPI_H

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../parth_b200.h"

namespace PARTH {

/* reference ParthTypes.h:11 */
enum ReorderingType { METIS = 0, AMD = 1, MORTON_CODE = 2 };

class ParthAPI {
public:
  //============================ OPTIONS (ParthAPI.h:19-25) =================
  ReorderingType reorder_type = ReorderingType::METIS;
  int num_cores = 10;
  bool verbose = true;
  bool activate_aggressive_reuse = false;
  bool lagging = true;
  int ND_levels = 8;
  int sim_dim = 1;

  void setReorderingType(ReorderingType type) { reorder_type = type; }
  ReorderingType getReorderingType() const { return reorder_type; }
  void setVerbose(bool v) { verbose = v; }
  bool getVerbose() const { return verbose; }
  void setNDLevels(int num) { ND_levels = num; }
  int getNDLevels() const { return ND_levels; }
  void setRelocatedLevel(int lvl) { integrator.relocate_lvl = lvl; }
  int getRelocatedLevel() const { return integrator.relocate_lvl; }
  void setLagLevel(int lvl) { integrator.lag_lvl = lvl; }
  int getLagLevel() const { return integrator.lag_lvl; }
  /* stored and never used, like the reference (ParthAPI.cpp:34) */
  void setNumberOfCores(int n) { num_cores = n; }
  int getNumberOfCores() const { return num_cores; }

  //============================ VARIABLES ================================
  /* host mirrors of the pieces of HMD / Integrator that callers read */
  struct HMDNodeMirror {
    int left_node_idx = -1, right_node_idx = -1, node_id = -1, parent_idx = -1, offset = -1, level = -1,
        org_size = -1;
    std::vector<int> DOFs, permuted_new_label;
    bool isLeaf() const { return left_node_idx == -1 && right_node_idx == -1; }
  };
  struct HMDMirror {
    bool HMD_init = false;
    int num_levels = 0, num_HMD_nodes = 0;
    std::vector<HMDNodeMirror> HMD_tree;
    std::vector<int> DOF_to_HMD_node;
  } hmd;
  struct IntegratorMirror {
    int relocate_lvl = 2, lag_lvl = 5; /* Integrator.h:26-28 */
    double reuse_ratio = 0.0;
    std::vector<int> dirty_fine_HMD_nodes_saved, dirty_coarse_HMD_nodes_saved;
  } integrator;

  std::vector<int> Mp_vec, Mi_vec; /* filled on demand by syncMeshMirror() */
  int *Mp = nullptr;
  int *Mi = nullptr;
  int M_n = 0;
  int M_n_prev = 0;
  std::vector<int> Mp_prev, Mi_prev; /* ParthAPI.h:72-73; filled on demand by syncSnapshotMirror() */
  std::vector<int> mesh_perm;
  std::vector<std::pair<int, int>> changed_dof_edges;

  /* ParthAPI.h:77-84: new_to_old_map is the caller's map as passed in; the other three are what
   * computeOldToNewDOFMap (ParthAPI.cpp:131-189) derived from it, copied back from the device by
   * setNewToOldDOFMap / computeOldToNewDOFMap */
  std::vector<int> new_to_old_map;
  std::vector<int> old_to_new_map;
  std::vector<int> added_dofs;
  std::vector<int> removed_dofs;

  double dof_change_integrator_time = 0.0;
  double change_computation_time = 0.0;
  double map_mesh_to_matrix_computation_time = 0.0;
  double compute_permutation_time = 0.0;
  double init_time = 0.0;

  //============================ FUNCTIONS =================================
  explicit ParthAPI(int device = -1) {
    int rc = parth_b200_create(&h_, device);
    if (rc != PARTH_B200_OK)
      throw std::runtime_error(std::string("PARTH (B200): ") + parth_b200_last_error(nullptr));
    parth_b200_use_builtin_decomposer(h_, 1);
  }
  ~ParthAPI() {
    if (h_) parth_b200_destroy(h_);
  }
  ParthAPI(const ParthAPI &) = delete;
  ParthAPI &operator=(const ParthAPI &) = delete;

  /* ParthAPI.cpp:62-74 */
  void clearParth() {
    check(parth_b200_clear(h_));
    mesh_perm.clear();
    changed_dof_edges.clear();
    new_to_old_map.clear();
    old_to_new_map.clear();
    added_dofs.clear();
    removed_dofs.clear();
    Mp_prev.clear();
    Mi_prev.clear();
    hmd = HMDMirror();
    M_n = M_n_prev = 0;
    Mp = Mi = nullptr;
    results_stale_ = false;
  }

  /* ParthAPI.cpp:76-98: the arrays are copied to the device during the call */
  void setMesh(int n, int *Mp_in, int *Mi_in, std::vector<int> &map) {
    setMesh(n, Mp_in, Mi_in);
    setNewToOldDOFMap(map);
  }
  void setMesh(int n, int *Mp_in, int *Mi_in) {
    pushOptions();
    check(parth_b200_set_mesh(h_, n, Mp_in, Mi_in));
    Mp = Mp_in;
    Mi = Mi_in;
    M_n = n;
    new_to_old_map.clear(); /* ParthAPI.cpp:97 */
  }

  /* ParthAPI.cpp:100-124, 366-418 */
  void setMatrix(int N, int *Ap, int *Ai, int dim = 1) {
    pushOptions();
    check(parth_b200_set_matrix(h_, N, Ap, Ai, dim));
    sim_dim = dim;
    int nnz = 0;
    parth_b200_mesh_dims(h_, &M_n, &nnz);
    Mp = Mi = nullptr; /* the graph lives on the device; syncMeshMirror() copies it back */
  }
  void setMatrix(int N, int *Ap, int *Ai, std::vector<int> &map, int dim = 1) {
    setMatrix(N, Ap, Ai, dim);
    setNewToOldDOFMap(map);
  }
  /* ParthAPI.cpp:366-418 (public in the reference): the graph is built on the device and, as in the reference,
   * left in Mp_vec / Mi_vec with Mp / Mi / M_n pointing at it; sim_dim is not touched */
  void computeMeshFromMatrix(int *Ap, int *Ai, int N, int dim) {
    const int keep = sim_dim;
    setMatrix(N, Ap, Ai, dim);
    sim_dim = keep;
    syncMeshMirror();
  }

  /* ParthAPI.cpp:126-189 */
  void setNewToOldDOFMap(std::vector<int> &map) {
    if (&map != &new_to_old_map) new_to_old_map = map;
    computeOldToNewDOFMap();
  }
  /* ParthAPI.cpp:131-189: inverts new_to_old_map on the device (cold-start rule for > 0.4 M_n added DOFs
   * included) and mirrors old_to_new_map / added_dofs / removed_dofs; set mirror_dof_maps = false to skip the copy */
  bool mirror_dof_maps = true;
  void computeOldToNewDOFMap() {
    check(parth_b200_set_new_to_old_map(h_, new_to_old_map.empty() ? nullptr : new_to_old_map.data(),
                                        (int)new_to_old_map.size()));
    map_set_ = !new_to_old_map.empty();
    old_to_new_map.clear();
    added_dofs.clear();
    removed_dofs.clear();
    if (!mirror_dof_maps || new_to_old_map.empty()) return;
    int a = 0, b = 0, c = 0;
    check(parth_b200_get_dof_maps(h_, &a, nullptr, &b, nullptr, &c, nullptr));
    old_to_new_map.resize((size_t)a);
    added_dofs.resize((size_t)b);
    removed_dofs.resize((size_t)c);
    check(parth_b200_get_dof_maps(h_, &a, old_to_new_map.data(), &b, added_dofs.data(), &c, removed_dofs.data()));
    if (a == 0) new_to_old_map.clear(); /* the map was dropped (no previous mesh / cold-start rule) */
  }

  /* ParthAPI.cpp:191-279; default dim is 3 as in the reference (ParthAPI.h:135) */
  void computePermutation(std::vector<int> &perm, int dim = 3) {
    pushOptions();
    sim_dim = dim;
    const bool warm = hmd.HMD_init && (M_n == M_n_prev || map_set_);
    map_set_ = false;
    perm.resize((size_t)M_n * dim);
    int rc = parth_b200_compute_permutation(h_, perm.data(), dim);
    if (rc != PARTH_B200_OK && rc != PARTH_B200_NEEDS_REDISSECT) check(rc);
    needs_redissect = rc == PARTH_B200_NEEDS_REDISSECT;
    double t[5];
    parth_b200_get_timers(h_, t);
    init_time = t[0];
    dof_change_integrator_time = t[1];
    change_computation_time = t[2];
    compute_permutation_time = t[3];
    map_mesh_to_matrix_computation_time = t[4];
    integrator.reuse_ratio = parth_b200_get_reuse(h_);
    const int nc = parth_b200_get_num_changes(h_);
    /* the reference prints this line on every warm update, even with verbose == false (ParthAPI.cpp:254) */
    if (warm) std::cout << "+++ PARTH: number of bad edges are " << nc << std::endl;
    int nf = 0, ncs = 0, tm = 0, lv = 0, nodes = 0;
    parth_b200_tree_dims(h_, &tm, &lv, &nodes);
    integrator.dirty_fine_HMD_nodes_saved.assign((size_t)(nodes > 0 ? nodes : 1), 0);
    integrator.dirty_coarse_HMD_nodes_saved.assign((size_t)(nodes > 0 ? nodes : 1), 0);
    parth_b200_get_dirty_nodes(h_, integrator.dirty_fine_HMD_nodes_saved.data(), &nf,
                               integrator.dirty_coarse_HMD_nodes_saved.data(), &ncs);
    integrator.dirty_fine_HMD_nodes_saved.resize((size_t)nf);
    integrator.dirty_coarse_HMD_nodes_saved.resize((size_t)ncs);
    M_n_prev = M_n;
    hmd.HMD_init = true;
    hmd.num_levels = lv;
    hmd.num_HMD_nodes = nodes;
    results_stale_ = true;
    if (!lazy_result_mirrors) syncResultMirror();
  }

  /* mesh_perm (M_n ints) and changed_dof_edges are public vectors in the reference (ParthAPI.h:75-76), so by
   * default they are copied back after every computePermutation.  Most callers read neither (the solver wrappers
   * only take `perm`): with lazy_result_mirrors = true the 4*M_n-byte device-to-host copy is skipped and
   * syncResultMirror() fetches both when somebody wants them. */
  bool lazy_result_mirrors = false;
  void syncResultMirror() {
    if (!results_stale_) return;
    const int nc = parth_b200_get_num_changes(h_);
    std::vector<int> flat((size_t)(nc > 0 ? nc : 0) * 2);
    if (nc > 0) {
      const int got = parth_b200_get_changed_edges(h_, flat.data(), nc); /* returns the count, negative = error */
      if (got < 0) check(got);
    }
    changed_dof_edges.resize((size_t)(nc > 0 ? nc : 0));
    for (int e = 0; e < nc; e++) changed_dof_edges[e] = {flat[2 * e], flat[2 * e + 1]};
    mesh_perm.resize((size_t)M_n);
    if (M_n > 0) {
      const int got = parth_b200_get_mesh_perm(h_, mesh_perm.data(), M_n); /* returns the length, negative = error */
      if (got < 0) check(got);
    }
    results_stale_ = false;
  }
  /* Mp_prev / Mi_prev (ParthAPI.h:72-73): the snapshot the next update is compared with */
  void syncSnapshotMirror() {
    int mn = 0, nnz = 0;
    check(parth_b200_get_snapshot(h_, &mn, &nnz, nullptr, nullptr));
    Mp_prev.assign((size_t)(mn > 0 ? mn + 1 : 0), 0);
    Mi_prev.resize((size_t)(nnz > 0 ? nnz : 0));
    if (mn > 0) {
      std::vector<int> tmp((size_t)(nnz > 0 ? nnz : 1));
      check(parth_b200_get_snapshot(h_, &mn, &nnz, Mp_prev.data(), tmp.data()));
      std::copy(tmp.begin(), tmp.begin() + nnz, Mi_prev.begin());
    }
  }

  /* ParthAPI.cpp:285-337; dim == -1 means sim_dim */
  void mapMeshPermToMatrixPerm(std::vector<int> &mesh_perm_in, std::vector<int> &matrix_perm, int dim = -1) {
    if (dim == -1) dim = sim_dim;
    matrix_perm.resize(mesh_perm_in.size() * (size_t)dim);
    check(parth_b200_map_mesh_perm_to_matrix_perm(h_, mesh_perm_in.data(), (int)mesh_perm_in.size(),
                                                  matrix_perm.data(), dim));
  }

  double getReuse() { return parth_b200_get_reuse(h_); }
  int getNumChanges() { return parth_b200_get_num_changes(h_); }

  /* ParthAPI.cpp:350-363 */
  void printTiming() {
    std::cout << "+++ PARTH: Timing START +++" << std::endl;
    std::cout << "+++ PARTH: Init Time: " << init_time << " sec" << std::endl;
    std::cout << "+++ PARTH: DOF Change Integration Time: " << dof_change_integrator_time << " sec" << std::endl;
    std::cout << "+++ PARTH: Change Computation Time: " << change_computation_time << " sec" << std::endl;
    std::cout << "+++ PARTH: Compute Permutation Time: " << compute_permutation_time << " sec" << std::endl;
    std::cout << "+++ PARTH: Map Mesh to Matrix Computation Time: " << map_mesh_to_matrix_computation_time << " sec"
              << std::endl;
    std::cout << "+++ PARTH: Overall reuse: " << getReuse() << std::endl;
    std::cout << "+++ PARTH: Timing END +++" << std::endl;
  }
  void resetTimers() {
    init_time = map_mesh_to_matrix_computation_time = dof_change_integrator_time = 0.0;
    change_computation_time = compute_permutation_time = 0.0;
  }

  //============================ B200 additions ============================
  bool needs_redissect = false; /* last computePermutation ended with PARTH_B200_NEEDS_REDISSECT */
  parth_b200 *handle() { return h_; }
  void setCoarsePolicy(int policy) { check(parth_b200_set_coarse_policy(h_, policy)); }
  void setAssumeSymmetric(bool on) { check(parth_b200_set_assume_symmetric(h_, on ? 1 : 0)); }
  void setDecomposer(parth_b200_decomposer fn, void *user) { check(parth_b200_set_decomposer(h_, fn, user)); }

  /* ---- ONE mesh over several GPUs (row blocks; include/parth_b200.h "ONE mesh over several GPUs").
   * One process (or thread) per GPU, each with its own ParthAPI object:
   *   rank 0: id = ParthAPI::commUniqueId(); hand the 128 bytes to every rank
   *   all   : p.commInit(id.data(), rank, world); p.importTree(tree); p.setMatrix(N, Ap, Ai, 1); p.acceptSnapshot();
   *   update: p.setMatrix(N, Ap, Ai, 1); p.computePermutation(perm, 1);   -- the collectives (NCCL) run inside
   * setMatrix then uploads only this rank's block of Ap / Ai; every rank gets the same full permutation. */
  static std::vector<unsigned char> commUniqueId() {
    std::vector<unsigned char> id(128);
    const int rc = parth_b200_comm_unique_id(id.data());
    if (rc != PARTH_B200_OK)
      throw std::runtime_error(std::string("PARTH (B200) status ") + std::to_string(rc) + ": " + parth_b200_last_error(nullptr));
    return id;
  }
  void commInit(const void *id128, int rank, int world) { check(parth_b200_comm_init(h_, id128, rank, world)); }
  /* the ranks are objects of THIS process, one host thread each (tests on one GPU) */
  static void commInitLocal(const std::vector<ParthAPI *> &ranks) {
    std::vector<parth_b200 *> hs;
    for (ParthAPI *p : ranks) hs.push_back(p->h_);
    if (parth_b200_comm_init_local(hs.data(), (int)hs.size()) != PARTH_B200_OK)
      throw std::runtime_error("PARTH (B200): comm_init_local failed");
  }
  void blockRange(int N, int nnz, int &c0, int &c1) { check(parth_b200_block_range(h_, N, nnz, &c0, &c1)); }
  /* the snapshot advances to the current graph (after the first matrix of a sharded handle, or after a REPORT) */
  void acceptSnapshot() {
    check(parth_b200_accept_snapshot(h_));
    M_n_prev = M_n;
  }
  /* a separator tree held in a host mirror (another object's `hmd` after syncTreeMirror, or one built by hand)
   * becomes this object's tree; no snapshot graph comes with it */
  void importTree(const HMDMirror &t) {
    const int nodes = (int)t.HMD_tree.size();
    std::vector<int> meta((size_t)nodes * 7), node_ptr((size_t)nodes + 1, 0), dofs, labels;
    for (int x = 0; x < nodes; x++) {
      const HMDNodeMirror &nd = t.HMD_tree[(size_t)x];
      int *m = &meta[(size_t)x * 7];
      m[0] = nd.left_node_idx; m[1] = nd.right_node_idx; m[2] = nd.node_id; m[3] = nd.parent_idx;
      m[4] = nd.offset; m[5] = nd.level; m[6] = nd.org_size;
      dofs.insert(dofs.end(), nd.DOFs.begin(), nd.DOFs.end());
      labels.insert(labels.end(), nd.permuted_new_label.begin(), nd.permuted_new_label.end());
      node_ptr[(size_t)x + 1] = (int)dofs.size();
    }
    ND_levels = t.num_levels;
    pushOptions();
    check(parth_b200_import_tree(h_, (int)dofs.size(), t.num_levels, nodes, meta.data(), node_ptr.data(), dofs.data(),
                                 labels.data(), nullptr, nullptr));
    hmd.HMD_init = true;
    hmd.num_levels = t.num_levels;
    hmd.num_HMD_nodes = nodes;
  }

  /* ReorderingType::MORTON_CODE: one 3-D point per current mesh DOF (point i, axis a at
   * xyz[i * point_stride + a * comp_stride]; an Eigen column-major V of n rows is (1, n)).  Fine nodes are then
   * ordered by ascending (Morton key, DOF id); mortonOrder returns the global order, perm[new] = old. */
  void setCoordinates(const double *xyz, int n, long long point_stride = 3, long long comp_stride = 1, int bits = 21) {
    check(parth_b200_set_coordinates(h_, xyz, n, point_stride, comp_stride, bits));
  }
  void mortonOrder(std::vector<int> &perm) {
    const int *d = nullptr;
    int n = 0;
    check(parth_b200_morton_order_dev(h_, &d, &n));
    perm.resize((size_t)n);
    if (n > 0) check(parth_b200_morton_order(h_, perm.data(), n));
  }

  /* nnz(L) of P*A*P' for the current graph and permutation (NULL: current mesh_perm); what the
   * reference ecosystem asks CHOLMOD for (src/utils/ParthTestUtils.cpp:15-59) */
  int64_t symbolic(const std::vector<int> *perm, std::vector<int> &parent, std::vector<int> &colcount) {
    parent.resize((size_t)M_n);
    colcount.resize((size_t)M_n);
    int64_t lnz = 0;
    check(parth_b200_symbolic(h_, perm ? perm->data() : nullptr, parent.data(), colcount.data(), &lnz));
    return lnz;
  }

  /* page-lock an array the caller keeps across updates (Ap, Ai, perm): the host-pointer calls then run at the PCIe
   * rate; unpin before the array is freed or reallocated.  Returns false when the driver refuses. */
  template <typename T>
  static bool pinHost(std::vector<T> &v) { return !v.empty() && parth_b200_pin_host(v.data(), v.size() * sizeof(T)) == PARTH_B200_OK; }
  template <typename T>
  static bool unpinHost(std::vector<T> &v) { return !v.empty() && parth_b200_unpin_host(v.data()) == PARTH_B200_OK; }

  /* one mesh over the ranks of a communicator (commInit*): the update of computePermutation with a DISTRIBUTED result --
   * this rank receives entries [first, first + perm_block.size()) of the expanded permutation (its row block times dim);
   * returns `first` */
  int computePermutationBlock(std::vector<int> &perm_block, int dim = 3) {
    int c0 = 0, c1 = 0;
    check(parth_b200_block_range(h_, M_n * sim_dim, 0, &c0, &c1));
    perm_block.resize((size_t)(c1 - c0) * (size_t)dim);
    int first = 0, count = 0;
    pushOptions();
    check(parth_b200_compute_permutation_block(h_, perm_block.data(), dim, &first, &count));
    perm_block.resize((size_t)count);
    return first;
  }

  /* fundamental supernodes of L from the arrays symbolic() returned (the first step of the supernodal analysis
   * behind reference src/LinSysSolver/CHOLMODSolver.cpp:108-151): super[s] = first column of supernode s
   * (super[nsuper] = n), sparent = supernodal elimination tree; returns the number of supernodes */
  int supernodes(const std::vector<int> &parent, const std::vector<int> &colcount, std::vector<int> &super,
                 std::vector<int> &supermap, std::vector<int> &sparent, int64_t *ssize = nullptr,
                 int64_t *xsize = nullptr) {
    const int n = (int)parent.size();
    int ns = 0;
    super.resize((size_t)n + 1);
    supermap.resize((size_t)n);
    sparent.resize((size_t)n);
    check(parth_b200_supernodes(h_, n, parent.data(), colcount.data(), super.data(), supermap.data(), sparent.data(),
                                &ns, ssize, xsize));
    super.resize((size_t)ns + 1);
    sparent.resize((size_t)ns);
    return ns;
  }

  /* copy the separator tree back into `hmd` (HMD.h:94-104 layout) */
  void syncTreeMirror() {
    int tm = 0, lv = 0, nodes = 0;
    if (parth_b200_tree_dims(h_, &tm, &lv, &nodes) != PARTH_B200_OK) return;
    std::vector<int> meta((size_t)nodes * 7), node_ptr((size_t)nodes + 1), dofs((size_t)tm), labels((size_t)tm);
    hmd.DOF_to_HMD_node.resize((size_t)tm);
    mesh_perm.resize((size_t)tm);
    check(parth_b200_export_tree(h_, meta.data(), node_ptr.data(), dofs.data(), labels.data(),
                                 hmd.DOF_to_HMD_node.data(), mesh_perm.data()));
    hmd.HMD_init = true;
    hmd.num_levels = lv;
    hmd.num_HMD_nodes = nodes;
    hmd.HMD_tree.assign((size_t)nodes, HMDNodeMirror());
    for (int x = 0; x < nodes; x++) {
      HMDNodeMirror &t = hmd.HMD_tree[x];
      const int *m = &meta[(size_t)x * 7];
      t.left_node_idx = m[0]; t.right_node_idx = m[1]; t.node_id = m[2]; t.parent_idx = m[3];
      t.offset = m[4]; t.level = m[5]; t.org_size = m[6];
      t.DOFs.assign(dofs.begin() + node_ptr[x], dofs.begin() + node_ptr[x + 1]);
      t.permuted_new_label.assign(labels.begin() + node_ptr[x], labels.begin() + node_ptr[x + 1]);
    }
  }
  /* Separator-tree fixture on disk (the on-disk format either side of the path; the reference keeps
   * its tree only in memory, HMD.h:94-104).  Layout, all little-endian int32: "PB2T", version 1, M_n, levels,
   * nodes, nnz, meta[nodes*7], node_ptr[nodes+1], dofs[M_n], labels[M_n], then the graph the tree belongs to --
   * Mp[M_n+1], Mi[nnz] (the snapshot the next update is compared with).  parth_b200.api.load_tree_file reads the
   * same file, so a tool written against this class and the Python tests update an identical tree. */
  bool saveTree(const std::string &path) {
    int tm = 0, lv = 0, nodes = 0, mn = 0, nnz = 0;
    if (parth_b200_tree_dims(h_, &tm, &lv, &nodes) != PARTH_B200_OK || tm <= 0) return false;
    check(parth_b200_mesh_dims(h_, &mn, &nnz));
    if (mn != tm) return false; /* the tree must describe the current graph */
    std::vector<int> meta((size_t)nodes * 7), node_ptr((size_t)nodes + 1), dofs((size_t)tm), labels((size_t)tm),
        d2n((size_t)tm), mp((size_t)tm), gp((size_t)tm + 1), gi((size_t)(nnz > 0 ? nnz : 1));
    check(parth_b200_export_tree(h_, meta.data(), node_ptr.data(), dofs.data(), labels.data(), d2n.data(), mp.data()));
    check(parth_b200_get_mesh(h_, gp.data(), gi.data()));
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const int hdr[6] = {0x54324250 /* "PB2T" */, 1, tm, lv, nodes, nnz};
    bool ok = std::fwrite(hdr, sizeof(int), 6, f) == 6;
    auto put = [&](const std::vector<int> &v, size_t n) { ok = ok && std::fwrite(v.data(), sizeof(int), n, f) == n; };
    put(meta, meta.size()); put(node_ptr, node_ptr.size()); put(dofs, dofs.size()); put(labels, labels.size());
    put(gp, gp.size()); put(gi, (size_t)nnz);
    return std::fclose(f) == 0 && ok;
  }
  bool loadTree(const std::string &path) {
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    int hdr[6] = {0, 0, 0, 0, 0, 0};
    bool ok = std::fread(hdr, sizeof(int), 6, f) == 6 && hdr[0] == 0x54324250 && hdr[1] == 1 && hdr[2] > 0 &&
              hdr[4] > 0 && hdr[5] >= 0;
    const int tm = hdr[2], lv = hdr[3], nodes = hdr[4], nnz = hdr[5];
    if (ok) {  // the header must describe exactly the bytes that follow (a corrupt file must not size the vectors)
      const long long want = 4ll * (6 + 7ll * nodes + (nodes + 1ll) + 2ll * tm + (tm + 1ll) + nnz);
      std::fseek(f, 0, SEEK_END);
      ok = std::ftell(f) == want && nodes <= 8191;
      std::fseek(f, 6 * (long)sizeof(int), SEEK_SET);
    }
    std::vector<int> meta, node_ptr, dofs, labels, gp, gi;
    auto get = [&](std::vector<int> &v, size_t n) { v.resize(n > 0 ? n : 1); ok = ok && std::fread(v.data(), sizeof(int), n, f) == n; };
    if (ok) {
      get(meta, (size_t)nodes * 7); get(node_ptr, (size_t)nodes + 1); get(dofs, (size_t)tm); get(labels, (size_t)tm);
      get(gp, (size_t)tm + 1); get(gi, (size_t)nnz);
    }
    std::fclose(f);
    if (ok) ok = node_ptr[(size_t)nodes] <= tm && gp[0] == 0 && gp[(size_t)tm] == nnz;
    if (!ok) return false;
    ND_levels = lv;
    pushOptions();
    check(parth_b200_import_tree(h_, tm, lv, nodes, meta.data(), node_ptr.data(), dofs.data(), labels.data(), gp.data(),
                                 gi.data()));
    M_n_prev = tm;
    M_n = tm;
    hmd.HMD_init = true;
    return true;
  }

  /* copy the mesh graph back (Mp_vec / Mi_vec, ParthAPI.h:63-67) and point Mp / Mi at it */
  void syncMeshMirror() {
    int nnz = 0;
    parth_b200_mesh_dims(h_, &M_n, &nnz);
    Mp_vec.resize((size_t)M_n + 1);
    Mi_vec.resize((size_t)(nnz > 0 ? nnz : 1));
    check(parth_b200_get_mesh(h_, Mp_vec.data(), Mi_vec.data()));
    Mi_vec.resize((size_t)nnz);
    Mp = Mp_vec.data();
    Mi = Mi_vec.data();
  }

private:
  parth_b200 *h_ = nullptr;
  bool map_set_ = false, results_stale_ = false;
  void pushOptions() {
    check(parth_b200_set_options(h_, (int)reorder_type, ND_levels, lagging ? 1 : 0, activate_aggressive_reuse ? 1 : 0,
                                 integrator.relocate_lvl, integrator.lag_lvl, verbose ? 1 : 0));
  }
  void check(int rc) {
    if (rc != PARTH_B200_OK)
      throw std::runtime_error(std::string("PARTH (B200) status ") + std::to_string(rc) + ": " +
                               parth_b200_last_error(h_));
  }
};

}  // namespace PARTH

#endif
