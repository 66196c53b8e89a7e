// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE (glue for the oracle/_ref build only).
//
// Exposes the UNMODIFIED reference (compiled from /root/reference/src/Parth/*.cpp by
// oracle/Makefile) behind the C ABI of cpu_parth.h, so Python tests and bench.py's
// "--impl reference" arm can drive it.  Everything Parth-specific happens inside the
// reference's own classes; this file only moves flat arrays in and out of their public
// members (ParthAPI.h:63-91, HMD.h:94-104, Integrator.h:17-28) and calls their public
// methods.  The tree injection recipe is the one verified in SURVEY.md section 7.1 step 0.
#include "cpu_parth.h"

#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include "ParthAPI.h"
#include "includes/def.h"
#include "includes/sparse_inspector.h"

struct cpu_parth {
  PARTH::ParthAPI api;
  std::vector<int> Mp_own, Mi_own;  // storage behind setMesh's borrowed pointers
  std::vector<int> map_own;
  bool quiet = true;
};

namespace {
// The reference prints "+++ PARTH: number of bad edges" even when verbose is off
// (ParthAPI.cpp:254); keep test logs readable by muting cout around calls.
struct Mute {
  std::streambuf *old = nullptr;
  std::ostringstream sink;
  explicit Mute(bool on) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
  ~Mute() { if (old) std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

const char *cpu_parth_kind(void) { return "reference"; }

cpu_parth *cpu_parth_create(void) {
  cpu_parth *h = new cpu_parth();
  h->api.setVerbose(false);
  return h;
}

void cpu_parth_destroy(cpu_parth *h) { delete h; }

void cpu_parth_set_options(cpu_parth *h, int reorder_type, int nd_levels, int lagging,
                           int aggressive, int relocate_lvl, int lag_lvl, int verbose) {
  h->api.setReorderingType((PARTH::ReorderingType)reorder_type);
  h->api.setNDLevels(nd_levels);
  h->api.lagging = lagging != 0;
  h->api.activate_aggressive_reuse = aggressive != 0;
  h->api.setRelocatedLevel(relocate_lvl);
  h->api.setLagLevel(lag_lvl);
  h->api.setVerbose(verbose != 0);
  h->quiet = verbose == 0;
}

int cpu_parth_set_matrix(cpu_parth *h, int N, const int *Ap, const int *Ai, int dim,
                         const int *map, int map_len) {
  Mute m(h->quiet);
  if (map && map_len > 0) {
    h->map_own.assign(map, map + map_len);
    h->api.setMatrix(N, const_cast<int *>(Ap), const_cast<int *>(Ai), h->map_own, dim);
  } else {
    // the two-argument overload leaves a stale map in place (ParthAPI.cpp:100-108); the
    // harness always wants "no map" here, which is what the map overload does for an empty map
    h->map_own.clear();
    h->api.setMatrix(N, const_cast<int *>(Ap), const_cast<int *>(Ai), h->map_own, dim);
  }
  return CPU_PARTH_OK;
}

int cpu_parth_set_mesh(cpu_parth *h, int n, const int *Mp, const int *Mi, const int *map,
                       int map_len) {
  Mute m(h->quiet);
  h->Mp_own.assign(Mp, Mp + n + 1);
  h->Mi_own.assign(Mi, Mi + Mp[n]);
  if (map && map_len > 0) {
    h->map_own.assign(map, map + map_len);
    h->api.setMesh(n, h->Mp_own.data(), h->Mi_own.data(), h->map_own);
  } else {
    h->api.setMesh(n, h->Mp_own.data(), h->Mi_own.data());
  }
  return CPU_PARTH_OK;
}

int cpu_parth_mesh_dims(cpu_parth *h, int *M_n, int *nnz) {
  *M_n = h->api.M_n;
  *nnz = (h->api.Mp && h->api.M_n > 0) ? h->api.Mp[h->api.M_n] : 0;
  return CPU_PARTH_OK;
}

int cpu_parth_mesh_copy(cpu_parth *h, int *Mp, int *Mi) {
  int n = h->api.M_n;
  std::memcpy(Mp, h->api.Mp, sizeof(int) * (size_t)(n + 1));
  std::memcpy(Mi, h->api.Mi, sizeof(int) * (size_t)h->api.Mp[n]);
  return CPU_PARTH_OK;
}

int cpu_parth_compute_permutation(cpu_parth *h, int dim, int *perm_out, int cap) {
  Mute m(h->quiet);
  std::vector<int> perm;
  h->api.computePermutation(perm, dim);
  if ((int)perm.size() > cap) return CPU_PARTH_ERR;
  std::memcpy(perm_out, perm.data(), sizeof(int) * perm.size());
  return (int)perm.size();
}

double cpu_parth_get_reuse(cpu_parth *h) { return h->api.getReuse(); }
int cpu_parth_get_num_changes(cpu_parth *h) { return h->api.getNumChanges(); }

int cpu_parth_get_changed_edges(cpu_parth *h, int *pairs, int cap_pairs) {
  int n = (int)h->api.changed_dof_edges.size();
  for (int i = 0; i < n && i < cap_pairs; i++) {
    pairs[2 * i] = h->api.changed_dof_edges[i].first;
    pairs[2 * i + 1] = h->api.changed_dof_edges[i].second;
  }
  return n;
}

int cpu_parth_get_dirty(cpu_parth *h, int *fine, int *nfine, int *coarse, int *ncoarse) {
  auto &f = h->api.integrator.dirty_fine_HMD_nodes_saved;
  auto &c = h->api.integrator.dirty_coarse_HMD_nodes_saved;
  *nfine = (int)f.size();
  *ncoarse = (int)c.size();
  if (fine) std::memcpy(fine, f.data(), sizeof(int) * f.size());
  if (coarse) std::memcpy(coarse, c.data(), sizeof(int) * c.size());
  return CPU_PARTH_OK;
}

int cpu_parth_get_timers(cpu_parth *h, double t[5]) {
  t[0] = h->api.init_time;
  t[1] = h->api.dof_change_integrator_time;
  t[2] = h->api.change_computation_time;
  t[3] = h->api.compute_permutation_time;
  t[4] = h->api.map_mesh_to_matrix_computation_time;
  return CPU_PARTH_OK;
}

int cpu_parth_get_mesh_perm(cpu_parth *h, int *mesh_perm, int cap) {
  int n = (int)h->api.mesh_perm.size();
  if (n > cap) return CPU_PARTH_ERR;
  std::memcpy(mesh_perm, h->api.mesh_perm.data(), sizeof(int) * (size_t)n);
  return n;
}

int cpu_parth_tree_dims(cpu_parth *h, int *M_n, int *levels, int *nodes) {
  *M_n = (int)h->api.hmd.DOF_to_HMD_node.size();
  *levels = h->api.hmd.num_levels;
  *nodes = h->api.hmd.num_HMD_nodes;
  return h->api.hmd.HMD_init ? CPU_PARTH_OK : CPU_PARTH_NO_TREE;
}

int cpu_parth_tree_export(cpu_parth *h, int *meta, int *node_ptr, int *dofs, int *labels,
                          int *dof2node, int *mesh_perm) {
  auto &hmd = h->api.hmd;
  int pos = 0;
  for (int i = 0; i < hmd.num_HMD_nodes; i++) {
    auto &nd = hmd.HMD_tree[i];
    int *m = meta + 7 * i;
    m[0] = nd.left_node_idx; m[1] = nd.right_node_idx; m[2] = nd.node_id; m[3] = nd.parent_idx;
    m[4] = nd.offset; m[5] = nd.level; m[6] = nd.org_size;
    node_ptr[i] = pos;
    for (size_t k = 0; k < nd.DOFs.size(); k++) {
      dofs[pos] = nd.DOFs[k];
      labels[pos] = nd.permuted_new_label[k];
      pos++;
    }
  }
  node_ptr[hmd.num_HMD_nodes] = pos;
  if (dof2node)
    std::memcpy(dof2node, hmd.DOF_to_HMD_node.data(), sizeof(int) * hmd.DOF_to_HMD_node.size());
  if (mesh_perm)
    std::memcpy(mesh_perm, h->api.mesh_perm.data(), sizeof(int) * h->api.mesh_perm.size());
  return CPU_PARTH_OK;
}

int cpu_parth_tree_import(cpu_parth *h, int M_n, int levels, int nodes, const int *meta,
                          const int *node_ptr, const int *dofs, const int *labels,
                          const int *Mp_prev, const int *Mi_prev) {
  auto &api = h->api;
  auto &hmd = api.hmd;
  hmd.HMD_tree.clear();
  hmd.HMD_tree.resize(nodes);
  hmd.num_levels = levels;
  hmd.num_HMD_nodes = nodes;
  hmd.DOF_to_HMD_node.assign(M_n, 0);
  api.ND_levels = levels;
  for (int i = 0; i < nodes; i++) {
    auto &nd = hmd.HMD_tree[i];
    const int *m = meta + 7 * i;
    if (m[2] == -1) continue;  // empty node keeps its default-constructed state
    nd.setIdentifier(m[0], m[1], m[2], m[3], m[4], m[5], m[6]);
    nd.setInitFlag();
    nd.DOFs.assign(dofs + node_ptr[i], dofs + node_ptr[i + 1]);
    nd.permuted_new_label.assign(labels + node_ptr[i], labels + node_ptr[i + 1]);
    nd.permuted_new_label_is_defined = true;
    for (int d : nd.DOFs) hmd.DOF_to_HMD_node[d] = i;
  }
  hmd.HMD_init = true;
  api.mesh_perm.assign(M_n, 0);
  hmd.assignCoarseLocalPermToGlobal(hmd.HMD_tree[0], api.mesh_perm);
  api.M_n_prev = M_n;
  api.Mp_prev.assign(Mp_prev, Mp_prev + M_n + 1);
  api.Mi_prev.assign(Mi_prev, Mi_prev + Mp_prev[M_n]);
  return CPU_PARTH_OK;
}

int cpu_parth_integrate(cpu_parth *h) {
  Mute m(h->quiet);
  auto &a = h->api;
  if (a.new_to_old_map.empty()) return CPU_PARTH_ERR;
  a.integrator.integrateDOFChanges(a.hmd, a.new_to_old_map, a.old_to_new_map, a.added_dofs,
                                   a.removed_dofs, a.M_n, a.Mp, a.Mi, a.M_n_prev, a.mesh_perm);
  return CPU_PARTH_OK;
}

int cpu_parth_detect(cpu_parth *h) {
  Mute m(h->quiet);
  auto &a = h->api;
  a.changed_dof_edges.clear();
  if (!a.new_to_old_map.empty())
    a.integrator.computeChangedEdges(a.hmd, a.M_n, a.Mp, a.Mi, a.M_n_prev, a.Mp_prev.data(),
                                     a.Mi_prev.data(), a.new_to_old_map.data(),
                                     a.old_to_new_map.data(), a.changed_dof_edges);
  else
    a.integrator.computeChangedEdges(a.hmd, a.M_n, a.Mp, a.Mi, a.M_n_prev, a.Mp_prev.data(),
                                     a.Mi_prev.data(), nullptr, nullptr, a.changed_dof_edges);
  return (int)a.changed_dof_edges.size();
}

int cpu_parth_permute_fine(cpu_parth *h, const int *nodes, int n) {
  Mute m(h->quiet);
  auto &a = h->api;
  std::vector<int> list(nodes, nodes + n);
  a.integrator.HMD_node_cache_flags.assign(a.hmd.num_HMD_nodes, 1);
  a.integrator.permuteFineHMDNodes(a.hmd, a.M_n, a.Mp, a.Mi, list, a.mesh_perm);
  return CPU_PARTH_OK;
}

int cpu_parth_submesh(cpu_parth *h, int node, int coarse, int *n, int *nnz, int *Mp, int *Mi,
                      int *local_to_global) {
  auto &a = h->api;
  PARTH::SubMesh sm;
  std::vector<PARTH::SubMesh> stack;
  PARTH::SubMesh *res = &sm;
  if (coarse) {
    std::vector<int> assigned;
    a.hmd.getCoarseMesh(a.hmd.HMD_tree[node], assigned, a.M_n, a.Mp, a.Mi, sm);
  } else {
    std::vector<int> flags(a.hmd.num_HMD_nodes, 0);
    flags[node] = 1;
    a.hmd.createMultiSubMesh(flags, a.hmd.DOF_to_HMD_node, a.M_n, a.Mp, a.Mi, stack);
    res = &stack[node];
  }
  *n = res->M_n;
  *nnz = res->M_nnz;
  if (Mp) std::memcpy(Mp, res->Mp.data(), sizeof(int) * res->Mp.size());
  if (Mi) std::memcpy(Mi, res->Mi.data(), sizeof(int) * res->Mi.size());
  if (local_to_global)
    std::memcpy(local_to_global, res->local_to_global_DOF_id.data(),
                sizeof(int) * res->local_to_global_DOF_id.size());
  return CPU_PARTH_OK;
}

int cpu_parth_map_mesh_perm(cpu_parth *h, const int *mesh_perm, int M_n, int dim, int *out) {
  PARTH::ParthAPI tmp;  // expansion only reads M_n and sim_dim (ParthAPI.cpp:285-318)
  tmp.setVerbose(false);
  tmp.M_n = M_n;
  tmp.sim_dim = h ? h->api.sim_dim : 1;
  std::vector<int> mp(mesh_perm, mesh_perm + M_n), res;
  int d = dim == -1 ? tmp.sim_dim : dim;
  res.resize((size_t)M_n * d);
  tmp.mapMeshPermToMatrixPerm(mp, res, dim);
  std::memcpy(out, res.data(), sizeof(int) * res.size());
  return (int)res.size();
}

int cpu_parth_amd(int n, const int *Ap, const int *Ai, int *P) {
  PARTH::HMD hmd;
  hmd.reorder_type = PARTH::ReorderingType::AMD;
  hmd.Permute(n, const_cast<int *>(Ap), const_cast<int *>(Ai), P, nullptr);  // HMD.cpp:58-87
  return CPU_PARTH_OK;
}

int cpu_parth_etree(int n, const int *Ap, const int *Ai, const int *perm, int *parent) {
  // upper-triangular CSC of P*A*P' (stype > 0), then the reference's own Liu etree
  // (src/utils/etree.cpp:39-115)
  std::vector<int> inv(n);
  for (int k = 0; k < n; k++) inv[perm[k]] = k;
  std::vector<int> cnt(n + 1, 0);
  for (int k = 0; k < n; k++) {
    int old = perm[k];
    for (int p = Ap[old]; p < Ap[old + 1]; p++)
      if (inv[Ai[p]] < k) cnt[k + 1]++;
  }
  for (int k = 0; k < n; k++) cnt[k + 1] += cnt[k];
  CSC A((size_t)n, (size_t)n, (size_t)(cnt[n] > 0 ? cnt[n] : 1), true);
  A.stype = 1;
  std::memcpy(A.p, cnt.data(), sizeof(int) * (size_t)(n + 1));
  for (int k = 0; k < n; k++) {
    int old = perm[k], q = cnt[k];
    for (int p = Ap[old]; p < Ap[old + 1]; p++)
      if (inv[Ai[p]] < k) A.i[q++] = inv[Ai[p]];
  }
  int ok = sym_lib::compute_etree(&A, parent);
  return ok ? CPU_PARTH_OK : CPU_PARTH_ERR;
}

int64_t cpu_parth_colcount(int, const int *, const int *, const int *, const int *, int *) {
  return -1;  // the reference takes column counts from CHOLMOD, which is not installed
}

}  // extern "C"
