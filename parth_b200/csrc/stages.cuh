// parth_b200/csrc/stages.cuh -- host-side stage drivers called from capi.cu.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "handle.cuh"

namespace pb {

struct StatusError : std::runtime_error {
  int code;
  StatusError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// Device-time bracket on the handle's stream.  The elapsed time lands in *dst the next time the
// host has synchronised with the stream anyway (read_ints, the host-pointer entry points) or asks
// for the statistics: the device-pointer entry points never synchronise just for a timer.
struct StageTimer {
  parth_b200 &h;
  int pair;
  double *dst, *dst2;
  bool accumulate;
  cudaEvent_t start_ev;  // != nullptr: an event already recorded on the stream marks the start (no record of our own)
  // Every cudaEventRecord costs 1.5-3 us of stream time on B200: brackets that share a boundary share the event.
  StageTimer(parth_b200 &hh, double *d, bool acc = false, double *d2 = nullptr, cudaEvent_t start = nullptr)
      : h(hh), dst(d), dst2(d2), accumulate(acc), start_ev(start) {
    pair = -1;
    if (!h.free_pairs.empty()) { pair = h.free_pairs.back(); h.free_pairs.pop_back(); }
    if (pair >= 0 && !start_ev) cudaEventRecord(h.tev[2 * pair], h.stream);
  }
  // stop_with: an event the caller records right after this call (or has just recorded) marks the end.
  // Returns the end event (nullptr when no timer pair was free).
  cudaEvent_t stop(cudaEvent_t stop_with = nullptr) {
    if (pair < 0) return nullptr;
    cudaEvent_t end = stop_with ? stop_with : h.tev[2 * pair + 1];
    if (!stop_with) cudaEventRecord(end, h.stream);
    h.pending_timers.push_back({dst, dst2, pair, accumulate, start_ev ? start_ev : h.tev[2 * pair], end});
    pair = -1;
    return end;
  }
  // nothing was enqueued inside the bracket: the span is zero, no event needed
  void cancel() {
    if (pair < 0) return;
    if (!accumulate) { *dst = 0; if (dst2) *dst2 = 0; }
    h.free_pairs.push_back(pair);
    pair = -1;
  }
  ~StageTimer() { if (pair >= 0) h.free_pairs.push_back(pair); }  // abandoned (exception): give the pair back
};
void resolve_timers(parth_b200 &h);  // synchronises the stream, then collect_timers
void collect_timers(parth_b200 &h);  // reads the stopped timers; call only right after a stream synchronisation

// small synchronous read-back of n ints through pinned memory
void read_ints(parth_b200 &h, const int *dsrc, int *dst, size_t n);

// (a1) ParthAPI::computeMeshFromMatrix
// may_defer: in the steady state (fused build + detection, edge kernels enqueued speculatively) return
// with everything in flight; stage_build_settle completes it (no-op when nothing is pending)
void stage_build(parth_b200 &h, int N, const int *dAp, const int *dAi, int nnz, int dim, bool may_defer = false);
void stage_build_settle(parth_b200 &h);

// stage_sharded.cu: one mesh over the ranks of h.comm, row blocks
void block_plan(parth_b200 &h, int M, int nnz);
void stage_build_block(parth_b200 &h, int N, int nnz, const int *dAp, const int *dAi, int q0, int q1);

// stage_perm.cu
void upload_tree_meta(parth_b200 &h, bool sync = true);
void assemble_all(parth_b200 &h);
void stage_import_tree(parth_b200 &h, int M_n, int levels, int nodes, const int *meta,
                       const int *node_ptr, const int *dofs, const int *labels, const int *Mp_prev,
                       const int *Mi_prev);
void stage_export_tree(parth_b200 &h, int *meta, int *node_ptr, int *dofs, int *labels, int *dof2node,
                       int *mesh_perm);
void stage_detect(parth_b200 &h);
void stage_dirty_sets(parth_b200 &h, const std::vector<int> &pre_fine);
void stage_expand(parth_b200 &h, const int *dmesh_perm, int n, int dim, int *dout);
void stage_snapshot(parth_b200 &h);
int stage_compute_permutation(parth_b200 &h, int *dperm, int dim);

// stage_maps.cu: (a3) DOF-map inversion, (a6) integration, (f2) relocation
void stage_set_map(parth_b200 &h, const int *map, int len, bool on_device = false);
void stage_integrate(parth_b200 &h);
void stage_relocate(parth_b200 &h, int threshold_node, std::vector<int> &dirty_fine_out, bool filter_edges = true);
void relocate_prepare(parth_b200 &h, int M);  // pre-sizes the relocation scratch (import_tree)
void check_reloc_flags(parth_b200 &h);         // deferred error flags of the regroup kernels (throws)

// stage_reorder.cu: (a12-a14) extraction + ordering of dirty fine nodes, coarse-node policy
void stage_permute_fine(parth_b200 &h, const std::vector<int> &nodes);
int stage_coarse(parth_b200 &h);
void stage_pack_labels(parth_b200 &h, const int *nodes, int n, int *dbuf, bool unpack);
void stage_finish_permutation(parth_b200 &h, int *dperm, int dim);
void stage_submesh(parth_b200 &h, int node, int coarse, int *n, int *nnz, int *Mp, int *Mi, int *l2g);
void stage_amd_batch(parth_b200 &h, int count, const int *graph_ptr, const int *Mp, const int *Mi, int *perm);

// morton.cu: ReorderingType MORTON_CODE (extension: the reference leaves it unimplemented, HMD.cpp:82-83)
void stage_set_coordinates(parth_b200 &h, const double *dxyz, int n, long long ps, long long cs, int bits);
void stage_morton_nodes(parth_b200 &h, int count, int total);  // batch in r_list / r_vtx -> r_perm

// host_decompose.cu: BFS level-structure nested dissection (host), parth_b200_decomposer signature
int builtin_decomposer(void *user, int M_n, const int *Mp, const int *Mi, int levels, int nodes, int *meta,
                       int *node_ptr, int *dofs, int *labels);

// stage_symbolic.cu: (d) etree + column counts
// on_device: perm_in / parent_out / colcount_out are device pointers (hand-off to a device solver)
void stage_symbolic(parth_b200 &h, const int *perm_in, int *parent_out, int *colcount_out, long long *lnz_out,
                    bool on_device = false);


// supernodes.cu: (f4) fundamental supernodes + supernodal etree from parent / colcount
void stage_supernodes(parth_b200 &h, int n, const int *parent_in, const int *colcount_in, int *super_out,
                      int *supermap_out, int *sparent_out, int *nsuper, long long *ssize, long long *xsize,
                      bool on_device);

}  // namespace pb
