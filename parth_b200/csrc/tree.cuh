// parth_b200/csrc/tree.cuh -- launchers of tree.cu / detect.cu.
#pragma once
#include "common.cuh"

namespace pb {

int launch_assemble_all(const int *node_ptr, const int *meta, const int *visit, int nodes,
                        const int *dofs, const int *labels, int total, int *mesh_perm, cudaStream_t s);
int launch_assemble_nodes(const int *list, int n_list, const int *node_ptr, const int *meta,
                          const int *dofs, const int *labels, int *mesh_perm, cudaStream_t s);
int launch_dof2node(const int *node_ptr, int nodes, const int *dofs, int total, int *dof2node,
                    cudaStream_t s);
int launch_expand(const int *mesh_perm, int M_n, int dim, int *out, cudaStream_t s);
int launch_mark_dirty_nodes(const int *edges, int n_edges, const int *dof2node, const int *meta,
                            int *within_flag, int *lca_flag, cudaStream_t s);
// same with the edge count read on the device (n_edges_dev, clamped to cap); also copies the count to *count_out
int launch_mark_dirty_nodes_dev(const int *edges, const int *n_edges_dev, int cap, const int *dof2node,
                                const int *meta, int *within_flag, int *lca_flag, int *count_out, cudaStream_t s);

// detect.cu
struct DetectCounters {
  int changed_rows;   // rows whose adjacency differs from the snapshot
  int dirty_tiles;    // row tiles that needed the per-row comparison
  int pad0;           // incremental symmetry proof: a mirror entry is missing
  int pad1;           // fused detection: tiles the build kernel left to diff_rows_kernel
};
int launch_diff_rows(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev, int M,
                     unsigned *row_bits, DetectCounters *cnt, cudaStream_t s);
int launch_diff_rows_mapped(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev,
                            const int *new_to_old, const int *old_to_new, int M, unsigned *row_bits,
                            DetectCounters *cnt, cudaStream_t s);
// counts != nullptr: also the problematic-edge count of the first spec_cap changed rows (count_row_edges fused in)
int launch_sym_check_rows(const int *rows, const int *n_rows_dev, int cap, const int *Mp, const int *Mi,
                          const int *Mp_prev, const int *Mi_prev, int *asym, const int *dof2node, int *counts,
                          int spec_cap, cudaStream_t s, int *zero_out = nullptr, int n_zero = 0);
// fused steady state (the pipeline build kernel wrote the bitmap words and per-tile counts of changed rows):
// ascending changed-row list from the tile counts (tile_off: num_tiles + 1 ints of scratch, scan_ws: scan work-space)
int launch_rows_from_tiles(const int *tile_cnt, int *tile_off, unsigned long long *scan_ws, int num_tiles, int tile_cols,
                           const unsigned *row_bits, int M, int *rows, int cap, cudaStream_t s);
// ordered edge emission (row offsets summed per CTA from the per-row counts) + dirty-node marking; out = w5 layout
// ([0] edge count or -1 = not computed, [2, 2+nn) within-node flags, [2+nn, 2+2nn) LCA flags, cleared by the
// launch_sym_check_rows in front of it through zero_out / n_zero)
constexpr int kSpecRows = 8192;
int launch_emit_mark(const int *rows, const DetectCounters *cnt, const int *counts, const int *Mp, const int *Mi,
                     const int *dof2node, const int *meta, int *edges, int cap_edges, int nn, int *out, cudaStream_t s);
int launch_list_changed_rows(const unsigned *row_bits, const int *word_prefix, int n_words, int M,
                             int *rows, int cap, cudaStream_t s);
int launch_count_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                           const int *dof2node, int *counts, cudaStream_t s);
int launch_emit_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                          const int *dof2node, const int *offsets, int *edges, int cap, cudaStream_t s);

#ifdef __CUDACC__
// HMD::isSeparator (HMD.cpp:366-379)
__device__ __forceinline__ bool is_separator_dev(int sep, int node) {
  int a = (node - 1) / 2;
  while (a > sep && a > 0) a = (a - 1) / 2;
  return a == sep;
}

// HMD::findCommonSeparator (HMD.cpp:381-419) on the stored level / parent fields
__device__ __forceinline__ int common_separator_dev(const int *__restrict__ meta, int first, int second) {
  int hi = first, lo = second;
  if (__ldg(meta + hi * 7 + 5) < __ldg(meta + lo * 7 + 5)) { hi = second; lo = first; }
  while (__ldg(meta + hi * 7 + 5) != __ldg(meta + lo * 7 + 5)) hi = __ldg(meta + hi * 7 + 3);
  while (hi != lo) { hi = __ldg(meta + hi * 7 + 3); lo = __ldg(meta + lo * 7 + 3); }
  return hi;
}
#endif

}  // namespace pb
