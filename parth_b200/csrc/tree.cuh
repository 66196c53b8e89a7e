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
// tile_dirty / tile_cols: optional per-tile verdict of the fused comparison in the pipeline build
// kernel (tile t = rows [t*tile_cols, (t+1)*tile_cols)); tile_cols == 0: compare everything
int launch_diff_rows(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev, int M,
                     unsigned *row_bits, DetectCounters *cnt, const int *tile_dirty, int tile_cols, cudaStream_t s);
int launch_diff_rows_mapped(const int *Mp, const int *Mi, const int *Mp_prev, const int *Mi_prev,
                            const int *new_to_old, const int *old_to_new, int M, unsigned *row_bits,
                            DetectCounters *cnt, cudaStream_t s);
int launch_sym_check_rows(const int *rows, const int *n_rows_dev, int cap, const int *Mp, const int *Mi,
                          const int *Mp_prev, const int *Mi_prev, int *asym, cudaStream_t s);
int launch_list_changed_rows(const unsigned *row_bits, const int *word_prefix, int n_words, int M,
                             int *rows, int cap, cudaStream_t s);
int launch_count_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                           const int *dof2node, int *counts, cudaStream_t s);
int launch_emit_row_edges(const int *rows, int n_rows, const int *n_rows_dev, const int *Mp, const int *Mi,
                          const int *dof2node, const int *offsets, int *edges, int cap, cudaStream_t s);

}  // namespace pb
