// parth_b200/csrc/build.cuh -- interface of graph_build.cu (stage a).
#pragma once
#include "common.cuh"

namespace pb {

enum : unsigned { BUILD_UNSORTED = 1u, BUILD_ASYM = 2u, BUILD_RANGE = 4u, BUILD_WIDE = 8u, BUILD_DUP = 16u, BUILD_NODIAG = 32u };

struct BuildStatus {
  unsigned flags;       // BUILD_* bits raised by the filter path
  unsigned diff_mod;    // (kept upper - kept lower) modulo 2^31; 0 <=> counts match
  int nnz_out;          // entries written to Mi
  int pad;
  unsigned long long diff_sum;  // column-aligned kernel: sum of (upper - lower) over all columns, two's complement
};

struct BuildArgs {
  const int *Ap, *Ai;
  const int *V;   // dim > 1: exclusive prefix of samples per block-column (M+1); dim == 1: unused
  int dim, M, Q;  // Q = number of samples (nnzA for dim == 1)
  int *tile_c;    // num_tiles + 1
  int *Mp, *Mi;
  BuildStatus *status;
  unsigned long long *desc;
  int check_mirror;
  cudaEvent_t ev_a, ev_b;  // optional: bracket the filter kernel alone (roofline timing)
  // optional (pipeline kernel only): fused change detection against the snapshot graph.  Every
  // tile is compared with the same rows of the snapshot while both are in shared memory and the
  // changed-row bitmap (row_bits, one bit per row) + counters are written by the build kernel;
  // tile_dirty[t] = 2 marks the tiles left to diff_rows_kernel.
  const int *Mp_prev, *Mi_prev;
  int nnz_prev;
  int *tile_dirty;
  unsigned *row_bits;
  struct DetectCounters *dcnt;
  int pipe_cfg;  // pipeline kernel: tile shape (graph_build_pipe.cu)
  int mail_prezeroed;  // pipeline kernel, fused: status + counters were zeroed behind the previous update's mailbox copy
  // pipeline kernel, row-block mode (stage_sharded.cu): blk_TC > 0 fixes the tile shape of the whole matrix
  // (M_rows columns); this launch owns tiles [blk_t0, blk_t1), M / Q / nnz_prev are the ends of the block's slices
  // and every pointer is biased so that global ids index it
  int blk_TC, blk_t0, blk_t1, M_rows;
};

int build_filter_tiles(int M, int Q);
// graph_build_cols.cu (dim == 1): spec = closed-form offsets (every column holds its diagonal)
int build_cols_tiles(int Q);
int build_cols_launch(const BuildArgs &a, bool spec, cudaStream_t s);
// full symmetry proof on the compacted output graph (used for dim > 1 instead of in-kernel probes)
int sym_proof_csr_launch(const int *Mp, const int *Mi, int M, int nnz_cap, BuildStatus *st, cudaStream_t s);
// graph_build_pipe.cu (dim == 1, aligned arrays): persistent bulk-copy pipeline, closed-form offsets
int build_pipe_supported(const BuildArgs &a);
int build_pipe_default_cfg();
int build_pipe_launch(const BuildArgs &a, int num_sms, cudaStream_t s);
int build_pipe_tile_cols(const BuildArgs &a);  // columns per tile of the pipeline kernel
int build_pipe_num_tiles(const BuildArgs &a);
int build_filter_launch(const BuildArgs &a, cudaStream_t s);
int build_sample_prefix_launch(const int *Ap, int M, int dim, int *ns_tmp, int *V,
                               unsigned long long *scan_ws, cudaStream_t s);
int build_general_count_launch(const BuildArgs &a, int *cnt, int *ptr, unsigned long long *scan_ws,
                               cudaStream_t s);
int build_general_finish_launch(const BuildArgs &a, int *cnt, const int *ptr, int *tmp, int *uniq,
                                unsigned long long *scan_ws, cudaStream_t s);
int build_general_gather_launch(const BuildArgs &a, const int *ptr, const int *tmp, cudaStream_t s);

}  // namespace pb
