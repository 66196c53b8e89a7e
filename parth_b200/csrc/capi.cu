// parth_b200/csrc/capi.cu -- extern "C" boundary (include/parth_b200.h) and host orchestration.
#include <algorithm>
#include <cstring>
#include <iostream>

#include "handle.cuh"
#include "stages.cuh"

using namespace pb;

// PB_API_BEGIN settles a build that parth_b200_set_matrix_dev left in flight (stage_build_settle);
// PB_API_BEGIN_NOSETTLE is for the entry points that handle it themselves
#define PB_API_BEGIN_NOSETTLE \
  if (!h) return PARTH_B200_ERR_ARG; \
  try {                 \
    PB_CUDA(cudaSetDevice(h->device));
#define PB_API_BEGIN_STREAM  \
  PB_API_BEGIN_NOSETTLE \
  stage_build_settle(*h);
// ... and reads the error flags a stream-ordered update left behind (one small read-back when there are any);
// PB_API_BEGIN_STREAM is for the stream-ordered entry points that must not wait for the previous update
#define PB_API_BEGIN  \
  PB_API_BEGIN_STREAM \
  check_reloc_flags(*h);
#define PB_API_END                                   \
  }                                                  \
  catch (const pb::CudaError &e) {                   \
    h->err = e.what();                               \
    return e.code;                                   \
  }                                                  \
  catch (const pb::StatusError &e) {                 \
    h->err = e.what();                               \
    return e.code;                                   \
  }                                                  \
  catch (const std::exception &e) {                  \
    h->err = e.what();                               \
    return PARTH_B200_ERR_INTERNAL;                  \
  }                                                  \
  return PARTH_B200_OK;

static thread_local std::string g_create_error;

// entry points outside the PB_API_* brackets that read what a build produces settle it first
static int settle_quiet(parth_b200 *h) {
  if (!h->pend.active) return PARTH_B200_OK;
  PB_API_BEGIN
  PB_API_END
}


// Row-block mode, host arrays in GLOBAL indexing: uploads Ap[c0..c1] and Ai[Ap[c0]..Ap[c1]) -- 1/world of the matrix --
// and builds this rank's rows.  The staging buffers are placed so that the biased pointers the kernels index with
// global ids keep the 16-byte alignment of the bulk copies (graph_build_pipe.cu).
static void set_matrix_block_host(parth_b200 &h, int N, int nnz, const int *Ap, const int *Ai, int dim) {
  if (dim != 1) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "row-block mode: dim must be 1");
  if (nnz < 0) throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: nnz < 0");
  block_plan(h, N, nnz);
  const int c0 = h.blk_c0, c1 = h.blk_c1;
  const int q0 = Ap[c0], q1 = Ap[c1];
  if (q0 < 0 || q1 < q0 || q1 > nnz) throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: Ap is not monotone over the block");
  cudaStream_t s = h.stream;
  const int off = q0 & 3;
  h.dAp.reserve((size_t)(c1 - c0) + 8, s);
  h.dAi.reserve((size_t)(q1 - q0) + 16, s);
  h.hostcopy.h2d(h.dAp.p, Ap + c0, sizeof(int) * ((size_t)(c1 - c0) + 1), s);
  if (q1 > q0) h.hostcopy.h2d(h.dAi.p + off, Ai + q0, sizeof(int) * (size_t)(q1 - q0), s);
  h.stats.h2d_bytes = (long long)sizeof(int) * ((long long)(c1 - c0) + 1 + (q1 - q0));
  h.stats.d2h_bytes = 0;
  stage_build_block(h, N, nnz, h.dAp.p - c0, h.dAi.p + off - q0, q0, q1);
  resolve_timers(h);
}

extern "C" {

const char *parth_b200_version(void) { return "parth_b200 0.1 (sm_100a)"; }

const char *parth_b200_last_error(const parth_b200 *h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

int parth_b200_create(parth_b200 **out, int device) {
  if (!out) return PARTH_B200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no usable CUDA device (this library has no CPU path): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return PARTH_B200_ERR_NO_DEVICE;
  }
  parth_b200 *h = new parth_b200();
  try {
    if (device < 0) PB_CUDA(cudaGetDevice(&device));
    h->device = device;
    PB_CUDA(cudaSetDevice(device));
    PB_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    PB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreate(&h->ev0));
    PB_CUDA(cudaEventCreate(&h->ev1));
    PB_CUDA(cudaEventCreate(&h->ev2));
    PB_CUDA(cudaEventCreate(&h->ev3));
    PB_CUDA(cudaEventCreate(&h->ev_mail));
    PB_CUDA(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->pin.reserve(1 << 16);
    {
      // keep freed work-space in the device's default pool instead of returning it to the driver
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess && pool) {
        unsigned long long keep_all = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
      }
      (void)cudaGetLastError();
    }
    for (auto &e : h->tev) PB_CUDA(cudaEventCreate(&e));
    for (int i = parth_b200::kTimerPairs - 1; i >= 0; i--) h->free_pairs.push_back(i);
    // [0..5] BuildStatus, [8..11] DetectCounters, [16..) stage (b)'s speculative result of the fused
    // steady state (edge count + node flags, <= 2 + 2 * 8191 ints): one read-back fetches all of it
    h->mail.reserve(16 + 2 + 2 * 8191 + 14);
  } catch (const std::exception &ex) {
    g_create_error = ex.what();
    delete h;
    return PARTH_B200_ERR_NO_DEVICE;
  }
  *out = h;
  return PARTH_B200_OK;
}

int parth_b200_destroy(parth_b200 *h) {
  if (!h) return PARTH_B200_ERR_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  host_trace().dump();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev2) cudaEventDestroy(h->ev2);
  if (h->ev3) cudaEventDestroy(h->ev3);
  if (h->ev_mail) cudaEventDestroy(h->ev_mail);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (auto &e : h->tev) if (e) cudaEventDestroy(e);
  delete h->comm;
  h->comm = nullptr;
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return PARTH_B200_OK;
}

int parth_b200_clear(parth_b200 *h) {
  PB_API_BEGIN
  PB_CUDA(cudaStreamSynchronize(h->stream));
  h->Mp = h->Mi = nullptr;
  h->M_n = h->nnzM = 0;
  h->M_n_prev = h->nnz_prev = 0;
  h->map_len = 0;
  h->morton_n = 0;
  h->tree_init = false;
  h->cur_sym_proven = h->prev_sym_proven = false;
  h->diff_cached = false;
  h->num_levels = h->num_nodes = h->tree_M = 0;
  h->meta.clear();
  h->node_ptr.clear();
  h->n_changed = 0;
  h->dirty_fine.clear(); h->dirty_coarse.clear();
  h->dirty_fine_saved.clear(); h->dirty_coarse_saved.clear();
  h->reuse = 1;
  // row-block mode: the partition is chosen again by the next matrix (the communicator stays)
  h->block_mode = false;
  h->blk_TC = h->blk_tiles = h->blk_t0 = h->blk_t1 = h->blk_c0 = h->blk_c1 = 0;
  h->blk_tiles_zeroed = false;
  h->pipe_cfg = pb::build_pipe_default_cfg();
  PB_API_END
}

int parth_b200_set_options(parth_b200 *h, int reorder_type, int nd_levels, int lagging,
                           int aggressive_reuse, int relocate_lvl, int lag_lvl, int verbose) {
  if (!h) return PARTH_B200_ERR_ARG;
  h->reorder_type = reorder_type;
  h->nd_levels = nd_levels;
  h->lagging = lagging != 0;
  h->aggressive = aggressive_reuse != 0;
  h->relocate_lvl = relocate_lvl;
  h->lag_lvl = lag_lvl;
  h->verbose = verbose != 0;
  return PARTH_B200_OK;
}

int parth_b200_set_coarse_policy(parth_b200 *h, int policy) {
  if (!h || policy < 0 || policy > 3) return PARTH_B200_ERR_ARG;
  h->coarse_policy = policy;
  return PARTH_B200_OK;
}

int parth_b200_set_assume_symmetric(parth_b200 *h, int on) {
  if (!h) return PARTH_B200_ERR_ARG;
  h->assume_symmetric = on != 0;
  return PARTH_B200_OK;
}

int parth_b200_set_shard(parth_b200 *h, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return PARTH_B200_ERR_ARG;
  h->rank = rank;
  h->world = world;
  return PARTH_B200_OK;
}

int parth_b200_set_decomposer(parth_b200 *h, parth_b200_decomposer fn, void *user) {
  if (!h) return PARTH_B200_ERR_ARG;
  h->decomposer = fn;
  h->decomposer_user = user;
  return PARTH_B200_OK;
}

int parth_b200_use_builtin_decomposer(parth_b200 *h, int on) {
  if (!h) return PARTH_B200_ERR_ARG;
  h->builtin_decomposer = on != 0;
  return PARTH_B200_OK;
}

void *parth_b200_stream(parth_b200 *h) { return h ? (void *)h->stream : nullptr; }

int parth_b200_pin_host(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return PARTH_B200_ERR_ARG;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); g_create_error = cudaGetErrorString(e); return (int)e; }
  return PARTH_B200_OK;
}

int parth_b200_unpin_host(void *ptr) {
  if (!ptr) return PARTH_B200_ERR_ARG;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); g_create_error = cudaGetErrorString(e); return (int)e; }
  return PARTH_B200_OK;
}

int parth_b200_synchronize(parth_b200 *h) {
  PB_API_BEGIN
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_get_stats(parth_b200 *h, parth_b200_stats *out) {
  if (!h || !out) return PARTH_B200_ERR_ARG;
  if (!h->pending_timers.empty() || h->pend.active) {  // device-pointer calls leave their timers to be read here
    PB_API_BEGIN
    resolve_timers(*h);
    *out = h->stats;
    PB_API_END
  }
  *out = h->stats;
  return PARTH_B200_OK;
}

// ------------------------------------------------------------------ matrix / mesh input
int parth_b200_set_matrix_dev(parth_b200 *h, int N, const int *dAp, const int *dAi, int nnz, int dim) {
  PB_MARK("set_matrix_dev: called");
  PB_API_BEGIN_STREAM
  if (!dAp || !dAi || N <= 0 || dim <= 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: bad argument");
  if (N % dim != 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: N is not a multiple of dim (ParthAPI.cpp:371 asserts it)");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  h->sim_dim = dim;
  if (h->comm) {
    // one mesh over the ranks of the communicator: only this rank's block of the arrays is read
    if (dim != 1) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "row-block mode: dim must be 1");
    block_plan(*h, N, nnz);
    int q[2] = {0, 0};
    read_ints(*h, dAp + h->blk_c0, &q[0], 1);
    read_ints(*h, dAp + h->blk_c1, &q[1], 1);
    stage_build_block(*h, N, nnz, dAp, dAi, q[0], q[1]);
    return PARTH_B200_OK;
  }
  // steady state: returns with the build, the changed-row proof and stage (b)'s edge kernels in flight;
  // whatever call needs their outcome next waits for the mailbox copy alone (stage_build_settle).
  // Otherwise it ends on its mailbox read-back and nothing is in flight.
  PB_MARK("set_matrix_dev: entry");
  stage_build(*h, N, dAp, dAi, nnz, dim, /*may_defer=*/true);
  PB_MARK("set_matrix_dev: build enqueued");
  PB_API_END
}

int parth_b200_set_matrix(parth_b200 *h, int N, const int *Ap, const int *Ai, int dim) {
  PB_API_BEGIN
  if (!Ap || !Ai || N <= 0 || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: bad argument");
  if (N % dim != 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: N is not a multiple of dim (ParthAPI.cpp:371 asserts it)");
  const int nnz = Ap[N];
  if (Ap[0] != 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: Ap[0] != 0 or Ap[N] < 0");
  if (h->comm) {
    h->sim_dim = dim;
    set_matrix_block_host(*h, N, nnz, Ap, Ai, dim);
    return PARTH_B200_OK;
  }
  h->dAp.reserve((size_t)N + 1, h->stream);
  h->dAi.reserve((size_t)(nnz > 0 ? nnz : 1), h->stream);
  h->hostcopy.h2d(h->dAp.p, Ap, sizeof(int) * ((size_t)N + 1), h->stream);
  if (nnz > 0) h->hostcopy.h2d(h->dAi.p, Ai, sizeof(int) * (size_t)nnz, h->stream);
  h->stats.h2d_bytes = (long long)sizeof(int) * ((long long)N + 1 + nnz);
  h->stats.d2h_bytes = 0;
  h->sim_dim = dim;
  stage_build(*h, N, h->dAp.p, h->dAi.p, nnz, dim);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_set_mesh_dev(parth_b200 *h, int n, const int *dMp, const int *dMi, int nnz) {
  PB_API_BEGIN
  if (!dMp || !dMi || n <= 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_ARG, "set_mesh: bad argument");
  h->Mp = dMp;
  h->Mi = dMi;
  h->M_n = n;
  h->nnzM = nnz;
  h->mesh_owned = false;
  h->map_len = 0;
  h->cur_sym_proven = false;
  h->diff_cached = false;
  PB_API_END
}

int parth_b200_set_mesh(parth_b200 *h, int n, const int *Mp, const int *Mi) {
  PB_API_BEGIN
  if (!Mp || !Mi || n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_mesh: bad argument");
  const int nnz = Mp[n];
  if (Mp[0] != 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_ARG, "set_mesh: Mp[0] != 0 or Mp[n] < 0");
  h->Mp_own.reserve((size_t)n + 1, h->stream);
  h->Mi_own.reserve((size_t)(nnz > 0 ? nnz : 1), h->stream);
  h->hostcopy.h2d(h->Mp_own.p, Mp, sizeof(int) * ((size_t)n + 1), h->stream);
  if (nnz > 0) h->hostcopy.h2d(h->Mi_own.p, Mi, sizeof(int) * (size_t)nnz, h->stream);
  PB_CUDA(cudaStreamSynchronize(h->stream));  // the caller may reuse its arrays
  h->stats.h2d_bytes = (long long)sizeof(int) * ((long long)n + 1 + nnz);
  h->Mp = h->Mp_own.p;
  h->Mi = h->Mi_own.p;
  h->M_n = n;
  h->nnzM = nnz;
  h->mesh_owned = true;
  h->map_len = 0;
  h->cur_sym_proven = false;
  h->diff_cached = false;
  PB_API_END
}

int parth_b200_mesh_dims(parth_b200 *h, int *M_n, int *nnz) {
  if (!h || !M_n || !nnz) return PARTH_B200_ERR_ARG;
  if (int rc = settle_quiet(h)) return rc;
  *M_n = h->M_n;
  *nnz = h->nnzM;
  return PARTH_B200_OK;
}

int parth_b200_get_mesh(parth_b200 *h, int *Mp, int *Mi) {
  PB_API_BEGIN
  if (!h->Mp || !Mp || !Mi) throw StatusError(PARTH_B200_ERR_ARG, "get_mesh: no mesh");
  if (h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "get_mesh: the handle holds one row block of a sharded mesh");
  PB_CUDA(cudaMemcpyAsync(Mp, h->Mp, sizeof(int) * ((size_t)h->M_n + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h->nnzM > 0)
    PB_CUDA(cudaMemcpyAsync(Mi, h->Mi, sizeof(int) * (size_t)h->nnzM, cudaMemcpyDeviceToHost, h->stream));
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_mesh_dev(parth_b200 *h, const int **dMp, const int **dMi) {
  if (!h || !dMp || !dMi) return PARTH_B200_ERR_ARG;
  if (int rc = settle_quiet(h)) return rc;
  *dMp = h->Mp;
  *dMi = h->Mi;
  return PARTH_B200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ tree, permutation, results
extern "C" {

int parth_b200_set_new_to_old_map(parth_b200 *h, const int *map, int len) {
  PB_API_BEGIN
  if (len < 0 || (len > 0 && !map)) throw StatusError(PARTH_B200_ERR_ARG, "set_new_to_old_map: bad argument");
  if (len > 0 && h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "set_new_to_old_map: DOF maps are not sharded (row-block mode)");
  stage_set_map(*h, map, len);
  PB_API_END
}

int parth_b200_set_new_to_old_map_dev(parth_b200 *h, const int *dmap, int len) {
  PB_API_BEGIN
  if (len < 0 || (len > 0 && !dmap)) throw StatusError(PARTH_B200_ERR_ARG, "set_new_to_old_map: bad argument");
  if (len > 0 && h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "set_new_to_old_map: DOF maps are not sharded (row-block mode)");
  stage_set_map(*h, dmap, len, /*on_device=*/true);
  PB_API_END
}

int parth_b200_import_tree(parth_b200 *h, int M_n, int levels, int nodes, const int *meta,
                           const int *node_ptr, const int *dofs, const int *labels,
                           const int *Mp_prev, const int *Mi_prev) {
  PB_API_BEGIN
  if (!meta || !node_ptr || !dofs || !labels) throw StatusError(PARTH_B200_ERR_ARG, "import_tree: null array");
  stage_import_tree(*h, M_n, levels, nodes, meta, node_ptr, dofs, labels, Mp_prev, Mi_prev);
  PB_API_END
}

int parth_b200_tree_dims(parth_b200 *h, int *M_n, int *levels, int *nodes) {
  if (!h || !M_n || !levels || !nodes) return PARTH_B200_ERR_ARG;
  *M_n = h->tree_M;
  *levels = h->num_levels;
  *nodes = h->num_nodes;
  return h->tree_init ? PARTH_B200_OK : PARTH_B200_ERR_NO_TREE;
}

int parth_b200_export_tree(parth_b200 *h, int *meta, int *node_ptr, int *dofs, int *labels,
                           int *dof2node, int *mesh_perm) {
  PB_API_BEGIN
  if (!meta || !node_ptr) throw StatusError(PARTH_B200_ERR_ARG, "export_tree: null array");
  stage_export_tree(*h, meta, node_ptr, dofs, labels, dof2node, mesh_perm);
  PB_API_END
}

int parth_b200_compute_permutation_dev(parth_b200 *h, int *dperm, int dim) {
  int rc = PARTH_B200_OK;
  PB_MARK("compute_permutation_dev: called");
  PB_API_BEGIN_NOSETTLE
  if (!dperm || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "computePermutation: bad argument");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  // stream-ordered: dperm is valid for work enqueued on the handle's stream after this call (or
  // after parth_b200_synchronize); the stage timers are read at the next synchronisation point
  h->defer_reloc_check = true;  // stream-ordered: the regroup kernels' error flags are read by the next synchronising call
  PB_MARK("compute_permutation_dev: entry");
  try {
    rc = stage_compute_permutation(*h, dperm, dim);
  } catch (...) {
    h->defer_reloc_check = false;
    throw;
  }
  PB_MARK("compute_permutation_dev: done");
  h->defer_reloc_check = false;
  if (rc != PARTH_B200_OK) return rc;
  PB_API_END
}

int parth_b200_compute_permutation(parth_b200 *h, int *perm, int dim) {
  int rc = PARTH_B200_OK;
  PB_API_BEGIN
  if (!perm || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "computePermutation: bad argument");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  const size_t n_out = (size_t)h->M_n * dim;
  h->d_out.reserve(n_out > 0 ? n_out : 1, h->stream);
  rc = stage_compute_permutation(*h, h->d_out.p, dim);
  if (rc == PARTH_B200_OK || rc == PARTH_B200_NEEDS_REDISSECT) {  // REPORT: the tree's current permutation
    h->hostcopy.d2h(perm, h->d_out.p, sizeof(int) * n_out, h->stream);
    h->stats.d2h_bytes += (long long)(sizeof(int) * n_out);
  }
  resolve_timers(*h);
  if (rc != PARTH_B200_OK) return rc;
  PB_API_END
}

// one mesh over the ranks of a communicator: the whole update runs as in compute_permutation, but this rank's host
// receives only ITS block of the expanded permutation, entries [c0 * dim, c1 * dim) with [c0, c1) its row block
int parth_b200_compute_permutation_block(parth_b200 *h, int *perm_block, int dim, int *first, int *count) {
  int rc = PARTH_B200_OK;
  PB_API_BEGIN
  if (!perm_block || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "computePermutationBlock: bad argument");
  if (!h->block_mode) throw StatusError(PARTH_B200_ERR_ARG, "computePermutationBlock: the handle holds no row block (set_matrix_block*)");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  const size_t n_out = (size_t)h->M_n * dim;
  h->d_out.reserve(n_out > 0 ? n_out : 1, h->stream);
  rc = stage_compute_permutation(*h, h->d_out.p, dim);
  const size_t b0 = (size_t)h->blk_c0 * dim, nb = (size_t)(h->blk_c1 - h->blk_c0) * dim;
  if (first) *first = (int)b0;
  if (count) *count = (int)nb;
  if ((rc == PARTH_B200_OK || rc == PARTH_B200_NEEDS_REDISSECT) && nb > 0) {
    h->hostcopy.d2h(perm_block, h->d_out.p + b0, sizeof(int) * nb, h->stream);
    h->stats.d2h_bytes += (long long)(sizeof(int) * nb);
  }
  resolve_timers(*h);
  if (rc != PARTH_B200_OK) return rc;
  PB_API_END
}

int parth_b200_get_snapshot(parth_b200 *h, int *M_n_prev, int *nnz_prev, int *Mp_prev, int *Mi_prev) {
  PB_API_BEGIN
  if (!M_n_prev || !nnz_prev) throw StatusError(PARTH_B200_ERR_ARG, "get_snapshot: null size pointer");
  if (h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "get_snapshot: the handle holds one row block of a sharded mesh");
  *M_n_prev = h->M_n_prev;
  *nnz_prev = h->M_n_prev > 0 ? h->nnz_prev : 0;
  if (h->M_n_prev > 0 && Mp_prev)
    PB_CUDA(cudaMemcpyAsync(Mp_prev, h->Mp_prev.p, sizeof(int) * ((size_t)h->M_n_prev + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h->M_n_prev > 0 && Mi_prev && h->nnz_prev > 0)
    PB_CUDA(cudaMemcpyAsync(Mi_prev, h->Mi_prev.p, sizeof(int) * (size_t)h->nnz_prev, cudaMemcpyDeviceToHost, h->stream));
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_get_dof_maps(parth_b200 *h, int *len_old, int *old_to_new, int *n_added, int *added, int *n_removed,
                            int *removed) {
  PB_API_BEGIN
  if (!len_old || !n_added || !n_removed) throw StatusError(PARTH_B200_ERR_ARG, "get_dof_maps: null size pointer");
  const bool have = h->map_len > 0;
  *len_old = have ? h->M_n_prev : 0;
  *n_added = have ? h->n_added : 0;
  *n_removed = have ? h->n_removed : 0;
  if (have && old_to_new)
    PB_CUDA(cudaMemcpyAsync(old_to_new, h->old_to_new.p, sizeof(int) * (size_t)h->M_n_prev, cudaMemcpyDeviceToHost, h->stream));
  if (have && added && h->n_added > 0)
    PB_CUDA(cudaMemcpyAsync(added, h->added.p, sizeof(int) * (size_t)h->n_added, cudaMemcpyDeviceToHost, h->stream));
  if (have && removed && h->n_removed > 0)
    PB_CUDA(cudaMemcpyAsync(removed, h->removed.p, sizeof(int) * (size_t)h->n_removed, cudaMemcpyDeviceToHost, h->stream));
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_accept_snapshot(parth_b200 *h) {
  PB_API_BEGIN
  if (!h->Mp || h->M_n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "accept_snapshot: no mesh set");
  stage_snapshot(*h);
  PB_API_END
}

int parth_b200_exchange_plan(parth_b200 *h, int *n_nodes, int *nodes, int *owner, int *sizes) {
  if (!h || !n_nodes) return PARTH_B200_ERR_ARG;
  *n_nodes = (int)h->xchg_nodes.size();
  for (size_t i = 0; i < h->xchg_nodes.size(); i++) {
    if (nodes) nodes[i] = h->xchg_nodes[i];
    if (owner) owner[i] = h->xchg_owner[i];
    if (sizes) sizes[i] = h->node_ptr[h->xchg_nodes[i] + 1] - h->node_ptr[h->xchg_nodes[i]];
  }
  return PARTH_B200_OK;
}

int parth_b200_pack_labels_dev(parth_b200 *h, const int *nodes, int n, int *dbuf) {
  PB_API_BEGIN
  if (!h->tree_init || n < 0 || (n > 0 && (!nodes || !dbuf))) throw StatusError(PARTH_B200_ERR_ARG, "pack_labels: bad argument");
  stage_pack_labels(*h, nodes, n, dbuf, false);
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_unpack_labels_dev(parth_b200 *h, const int *nodes, int n, const int *dbuf) {
  PB_API_BEGIN
  if (!h->tree_init || n < 0 || (n > 0 && (!nodes || !dbuf))) throw StatusError(PARTH_B200_ERR_ARG, "unpack_labels: bad argument");
  stage_pack_labels(*h, nodes, n, const_cast<int *>(dbuf), true);
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_finish_permutation_dev(parth_b200 *h, int *dperm, int dim) {
  PB_API_BEGIN
  if (!dperm || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "finish_permutation: bad argument");
  stage_finish_permutation(*h, dperm, dim);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_finish_permutation(parth_b200 *h, int *perm, int dim) {
  PB_API_BEGIN
  if (!perm || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "finish_permutation: bad argument");
  const size_t n_out = (size_t)h->M_n * dim;
  h->d_out.reserve(n_out > 0 ? n_out : 1, h->stream);
  stage_finish_permutation(*h, h->d_out.p, dim);
  PB_CUDA(cudaMemcpyAsync(perm, h->d_out.p, sizeof(int) * n_out, cudaMemcpyDeviceToHost, h->stream));
  h->stats.d2h_bytes += (long long)(sizeof(int) * n_out);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_map_mesh_perm_to_matrix_perm_dev(parth_b200 *h, const int *dmesh_perm, int n,
                                                int *dout, int dim) {
  PB_API_BEGIN
  if (!dmesh_perm || !dout || n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "mapMeshPermToMatrixPerm: bad argument");
  if (dim == -1) dim = h->sim_dim;
  if (dim < 1) throw StatusError(PARTH_B200_ERR_ARG, "mapMeshPermToMatrixPerm: dim must be >= 1 (or -1 for sim_dim)");
  stage_expand(*h, dmesh_perm, n, dim, dout);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_map_mesh_perm_to_matrix_perm(parth_b200 *h, const int *mesh_perm, int n, int *out,
                                            int dim) {
  PB_API_BEGIN
  if (!mesh_perm || !out || n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "mapMeshPermToMatrixPerm: bad argument");
  if (dim == -1) dim = h->sim_dim;
  if (dim < 1) throw StatusError(PARTH_B200_ERR_ARG, "mapMeshPermToMatrixPerm: dim must be >= 1 (or -1 for sim_dim)");
  h->w0.reserve((size_t)n, h->stream);
  h->d_out.reserve((size_t)n * dim, h->stream);
  PB_CUDA(cudaMemcpyAsync(h->w0.p, mesh_perm, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  stage_expand(*h, h->w0.p, n, dim, h->d_out.p);
  PB_CUDA(cudaMemcpyAsync(out, h->d_out.p, sizeof(int) * (size_t)n * dim, cudaMemcpyDeviceToHost, h->stream));
  resolve_timers(*h);
  h->stats.h2d_bytes = (long long)sizeof(int) * n;
  h->stats.d2h_bytes = (long long)sizeof(int) * n * dim;
  PB_API_END
}

double parth_b200_get_reuse(parth_b200 *h) { return h ? h->reuse : 0.0; }
int parth_b200_get_num_changes(parth_b200 *h) { return h ? h->n_changed : 0; }

int parth_b200_get_changed_edges(parth_b200 *h, int *pairs, int cap_pairs) {
  if (!h) return PARTH_B200_ERR_ARG;
  try {
    PB_CUDA(cudaSetDevice(h->device));
    const int n = h->n_changed < cap_pairs ? h->n_changed : cap_pairs;
    if (n > 0 && pairs) {
      PB_CUDA(cudaMemcpyAsync(pairs, h->d_changed.p, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
      PB_CUDA(cudaStreamSynchronize(h->stream));
    }
  } catch (const std::exception &e) {
    h->err = e.what();
    return PARTH_B200_ERR_INTERNAL;
  }
  return h->n_changed;
}

int parth_b200_get_dirty_nodes(parth_b200 *h, int *fine, int *nfine, int *coarse, int *ncoarse) {
  if (!h || !nfine || !ncoarse) return PARTH_B200_ERR_ARG;
  *nfine = (int)h->dirty_fine_saved.size();
  *ncoarse = (int)h->dirty_coarse_saved.size();
  if (fine) std::memcpy(fine, h->dirty_fine_saved.data(), sizeof(int) * h->dirty_fine_saved.size());
  if (coarse) std::memcpy(coarse, h->dirty_coarse_saved.data(), sizeof(int) * h->dirty_coarse_saved.size());
  return PARTH_B200_OK;
}

int parth_b200_get_timers(parth_b200 *h, double t[5]) {
  if (!h || !t) return PARTH_B200_ERR_ARG;
  if (!h->pending_timers.empty() || h->pend.active) {
    PB_API_BEGIN
    resolve_timers(*h);
    for (int i = 0; i < 5; i++) t[i] = h->timers[i];
    PB_API_END
  }
  for (int i = 0; i < 5; i++) t[i] = h->timers[i];
  return PARTH_B200_OK;
}

int parth_b200_get_mesh_perm(parth_b200 *h, int *mesh_perm, int cap) {
  if (!h || !mesh_perm) return PARTH_B200_ERR_ARG;
  if (!h->tree_init) return 0;
  if (h->tree_M > cap) return PARTH_B200_ERR_ARG;
  try {
    PB_CUDA(cudaSetDevice(h->device));
    PB_CUDA(cudaMemcpyAsync(mesh_perm, h->d_mesh_perm.p, sizeof(int) * (size_t)h->tree_M, cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
  } catch (const std::exception &e) {
    h->err = e.what();
    return PARTH_B200_ERR_INTERNAL;
  }
  return h->tree_M;
}

int parth_b200_mesh_perm_dev(parth_b200 *h, const int **dmesh_perm) {
  if (!h || !dmesh_perm) return PARTH_B200_ERR_ARG;
  *dmesh_perm = h->tree_init ? h->d_mesh_perm.p : nullptr;
  return PARTH_B200_OK;
}

int parth_b200_stage_integrate(parth_b200 *h) {
  PB_API_BEGIN
  if (!h->tree_init || h->map_len == 0) throw StatusError(PARTH_B200_ERR_ARG, "stage_integrate: needs a tree and a map");
  stage_integrate(*h);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_stage_detect(parth_b200 *h, int *nchanged) {
  PB_API_BEGIN
  if (!h->tree_init || !h->Mp) throw StatusError(PARTH_B200_ERR_ARG, "stage_detect: needs a tree and a mesh");
  stage_detect(*h);
  resolve_timers(*h);
  if (nchanged) *nchanged = h->n_changed;
  PB_API_END
}

int parth_b200_stage_permute_fine(parth_b200 *h, const int *nodes, int n) {
  PB_API_BEGIN
  if (!h->tree_init || !h->Mp || (n > 0 && !nodes)) throw StatusError(PARTH_B200_ERR_ARG, "stage_permute_fine: bad argument");
  stage_permute_fine(*h, std::vector<int>(nodes, nodes + n));
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_stage_submesh(parth_b200 *h, int node, int coarse, int *n, int *nnz, int *Mp, int *Mi,
                             int *local_to_global) {
  PB_API_BEGIN
  if (!h->tree_init || !h->Mp || !n || !nnz || node < 0 || node >= h->num_nodes)
    throw StatusError(PARTH_B200_ERR_ARG, "stage_submesh: bad argument");
  if (h->block_mode && coarse) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "stage_submesh(coarse): needs the whole graph (row-block mode)");
  stage_submesh(*h, node, coarse, n, nnz, Mp, Mi, local_to_global);
  PB_API_END
}

int parth_b200_amd_batch(parth_b200 *h, int count, const int *graph_ptr, const int *Mp, const int *Mi,
                         int *perm) {
  PB_API_BEGIN
  if (count < 0 || !graph_ptr || !Mp || !Mi || !perm) throw StatusError(PARTH_B200_ERR_ARG, "amd_batch: bad argument");
  stage_amd_batch(*h, count, graph_ptr, Mp, Mi, perm);
  PB_API_END
}

// ------------------------------------------------------------------ ReorderingType MORTON_CODE (extension)
int parth_b200_set_coordinates_dev(parth_b200 *h, const double *dxyz, int n, long long point_stride,
                                   long long comp_stride, int bits) {
  PB_API_BEGIN
  if (n < 0 || (n > 0 && !dxyz) || point_stride < 0 || comp_stride < 0)
    throw StatusError(PARTH_B200_ERR_ARG, "set_coordinates: bad argument");
  stage_set_coordinates(*h, dxyz, n, point_stride, comp_stride, bits);
  PB_API_END
}

int parth_b200_set_coordinates(parth_b200 *h, const double *xyz, int n, long long point_stride,
                               long long comp_stride, int bits) {
  PB_API_BEGIN
  if (n < 0 || (n > 0 && !xyz) || point_stride < 0 || comp_stride < 0)
    throw StatusError(PARTH_B200_ERR_ARG, "set_coordinates: bad argument");
  // extent of the strided array in doubles
  const size_t span = n > 0 ? (size_t)(n - 1) * (size_t)point_stride + 2 * (size_t)comp_stride + 1 : 0;
  h->m_xyz.reserve(std::max<size_t>(span, 1), h->stream);
  if (span) PB_CUDA(cudaMemcpyAsync(h->m_xyz.p, xyz, sizeof(double) * span, cudaMemcpyHostToDevice, h->stream));
  h->stats.h2d_bytes = (long long)(sizeof(double) * span);
  stage_set_coordinates(*h, h->m_xyz.p, n, point_stride, comp_stride, bits);
  resolve_timers(*h);
  PB_API_END
}

int parth_b200_morton_order_dev(parth_b200 *h, const int **dperm, int *n) {
  if (!h || !dperm || !n) return PARTH_B200_ERR_ARG;
  *dperm = h->morton_n > 0 ? (h->morton_cur ? h->m_v1.p : h->m_v0.p) : nullptr;
  *n = h->morton_n;
  return PARTH_B200_OK;
}

int parth_b200_morton_order(parth_b200 *h, int *perm, int cap) {
  PB_API_BEGIN
  if (!perm || cap < h->morton_n) throw StatusError(PARTH_B200_ERR_ARG, "morton_order: buffer too small");
  if (h->morton_n > 0) {
    PB_CUDA(cudaMemcpyAsync(perm, h->morton_cur ? h->m_v1.p : h->m_v0.p, sizeof(int) * (size_t)h->morton_n,
                            cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
  }
  PB_API_END
}

int parth_b200_morton_keys(parth_b200 *h, unsigned long long *keys_sorted, int cap) {
  PB_API_BEGIN
  if (!keys_sorted || cap < h->morton_n) throw StatusError(PARTH_B200_ERR_ARG, "morton_keys: buffer too small");
  if (h->morton_n > 0) {
    PB_CUDA(cudaMemcpyAsync(keys_sorted, h->morton_cur ? h->m_k1.p : h->m_k0.p,
                            sizeof(unsigned long long) * (size_t)h->morton_n, cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
  }
  PB_API_END
}

int parth_b200_symbolic(parth_b200 *h, const int *perm, int *parent, int *colcount, int64_t *lnz) {
  PB_API_BEGIN
  if (!h->Mp || !lnz) throw StatusError(PARTH_B200_ERR_ARG, "symbolic: needs a mesh");
  if (h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "symbolic: the handle holds one row block of a sharded mesh");
  long long v = 0;
  stage_symbolic(*h, perm, parent, colcount, &v);
  resolve_timers(*h);
  *lnz = v;
  PB_API_END
}

// device-pointer variant: dperm (may be NULL: current mesh_perm), dparent and dcolcount (each may be
// NULL) live on the device; lnz is returned on the host
int parth_b200_symbolic_dev(parth_b200 *h, const int *dperm, int *dparent, int *dcolcount, int64_t *lnz) {
  PB_API_BEGIN
  if (!h->Mp || !lnz) throw StatusError(PARTH_B200_ERR_ARG, "symbolic: needs a mesh");
  if (h->block_mode) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "symbolic: the handle holds one row block of a sharded mesh");
  long long v = 0;
  stage_symbolic(*h, dperm, dparent, dcolcount, &v, /*on_device=*/true);
  resolve_timers(*h);
  *lnz = v;
  PB_API_END
}

int parth_b200_supernodes(parth_b200 *h, int n, const int *parent, const int *colcount, int *super, int *supermap,
                          int *sparent, int *nsuper, int64_t *ssize, int64_t *xsize) {
  PB_API_BEGIN
  if (n < 0 || (n > 0 && (!parent || !colcount))) throw StatusError(PARTH_B200_ERR_ARG, "supernodes: parent / colcount missing");
  long long a = 0, b = 0;
  int ns = 0;
  stage_supernodes(*h, n, parent, colcount, super, supermap, sparent, &ns, &a, &b, /*on_device=*/false);
  resolve_timers(*h);
  if (nsuper) *nsuper = ns;
  if (ssize) *ssize = a;
  if (xsize) *xsize = b;
  PB_API_END
}

int parth_b200_supernodes_dev(parth_b200 *h, int n, const int *dparent, const int *dcolcount, int *dsuper,
                              int *dsupermap, int *dsparent, int *nsuper, int64_t *ssize, int64_t *xsize) {
  PB_API_BEGIN
  if (n < 0 || (n > 0 && (!dparent || !dcolcount))) throw StatusError(PARTH_B200_ERR_ARG, "supernodes: parent / colcount missing");
  long long a = 0, b = 0;
  int ns = 0;
  stage_supernodes(*h, n, dparent, dcolcount, dsuper, dsupermap, dsparent, &ns, &a, &b, /*on_device=*/true);
  resolve_timers(*h);
  if (nsuper) *nsuper = ns;
  if (ssize) *ssize = a;
  if (xsize) *xsize = b;
  PB_API_END
}

// ------------------------------------------------------------------ one mesh over several GPUs (row blocks)
int parth_b200_comm_unique_id(void *id128) {
  if (!id128) return PARTH_B200_ERR_ARG;
  try {
    nccl_unique_id(id128);
  } catch (const pb::StatusError &e) {
    g_create_error = e.what();
    return e.code;
  } catch (const std::exception &e) {
    g_create_error = e.what();
    return PARTH_B200_ERR_INTERNAL;
  }
  return PARTH_B200_OK;
}

int parth_b200_comm_init(parth_b200 *h, const void *id128, int rank, int world) {
  PB_API_BEGIN
  if (!id128 || world < 1 || rank < 0 || rank >= world) throw StatusError(PARTH_B200_ERR_ARG, "comm_init: bad argument");
  if (h->comm) throw StatusError(PARTH_B200_ERR_ARG, "comm_init: the handle already has a communicator");
  h->comm = nccl_comm_create(id128, rank, world);
  PB_API_END
}

int parth_b200_comm_init_local(parth_b200 **handles, int world) {
  if (!handles || world < 1) return PARTH_B200_ERR_ARG;
  for (int r = 0; r < world; r++)
    if (!handles[r] || handles[r]->comm) return PARTH_B200_ERR_ARG;
  LocalGroup *g = local_group_create(world);
  for (int r = 0; r < world; r++) handles[r]->comm = local_comm_create(g, r);
  return PARTH_B200_OK;
}

int parth_b200_comm_info(parth_b200 *h, int *rank, int *world, const char **kind) {
  if (!h) return PARTH_B200_ERR_ARG;
  if (rank) *rank = h->comm ? h->comm->rank : 0;
  if (world) *world = h->comm ? h->comm->world : 1;
  if (kind) *kind = h->comm ? h->comm->kind() : "none";
  return PARTH_B200_OK;
}

int parth_b200_block_range(parth_b200 *h, int N, int nnz, int *c0, int *c1) {
  PB_API_BEGIN
  if (N <= 0 || nnz < 0 || !c0 || !c1) throw StatusError(PARTH_B200_ERR_ARG, "block_range: bad argument");
  block_plan(*h, N, nnz);
  *c0 = h->blk_c0;
  *c1 = h->blk_c1;
  PB_API_END
}

int parth_b200_set_matrix_block(parth_b200 *h, int N, int nnz, const int *Ap, const int *Ai, int dim) {
  PB_API_BEGIN
  if (!Ap || !Ai || N <= 0 || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix_block: bad argument");
  if (!h->comm) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix_block: no communicator (parth_b200_comm_init*)");
  h->sim_dim = dim;
  set_matrix_block_host(*h, N, nnz, Ap, Ai, dim);
  PB_API_END
}

int parth_b200_set_matrix_block_dev(parth_b200 *h, int N, int nnz, const int *dAp, const int *dAi, int dim, int q0,
                                    int q1) {
  PB_API_BEGIN_STREAM
  if (!dAp || !dAi || N <= 0 || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix_block: bad argument");
  if (!h->comm) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix_block: no communicator (parth_b200_comm_init*)");
  if (dim != 1) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "row-block mode: dim must be 1");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  h->sim_dim = dim;
  stage_build_block(*h, N, nnz, dAp, dAi, q0, q1);
  PB_API_END
}

}  // extern "C"
