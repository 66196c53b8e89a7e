// parth_b200/csrc/capi.cu -- extern "C" boundary (include/parth_b200.h) and host orchestration.
#include <cstring>
#include <iostream>

#include "handle.cuh"
#include "stages.cuh"

using namespace pb;

#define PB_API_BEGIN  \
  if (!h) return PARTH_B200_ERR_ARG; \
  try {                 \
    PB_CUDA(cudaSetDevice(h->device));
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
    PB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreate(&h->ev0));
    PB_CUDA(cudaEventCreate(&h->ev1));
    PB_CUDA(cudaEventCreate(&h->ev2));
    PB_CUDA(cudaEventCreate(&h->ev3));
    h->pin.reserve(1 << 16);
    h->bstatus.reserve(1);
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
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev2) cudaEventDestroy(h->ev2);
  if (h->ev3) cudaEventDestroy(h->ev3);
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
  h->tree_init = false;
  h->num_levels = h->num_nodes = h->tree_M = 0;
  h->meta.clear();
  h->node_ptr.clear();
  h->n_changed = 0;
  h->dirty_fine.clear(); h->dirty_coarse.clear();
  h->dirty_fine_saved.clear(); h->dirty_coarse_saved.clear();
  h->reuse = 1;
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
  if (!h || policy < 0 || policy > 2) return PARTH_B200_ERR_ARG;
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

void *parth_b200_stream(parth_b200 *h) { return h ? (void *)h->stream : nullptr; }

int parth_b200_synchronize(parth_b200 *h) {
  PB_API_BEGIN
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_get_stats(parth_b200 *h, parth_b200_stats *out) {
  if (!h || !out) return PARTH_B200_ERR_ARG;
  *out = h->stats;
  return PARTH_B200_OK;
}

// ------------------------------------------------------------------ matrix / mesh input
int parth_b200_set_matrix_dev(parth_b200 *h, int N, const int *dAp, const int *dAi, int nnz, int dim) {
  PB_API_BEGIN
  if (!dAp || !dAi || N <= 0 || dim <= 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: bad argument");
  h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
  h->sim_dim = dim;
  stage_build(*h, N, dAp, dAi, nnz, dim);
  PB_API_END
}

int parth_b200_set_matrix(parth_b200 *h, int N, const int *Ap, const int *Ai, int dim) {
  PB_API_BEGIN
  if (!Ap || !Ai || N <= 0 || dim <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_matrix: bad argument");
  const int nnz = Ap[N];
  if (Ap[0] != 0 || nnz < 0) throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: Ap[0] != 0 or Ap[N] < 0");
  h->dAp.reserve((size_t)N + 1, h->stream);
  h->dAi.reserve((size_t)(nnz > 0 ? nnz : 1), h->stream);
  PB_CUDA(cudaMemcpyAsync(h->dAp.p, Ap, sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, h->stream));
  if (nnz > 0)
    PB_CUDA(cudaMemcpyAsync(h->dAi.p, Ai, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  h->stats.h2d_bytes = (long long)sizeof(int) * ((long long)N + 1 + nnz);
  h->stats.d2h_bytes = 0;
  h->sim_dim = dim;
  stage_build(*h, N, h->dAp.p, h->dAi.p, nnz, dim);
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
  PB_API_END
}

int parth_b200_set_mesh(parth_b200 *h, int n, const int *Mp, const int *Mi) {
  PB_API_BEGIN
  if (!Mp || !Mi || n <= 0) throw StatusError(PARTH_B200_ERR_ARG, "set_mesh: bad argument");
  const int nnz = Mp[n];
  h->Mp_own.reserve((size_t)n + 1, h->stream);
  h->Mi_own.reserve((size_t)(nnz > 0 ? nnz : 1), h->stream);
  PB_CUDA(cudaMemcpyAsync(h->Mp_own.p, Mp, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, h->stream));
  if (nnz > 0)
    PB_CUDA(cudaMemcpyAsync(h->Mi_own.p, Mi, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  PB_CUDA(cudaStreamSynchronize(h->stream));  // the caller may reuse its arrays
  h->stats.h2d_bytes = (long long)sizeof(int) * ((long long)n + 1 + nnz);
  h->Mp = h->Mp_own.p;
  h->Mi = h->Mi_own.p;
  h->M_n = n;
  h->nnzM = nnz;
  h->mesh_owned = true;
  h->map_len = 0;
  PB_API_END
}

int parth_b200_mesh_dims(parth_b200 *h, int *M_n, int *nnz) {
  if (!h || !M_n || !nnz) return PARTH_B200_ERR_ARG;
  *M_n = h->M_n;
  *nnz = h->nnzM;
  return PARTH_B200_OK;
}

int parth_b200_get_mesh(parth_b200 *h, int *Mp, int *Mi) {
  PB_API_BEGIN
  if (!h->Mp || !Mp || !Mi) throw StatusError(PARTH_B200_ERR_ARG, "get_mesh: no mesh");
  PB_CUDA(cudaMemcpyAsync(Mp, h->Mp, sizeof(int) * ((size_t)h->M_n + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h->nnzM > 0)
    PB_CUDA(cudaMemcpyAsync(Mi, h->Mi, sizeof(int) * (size_t)h->nnzM, cudaMemcpyDeviceToHost, h->stream));
  PB_CUDA(cudaStreamSynchronize(h->stream));
  PB_API_END
}

int parth_b200_mesh_dev(parth_b200 *h, const int **dMp, const int **dMi) {
  if (!h || !dMp || !dMi) return PARTH_B200_ERR_ARG;
  *dMp = h->Mp;
  *dMi = h->Mi;
  return PARTH_B200_OK;
}

}  // extern "C"
