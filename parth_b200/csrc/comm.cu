// parth_b200/csrc/comm.cu -- see comm.cuh.
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the symbols are resolved with dlsym, nothing is linked

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "stages.cuh"

namespace pb {

namespace {

// ------------------------------------------------------------------ NCCL through dlopen
struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static std::string why;
  std::call_once(once, [] {
    std::vector<std::string> names;
    if (const char *e = std::getenv("PARTH_B200_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2");  // a process that has torch loaded resolves to torch's bundled copy
    names.push_back("libnccl.so");
    for (const std::string &n : names) {
      api.lib = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
      why += std::string(dlerror() ? dlerror() : "?") + "; ";
    }
    if (!api.lib) return;
#define PB_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name))
    PB_SYM(GetUniqueId); PB_SYM(CommInitRank); PB_SYM(CommDestroy); PB_SYM(AllGather); PB_SYM(AllReduce);
    PB_SYM(Broadcast); PB_SYM(GroupStart); PB_SYM(GroupEnd); PB_SYM(GetErrorString); PB_SYM(GetVersion);
#undef PB_SYM
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.AllReduce || !api.Broadcast ||
      !api.GroupStart || !api.GroupEnd)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2 failed: " + why +
                                                      "set PARTH_B200_NCCL_LIB to its path)");
  return api;
}

#define PB_NCCL(expr)                                                                                       \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess)                                                                                  \
      throw StatusError(PARTH_B200_ERR_INTERNAL, std::string(#expr) + ": " +                                \
                                                     (nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "NCCL error")); \
  } while (0)

struct NcclComm final : Comm {
  ncclComm_t comm = nullptr;
  ~NcclComm() override {
    if (comm && nccl_api().CommDestroy) nccl_api().CommDestroy(comm);
  }
  void all_gather(const int *send, int *recv, size_t n, cudaStream_t s) override {
    PB_NCCL(nccl_api().AllGather(send, recv, n, ncclInt32, comm, s));
  }
  void all_reduce_sum(int *buf, size_t n, cudaStream_t s) override {
    PB_NCCL(nccl_api().AllReduce(buf, buf, n, ncclInt32, ncclSum, comm, s));
  }
  void all_gather_v(int *buf, const size_t *offsets, const size_t *counts, cudaStream_t s) override {
    NcclApi &a = nccl_api();
    PB_NCCL(a.GroupStart());
    for (int r = 0; r < world; r++)
      if (counts[r] > 0) PB_NCCL(a.Broadcast(buf + offsets[r], buf + offsets[r], counts[r], ncclInt32, r, comm, s));
    PB_NCCL(a.GroupEnd());
  }
  const char *kind() const override { return "nccl"; }
};

// ------------------------------------------------------------------ in-process group
__global__ void sum_slices_kernel(const int *__restrict__ all, size_t n, int world, int *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = 0;
  for (int r = 0; r < world; r++) v += all[(size_t)r * n + i];
  out[i] = v;
}

}  // namespace

struct LocalGroup {
  int world = 1;
  int refs = 0;  // communicators alive; the last one frees the group
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long long generation = 0;
  std::vector<const int *> slot;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long gen = generation;
    if (++waiting == world) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else if (!cv.wait_for(lk, std::chrono::seconds(60), [&] { return generation != gen; })) {
      waiting--;  // a rank failed before it got here: give up instead of hanging the process
      throw StatusError(PARTH_B200_ERR_INTERNAL, "local communicator: a rank did not reach the collective within 60 s");
    }
  }
};

namespace {

struct LocalComm final : Comm {
  LocalGroup *g = nullptr;
  ~LocalComm() override {
    bool last;
    {
      std::lock_guard<std::mutex> lk(g->mu);
      last = --g->refs == 0;
    }
    if (last) delete g;
  }
  void publish(const int *p, cudaStream_t s) {
    PB_CUDA(cudaStreamSynchronize(s));  // what this rank contributes is complete
    g->slot[rank] = p;
    g->barrier();
  }
  void all_gather(const int *send, int *recv, size_t n, cudaStream_t s) override {
    publish(send, s);
    for (int r = 0; r < world; r++)
      if (n > 0) PB_CUDA(cudaMemcpyAsync(recv + (size_t)r * n, g->slot[r], sizeof(int) * n, cudaMemcpyDefault, s));
    PB_CUDA(cudaStreamSynchronize(s));
    g->barrier();  // nobody re-uses its send buffer before everybody has read it
  }
  void all_reduce_sum(int *buf, size_t n, cudaStream_t s) override {
    int *tmp = nullptr;
    PB_CUDA(cudaMallocAsync(&tmp, sizeof(int) * n * (size_t)world + 4, s));
    all_gather(buf, tmp, n, s);
    if (n > 0) sum_slices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tmp, n, world, buf);
    PB_CUDA(cudaFreeAsync(tmp, s));
  }
  void all_gather_v(int *buf, const size_t *offsets, const size_t *counts, cudaStream_t s) override {
    publish(buf, s);
    for (int r = 0; r < world; r++)
      if (r != rank && counts[r] > 0)
        PB_CUDA(cudaMemcpyAsync(buf + offsets[r], g->slot[r] + offsets[r], sizeof(int) * counts[r], cudaMemcpyDefault, s));
    PB_CUDA(cudaStreamSynchronize(s));
    g->barrier();
  }
  const char *kind() const override { return "local"; }
};

}  // namespace

void nccl_unique_id(void *out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  PB_NCCL(nccl_api().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

Comm *nccl_comm_create(const void *id128, int rank, int world) {
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm *c = new NcclComm();
  c->rank = rank;
  c->world = world;
  try {
    PB_NCCL(nccl_api().CommInitRank(&c->comm, world, id, rank));
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

LocalGroup *local_group_create(int world) {
  LocalGroup *g = new LocalGroup();
  g->world = world;
  g->slot.assign((size_t)world, nullptr);
  return g;
}

void local_group_destroy(LocalGroup *g) { delete g; }

Comm *local_comm_create(LocalGroup *g, int rank) {
  LocalComm *c = new LocalComm();
  c->g = g;
  c->rank = rank;
  c->world = g->world;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->refs++;
  }
  return c;
}

}  // namespace pb
