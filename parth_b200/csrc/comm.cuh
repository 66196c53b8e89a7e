// parth_b200/csrc/comm.cuh -- the collectives of the sharded (one mesh over N GPUs) path.
//
// Two implementations behind one interface:
//   * NcclComm  -- NCCL over NVLink / NVSwitch, one process per GPU.  libnccl.so.2 is loaded with dlopen at
//                  parth_b200_comm_init time (the library keeps no link-time dependency: a single-GPU caller
//                  never needs NCCL to be installed); the communicator is bootstrapped from a ncclUniqueId the
//                  caller distributes (file, MPI, torch.distributed -- anything that moves 128 bytes).
//   * LocalComm -- ranks are handles of ONE process driven by one host thread each (tests on a single GPU,
//                  or several GPUs of one process): host barriers + device-to-device copies.  TEST PLUMBING
//                  for the sharded logic; same call sequence as NCCL, so a hang or a mismatch shows up there too.
// All calls are stream-ordered on the handle's stream like NCCL's.
#pragma once
#include <cstddef>

#include <cuda_runtime.h>

namespace pb {

struct Comm {
  int rank = 0, world = 1;
  virtual ~Comm() {}
  // recv[r * n .. (r+1) * n) = rank r's send[0..n)   (ints)
  virtual void all_gather(const int *send, int *recv, size_t n, cudaStream_t s) = 0;
  // buf[i] = sum over ranks of buf[i]   (ints, in place)
  virtual void all_reduce_sum(int *buf, size_t n, cudaStream_t s) = 0;
  // every rank r contributes counts[r] ints at bufs_at[r] of its own `buf`; afterwards every rank holds all of
  // them (a group of broadcasts, one per rank with a non-empty part): variable-size all-gather in place
  virtual void all_gather_v(int *buf, const size_t *offsets, const size_t *counts, cudaStream_t s) = 0;
  virtual const char *kind() const = 0;
};

// NCCL (dlopen).  id: 128 bytes from nccl_unique_id on rank 0.
void nccl_unique_id(void *out128);
Comm *nccl_comm_create(const void *id128, int rank, int world);

// in-process group
struct LocalGroup;
LocalGroup *local_group_create(int world);
void local_group_destroy(LocalGroup *g);
Comm *local_comm_create(LocalGroup *g, int rank);

}  // namespace pb
