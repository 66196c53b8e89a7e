// parth_b200/csrc/stages.cuh -- host-side stage drivers called from capi.cu.
#pragma once
#include <stdexcept>
#include <string>

#include "handle.cuh"

namespace pb {

struct StatusError : std::runtime_error {
  int code;
  StatusError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// RAII device-time bracket on the handle's stream (events ev0/ev1); read with ms() after a sync.
struct StageTimer {
  parth_b200 &h;
  cudaEvent_t a, b;
  StageTimer(parth_b200 &hh, cudaEvent_t ea, cudaEvent_t eb) : h(hh), a(ea), b(eb) {
    cudaEventRecord(a, h.stream);
  }
  void stop() { cudaEventRecord(b, h.stream); }
  double ms() {
    float t = 0;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&t, a, b);
    return t;
  }
};

// small synchronous read-back of n ints through pinned memory
void read_ints(parth_b200 &h, const int *dsrc, int *dst, size_t n);

// (a1) ParthAPI::computeMeshFromMatrix
void stage_build(parth_b200 &h, int N, const int *dAp, const int *dAi, int nnz, int dim);

}  // namespace pb
