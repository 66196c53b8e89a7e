// tests/amd_model/amd_model.cpp -- TEST INFRASTRUCTURE.
// Runs the product's cooperative AMD source (parth_b200/csrc/amd_core.cuh) on the CPU: one OS
// thread per lane (lanes_host.h).  Built by tests/test_amd_model.py:
//   g++ -O1 -std=c++17 -shared -fPIC -pthread -I parth_b200/csrc tests/amd_model/amd_model.cpp
#include "lanes_host.h"

#include "amd_core.cuh"

#include <cstring>

extern "C" {

// orders one graph with a CTA of `nthreads` emulated threads.  extra_iw: slack of the Iw work-space
// beyond nzaat + n (0 = the legal minimum, which forces frequent workspace compaction).
// returns 1 if the rows were symmetric (fast list build), 0 for the general build.
int amd_model_order(int n, const int *Mp, const int *Mi, int *perm, int nthreads, int extra_iw) {
  using namespace pb::lanes;
  const int nz = Mp[n];
  if (nz == 0) {
    for (int i = 0; i < n; i++) perm[i] = i;
    return 1;
  }
  CtaShared cta(nthreads);
  int sh[36] = {0};
  int sym_out = 0;
  const long long slab_ints = 8LL * n + 2LL * nz + n + extra_iw;
  std::vector<int> slab((size_t)slab_ints + 16, -12345);
  auto body = [&](int tid) {
    t_tid = tid;
    t_cta = &cta;
    const bool sym = pb::amdk::rows_are_symmetric(n, Mp, 0, Mi, &sh[0]);
    const long long use = sym ? 8LL * n + nz + n + extra_iw : slab_ints;
    pb::amdk::order_graph(n, Mp, 0, Mi, sym, slab.data(), use, perm, sh);
    if (tid == 0) sym_out = sym ? 1 : 0;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(body, t);
  for (auto &t : th) t.join();
  return sym_out;
}

int amd_model_warp() { return pb::lanes::WARP; }
}
