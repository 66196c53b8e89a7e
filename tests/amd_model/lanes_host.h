// tests/amd_model/lanes_host.h -- TEST INFRASTRUCTURE.
// Host implementation of the lane vocabulary of parth_b200/csrc/lanes.cuh: every lane of a CTA is
// an OS thread, warp collectives are barrier + exchange slots.  The warp is PB_MODEL_WARP lanes wide
// (8 by default: lists longer than a chunk, chunk boundaries and cross-chunk ordering are exercised
// far more often than with 32).  Used only by tests/test_amd_model.py to run the product's
// cooperative AMD source (parth_b200/csrc/amd_core.cuh) against the oracle on the CPU.
#pragma once
#include <atomic>
#include <climits>
#include <thread>
#include <vector>

#ifndef PB_MODEL_WARP
#define PB_MODEL_WARP 8
#endif

#define PB_DEV static inline
#define PB_DEVFN static

namespace pb {
namespace lanes {

constexpr int WARP = PB_MODEL_WARP;
constexpr unsigned FULL = WARP == 32 ? 0xffffffffu : ((1u << WARP) - 1u);

struct Barrier {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  int n = 1;
  void wait() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_acq_rel);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 64) std::this_thread::yield();
    }
  }
};

struct WarpShared {
  Barrier bar;
  int slot[32];
};

struct CtaShared {
  Barrier bar;
  int nthreads = 0;
  std::vector<WarpShared> warps;
  explicit CtaShared(int nt) : nthreads(nt), warps((nt + WARP - 1) / WARP) {
    bar.n = nt;
    for (auto &w : warps) w.bar.n = WARP;
  }
};

inline thread_local int t_tid = 0;
inline thread_local CtaShared *t_cta = nullptr;

PB_DEV int lane_id() { return t_tid % WARP; }
PB_DEV int thread_id() { return t_tid; }
PB_DEV int num_threads() { return t_cta->nthreads; }
PB_DEV WarpShared &my_warp() { return t_cta->warps[t_tid / WARP]; }
PB_DEV void sync_warp() { my_warp().bar.wait(); }
PB_DEV void sync_cta() { t_cta->bar.wait(); }

PB_DEV int shfl(int v, int src) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  const int r = w.slot[src];
  w.bar.wait();
  return r;
}
PB_DEV unsigned ballot(bool p) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = p ? 1 : 0;
  w.bar.wait();
  unsigned m = 0;
  for (int l = 0; l < WARP; l++) if (w.slot[l]) m |= 1u << l;
  w.bar.wait();
  return m;
}
PB_DEV unsigned match_any(int v) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  unsigned m = 0;
  for (int l = 0; l < WARP; l++) if (w.slot[l] == v) m |= 1u << l;
  w.bar.wait();
  return m;
}
PB_DEV int warp_sum(int v) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  int s = 0;
  for (int l = 0; l < WARP; l++) s += w.slot[l];
  w.bar.wait();
  return s;
}
PB_DEV int warp_min(int v) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  int s = INT_MAX;
  for (int l = 0; l < WARP; l++) s = w.slot[l] < s ? w.slot[l] : s;
  w.bar.wait();
  return s;
}
PB_DEV int warp_max(int v) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  int s = INT_MIN;
  for (int l = 0; l < WARP; l++) s = w.slot[l] > s ? w.slot[l] : s;
  w.bar.wait();
  return s;
}
PB_DEV int warp_incl_scan(int v) {
  WarpShared &w = my_warp();
  w.slot[lane_id()] = v;
  w.bar.wait();
  int s = 0;
  for (int l = 0; l <= lane_id(); l++) s += w.slot[l];
  w.bar.wait();
  return s;
}
PB_DEV unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }
PB_DEV unsigned lanemask_gt() { return FULL & ~((2u << lane_id()) - 1u); }
PB_DEV int popc(unsigned m) { return __builtin_popcount(m); }
PB_DEV int lowest(unsigned m) { return __builtin_ctz(m); }
PB_DEV int highest(unsigned m) { return 31 - __builtin_clz(m); }
PB_DEV void atomic_add(int *p, int v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
PB_DEV int atomic_add_ret(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

}  // namespace lanes
}  // namespace pb
