// parth_b200/csrc/hostcopy.cuh -- big transfers between the CALLER's host arrays and the device.
//
// The host-pointer entry points (the reference's own call shape: setMatrix(N, Ap, Ai), computePermutation(perm)) move
// 255 MB in and 32 MB out per update at 200^3.  From page-locked memory that is one DMA at the PCIe rate.  From a
// plain std::vector the driver stages the copy itself at ~10-12 GB/s (25 ms per update, examples/e2e_update.cpp).
// Unregistered memory therefore goes through a ring of page-locked chunks of our own that a few host threads fill
// (memcpy, pure plumbing) while the DMA of the previous chunk runs.
#pragma once
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pb {

class HostCopier {
 public:
  HostCopier() = default;
  HostCopier(const HostCopier &) = delete;
  HostCopier &operator=(const HostCopier &) = delete;
  ~HostCopier() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
    for (auto &e : ev_) if (e) cudaEventDestroy(e);
    if (ring_) cudaFreeHost(ring_);
  }

  static bool unregistered(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
  }

  // host -> device, stream-ordered on s; returns after the last chunk has been handed to the DMA engine (the source
  // has been read completely: the caller may reuse it)
  void h2d(void *ddst, const void *hsrc, size_t bytes, cudaStream_t s) {
    if (bytes < kMinStaged || !unregistered(hsrc)) {
      PB_CUDA(cudaMemcpyAsync(ddst, hsrc, bytes, cudaMemcpyHostToDevice, s));
      return;
    }
    prepare();
    // the slot sequence and the per-slot events outlive the call: the DMA of the LAST chunks of one transfer may still
    // be reading its slots when the next transfer (Ap, then Ai) starts to fill the ring
    for (size_t off = 0; off < bytes; off += kChunk, next_++) {
      const int slot = (int)(next_ % kSlots);
      const size_t len = std::min(kChunk, bytes - off);
      if (busy_[slot]) PB_CUDA(cudaEventSynchronize(ev_[slot]));  // the DMA that read this slot has finished
      parallel_copy(ring_ + (size_t)slot * kChunk, static_cast<const char *>(hsrc) + off, len);
      PB_CUDA(cudaMemcpyAsync(static_cast<char *>(ddst) + off, ring_ + (size_t)slot * kChunk, len, cudaMemcpyHostToDevice, s));
      PB_CUDA(cudaEventRecord(ev_[slot], s));
      busy_[slot] = true;
    }
  }

  // device -> host; synchronous: returns with the bytes in hdst
  void d2h(void *hdst, const void *dsrc, size_t bytes, cudaStream_t s) {
    if (bytes < kMinStaged || !unregistered(hdst)) {
      PB_CUDA(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, s));
      PB_CUDA(cudaStreamSynchronize(s));
      return;
    }
    prepare();
    for (int i = 0; i < kSlots; i++)  // uploads of an earlier call that still read the ring (possibly on another stream)
      if (busy_[i]) { PB_CUDA(cudaEventSynchronize(ev_[i])); busy_[i] = false; }
    const size_t chunks = (bytes + kChunk - 1) / kChunk;
    size_t issued = 0, drained = 0;
    while (drained < chunks) {
      while (issued < chunks && issued < drained + kSlots) {
        const int slot = (int)(issued % kSlots);
        const size_t off = issued * kChunk, len = std::min(kChunk, bytes - off);
        PB_CUDA(cudaMemcpyAsync(ring_ + (size_t)slot * kChunk, static_cast<const char *>(dsrc) + off, len, cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaEventRecord(ev_[slot], s));
        issued++;
      }
      const int slot = (int)(drained % kSlots);
      const size_t off = drained * kChunk, len = std::min(kChunk, bytes - off);
      PB_CUDA(cudaEventSynchronize(ev_[slot]));
      parallel_copy(static_cast<char *>(hdst) + off, ring_ + (size_t)slot * kChunk, len);
      drained++;
    }
  }

 private:
  static constexpr size_t kChunk = 8u << 20, kMinStaged = 4u << 20;
  static constexpr int kSlots = 4;
  struct Task { char *dst; const char *src; size_t n; };

  void prepare() {
    if (ring_) return;
    PB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ring_), kChunk * kSlots));
    ev_.assign(kSlots, nullptr);
    for (auto &e : ev_) PB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)std::max(1u, std::min(8u, hw / 2));
    for (int i = 0; i < n - 1; i++) workers_.emplace_back([this] { work(); });  // the calling thread copies a share too
  }

  void work() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !tasks_.empty(); });
        if (stop_ && tasks_.empty()) return;
        t = tasks_.back();
        tasks_.pop_back();
      }
      std::memcpy(t.dst, t.src, t.n);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) cv_done_.notify_all();
      }
    }
  }

  void parallel_copy(char *dst, const char *src, size_t n) {
    const size_t parts = workers_.size() + 1;
    const size_t piece = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
    if (workers_.empty() || n <= piece) { std::memcpy(dst, src, n); return; }
    size_t first = std::min(piece, n);
    {
      std::lock_guard<std::mutex> g(m_);
      for (size_t off = first; off < n; off += piece) {
        tasks_.push_back({dst + off, src + off, std::min(piece, n - off)});
        pending_++;
      }
    }
    cv_.notify_all();
    std::memcpy(dst, src, first);
    std::unique_lock<std::mutex> g(m_);
    cv_done_.wait(g, [this] { return pending_ == 0; });
  }

  char *ring_ = nullptr;
  std::vector<cudaEvent_t> ev_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, cv_done_;
  std::vector<Task> tasks_;
  int pending_ = 0;
  bool stop_ = false;
  size_t next_ = 0;             // ring position of the next uploaded chunk
  bool busy_[kSlots] = {};      // slot has an upload recorded on ev_[slot]
};

}  // namespace pb
