// parth_b200/csrc/hosttrace.cuh -- developer aid: host-side timeline of one update (PARTH_B200_HOSTTRACE=1).
// PB_MARK("label") stamps the host clock; parth_b200_destroy prints, per pair of consecutive labels, the median and mean time between them
// over the recorded updates.  Off (one predictable branch per mark) unless the variable is set.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace pb {

struct HostTrace {
  bool on = std::getenv("PARTH_B200_HOSTTRACE") != nullptr;
  std::vector<std::pair<const char *, long long>> ev;
  void mark(const char *label) {
    if (!on || ev.size() >= (1u << 20)) return;
    ev.emplace_back(label, (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
                               std::chrono::steady_clock::now().time_since_epoch()).count());
  }
  void dump() {
    if (!on || ev.size() < 2) return;
    // the second half of the marks only: the steady state
    std::vector<std::string> order;
    std::map<std::string, std::vector<double>> acc;
    for (size_t i = ev.size() / 2 + 1; i < ev.size(); i++) {
      const std::string key = std::string(ev[i - 1].first) + " -> " + ev[i].first;
      auto it = acc.find(key);
      if (it == acc.end()) { order.push_back(key); it = acc.emplace(key, std::vector<double>()).first; }
      it->second.push_back((double)(ev[i].second - ev[i - 1].second) / 1e3);
    }
    std::fprintf(stderr, "[hosttrace] host time between consecutive marks (us): median / mean, second half of %zu marks\n", ev.size());
    for (auto &k : order) {
      auto &v = acc[k];
      std::sort(v.begin(), v.end());
      double sum = 0;
      for (double x : v) sum += x;
      std::fprintf(stderr, "[hosttrace] %-72s %9.2f %9.2f  x%zu\n", k.c_str(), v[v.size() / 2], sum / v.size(), v.size());
    }
    ev.clear();
  }
};

inline HostTrace &host_trace() {
  static HostTrace t;
  return t;
}

}  // namespace pb

#define PB_MARK(label) ::pb::host_trace().mark(label)
