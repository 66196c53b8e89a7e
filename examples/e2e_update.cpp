// examples/e2e_update.cpp -- what a C++ caller of the drop-in class sees end to end: warm reuse updates through
// PARTH::ParthAPI with host arrays (std::vector, as the reference's callers hold them), pageable versus page-locked.
//
//   e2e_update [n = 100] [updates = 10]
//
// 7-point Poisson grid n^3 (dim = 1), cold start through the built-in dissection, then `updates` warm updates that
// alternate between the base matrix and one with a few extra couplings inside one leaf (a local change: the update
// re-orders one node).  Prints the wall time of setMatrix + computePermutation per update for both kinds of memory.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <parth/parth.h>

static void grid(int n, bool extra, int &N, std::vector<int> &Ap, std::vector<int> &Ai) {
  N = n * n * n;
  Ap.assign(1, 0);
  Ai.clear();
  Ai.reserve((size_t)7 * N + 16);
  auto id = [n](int x, int y, int z) { return z + y * n + x * n * n; };
  const int a0 = id(2, 2, 2), b0 = id(2, 2, 4);  // extra coupling two cells apart (both directions)
  for (int x = 0; x < n; x++)
    for (int y = 0; y < n; y++)
      for (int z = 0; z < n; z++) {
        int c[9], k = 0;
        const int me = id(x, y, z);
        if (x > 0) c[k++] = id(x - 1, y, z);
        if (y > 0) c[k++] = id(x, y - 1, z);
        if (z > 0) c[k++] = id(x, y, z - 1);
        c[k++] = me;
        if (z + 1 < n) c[k++] = id(x, y, z + 1);
        if (y + 1 < n) c[k++] = id(x, y + 1, z);
        if (x + 1 < n) c[k++] = id(x + 1, y, z);
        if (extra && me == a0) c[k++] = b0;
        if (extra && me == b0) c[k++] = a0;
        std::sort(c, c + k);
        Ai.insert(Ai.end(), c, c + k);
        Ap.push_back((int)Ai.size());
      }
}

static double run(PARTH::ParthAPI &parth, int N, std::vector<int> *Ap, std::vector<int> *Ai, std::vector<int> &perm, int updates) {
  double total = 0;
  for (int k = 0; k < updates; k++) {
    const int m = (k + 1) & 1;
    const auto t0 = std::chrono::steady_clock::now();
    parth.setMatrix(N, Ap[m].data(), Ai[m].data(), 1);
    parth.computePermutation(perm, 1);
    total += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  return total / updates;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 100, updates = argc > 2 ? std::atoi(argv[2]) : 10;
  int N = 0;
  std::vector<int> Ap[2], Ai[2], perm;
  grid(n, false, N, Ap[0], Ai[0]);
  grid(n, true, N, Ap[1], Ai[1]);
  perm.resize((size_t)N);
  try {
    PARTH::ParthAPI parth;
    parth.setVerbose(false);
    parth.setReorderingType(PARTH::ReorderingType::AMD);
    parth.setNDLevels(n >= 64 ? 8 : 4);
    parth.lazy_result_mirrors = true;  // mesh_perm / changed_dof_edges are fetched on demand (syncResultMirror)
    parth.setMatrix(N, Ap[0].data(), Ai[0].data(), 1);
    parth.computePermutation(perm, 1);  // cold start
    run(parth, N, Ap, Ai, perm, 2);     // warm-up
    const double pageable = run(parth, N, Ap, Ai, perm, updates);
    bool pinned = true;
    for (int m = 0; m < 2; m++) pinned = PARTH::ParthAPI::pinHost(Ap[m]) && PARTH::ParthAPI::pinHost(Ai[m]) && pinned;
    pinned = PARTH::ParthAPI::pinHost(perm) && pinned;
    const double locked = pinned ? run(parth, N, Ap, Ai, perm, updates) : -1.0;
    std::vector<char> seen((size_t)N, 0);
    bool ok = true;
    for (int v : perm) { if (v < 0 || v >= N || seen[(size_t)v]) { ok = false; break; } seen[(size_t)v] = 1; }
    for (int m = 0; m < 2; m++) { PARTH::ParthAPI::unpinHost(Ap[m]); PARTH::ParthAPI::unpinHost(Ai[m]); }
    PARTH::ParthAPI::unpinHost(perm);
    std::printf("{\"n\": %d, \"N\": %d, \"nnz\": %d, \"updates\": %d, \"ms_per_update_pageable\": %.3f, "
                "\"ms_per_update_pinned\": %.3f, \"changes\": %d, \"reuse\": %.6f, \"valid_permutation\": %s}\n",
                n, N, Ap[1][(size_t)N], updates, pageable, locked, parth.getNumChanges(), parth.getReuse(), ok ? "true" : "false");
    return ok ? 0 : 1;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "e2e_update: %s\n", e.what());
    return 2;
  }
}
