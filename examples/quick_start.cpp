// examples/quick_start.cpp -- the flow of the reference's examples/api_demos/quick_start.cpp:72-118
// against the B200 drop-in class: setMatrix(original) -> computePermutation (cold) ->
// setMatrix(modified) -> computePermutation (warm, with reuse) -> printTiming.
//
// The reference demo loads two .mtx files through Eigen (not needed here): this one reads
// Matrix-Market coordinate files directly, or builds a 7-point grid when no file is given.
//   quick_start [original.mtx modified.mtx]
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <parth/parth.h>

static bool read_mtx(const std::string &path, int &N, std::vector<int> &Ap, std::vector<int> &Ai) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  std::getline(f, line);
  const bool sym = line.find("symmetric") != std::string::npos;
  while (std::getline(f, line) && !line.empty() && line[0] == '%') {}
  int nr, nc;
  long nnz;
  std::istringstream(line) >> nr >> nc >> nnz;
  std::vector<std::pair<int, int>> e;  // (col, row)
  for (long k = 0; k < nnz && std::getline(f, line); k++) {
    int r, c;
    std::istringstream(line) >> r >> c;
    e.push_back({c - 1, r - 1});
    if (sym && r != c) e.push_back({r - 1, c - 1});
  }
  std::sort(e.begin(), e.end());
  e.erase(std::unique(e.begin(), e.end()), e.end());
  N = nc;
  Ap.assign((size_t)N + 1, 0);
  Ai.resize(e.size());
  for (size_t k = 0; k < e.size(); k++) { Ap[e[k].first + 1]++; Ai[k] = e[k].second; }
  for (int c = 0; c < N; c++) Ap[c + 1] += Ap[c];
  return true;
}

static void grid(int n, bool extra, int &N, std::vector<int> &Ap, std::vector<int> &Ai) {
  N = n * n * n;
  std::vector<std::vector<int>> col((size_t)N);
  auto id = [n](int x, int y, int z) { return z + y * n + x * n * n; };
  for (int x = 0; x < n; x++)
    for (int y = 0; y < n; y++)
      for (int z = 0; z < n; z++) {
        std::vector<int> &c = col[id(x, y, z)];
        c.push_back(id(x, y, z));
        if (x > 0) c.push_back(id(x - 1, y, z));
        if (x + 1 < n) c.push_back(id(x + 1, y, z));
        if (y > 0) c.push_back(id(x, y - 1, z));
        if (y + 1 < n) c.push_back(id(x, y + 1, z));
        if (z > 0) c.push_back(id(x, y, z - 1));
        if (z + 1 < n) c.push_back(id(x, y, z + 1));
      }
  if (extra) {  // a handful of new couplings, as in the reference's {0..4}_matrix.mtx fixtures
    for (int k = 0; k < 4; k++) {
      int a = id(3 + k, 3, 3), b = id(3 + k, 5, 3);  // two cells apart: a local change
      col[a].push_back(b);
      col[b].push_back(a);
    }
  }
  Ap.assign(1, 0);
  Ai.clear();
  for (auto &c : col) {
    std::sort(c.begin(), c.end());
    Ai.insert(Ai.end(), c.begin(), c.end());
    Ap.push_back((int)Ai.size());
  }
}

int main(int argc, char **argv) {
  int N0, N1;
  std::vector<int> Ap0, Ai0, Ap1, Ai1;
  if (argc >= 3) {
    if (!read_mtx(argv[1], N0, Ap0, Ai0) || !read_mtx(argv[2], N1, Ap1, Ai1)) {
      std::fprintf(stderr, "cannot read %s / %s\n", argv[1], argv[2]);
      return 1;
    }
  } else {
    grid(24, false, N0, Ap0, Ai0);
    grid(24, true, N1, Ap1, Ai1);
  }
  std::printf("original: %d x %d, %d nnz; modified: %d nnz\n", N0, N0, Ap0[N0], Ap1[N1]);

  PARTH::ParthAPI parth;
  parth.setMatrix(N0, Ap0.data(), Ai0.data(), 1);
  std::vector<int> perm;
  parth.computePermutation(perm);  // default dim = 3, like the reference demo (quick_start.cpp:88)
  std::printf("=== TIMING FOR ORIGINAL MATRIX (Cold Start) ===\n");
  parth.printTiming();
  std::printf("Permutation vector size: %zu\n", perm.size());

  parth.setMatrix(N1, Ap1.data(), Ai1.data(), 1);
  parth.computePermutation(perm);
  std::printf("=== TIMING FOR MODIFIED MATRIX (With Reuse) ===\n");
  parth.printTiming();
  std::printf("Permutation vector size: %zu, changed edges: %d, reuse: %.6f\n", perm.size(), parth.getNumChanges(),
              parth.getReuse());

  // validity: perm (dim 3) must be a permutation of 0..3*M_n-1
  std::vector<char> seen(perm.size(), 0);
  for (int v : perm) {
    if (v < 0 || (size_t)v >= perm.size() || seen[(size_t)v]) { std::printf("INVALID permutation\n"); return 2; }
    seen[(size_t)v] = 1;
  }
  std::vector<int> parent, colcount;
  const long long lnz = parth.symbolic(nullptr, parent, colcount);
  std::printf("nnz(L) = %lld (CHOLMOD-style L_NNZ = 2*lnz - N = %lld)\n", lnz, 2 * lnz - parth.M_n);
  std::printf("Both matrices processed successfully!\n");
  return 0;
}
