// examples/example_creator.cpp -- the reference's test-matrix generator (tests/example_creator.cpp:61-164) against
// the B200 drop-in class.  For each of the reference's five scenarios (pairs of separator-tree nodes, :115-128) it
// starts from the original matrix, reads the DOFs of the two nodes out of the tree mirror
// (parth.hmd.HMD_tree[i].DOFs, :66-67), couples random DOFs of one with random DOFs of the other (both
// directions, :72-83), runs the warm update and writes the modified pattern as <k>_matrix.mtx.
//
// The reference needs libigl (cotangent Laplacian of a triangle mesh), Eigen and CLI11; none is needed here: the
// original matrix is a Matrix-Market pattern (-i) or a 7-point grid (-g n), and the RNG is std::mt19937(seed)
// instead of rand() so that the files are reproducible.  --save-tree / --load-tree exchange the separator tree as
// a fixture (ParthAPI::saveTree) so that another run -- or the Python tests -- update the same tree.
//
//   example_creator -o outdir [-i original.mtx | -g 24] [-e edges_per_pair] [-s seed] [-l levels]
//                   [--save-tree file] [--load-tree file]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <parth/parth.h>

struct Pattern {
  int N = 0;
  std::vector<std::vector<int>> col;  // sorted rows per column
  long nnz() const { long z = 0; for (auto &c : col) z += (long)c.size(); return z; }
  void csc(std::vector<int> &Ap, std::vector<int> &Ai) const {
    Ap.assign(1, 0);
    Ai.clear();
    for (auto &c : col) { Ai.insert(Ai.end(), c.begin(), c.end()); Ap.push_back((int)Ai.size()); }
  }
  // tests/example_creator.cpp:33-58: only add the entry if it does not exist yet
  bool add(int row, int c) {
    auto &v = col[(size_t)c];
    auto it = std::lower_bound(v.begin(), v.end(), row);
    if (it != v.end() && *it == row) return false;
    v.insert(it, row);
    return true;
  }
};

static bool read_mtx(const std::string &path, Pattern &P) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  std::getline(f, line);
  const bool sym = line.find("symmetric") != std::string::npos;
  while (std::getline(f, line) && !line.empty() && line[0] == '%') {}
  int nr = 0, nc = 0;
  long nnz = 0;
  std::istringstream(line) >> nr >> nc >> nnz;
  if (nr <= 0 || nr != nc) return false;
  P.N = nc;
  P.col.assign((size_t)nc, {});
  for (long k = 0; k < nnz && std::getline(f, line); k++) {
    int r = 0, c = 0;
    std::istringstream(line) >> r >> c;
    if (r < 1 || c < 1 || r > nr || c > nc) return false;
    P.col[(size_t)c - 1].push_back(r - 1);
    if (sym && r != c) P.col[(size_t)r - 1].push_back(c - 1);
  }
  for (auto &c : P.col) { std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end()); }
  return true;
}

static bool write_mtx(const std::string &path, const Pattern &P) {
  std::FILE *f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  std::fprintf(f, "%%%%MatrixMarket matrix coordinate pattern general\n%d %d %ld\n", P.N, P.N, P.nnz());
  for (int c = 0; c < P.N; c++)
    for (int r : P.col[(size_t)c]) std::fprintf(f, "%d %d\n", r + 1, c + 1);
  return std::fclose(f) == 0;
}

static void grid(int n, Pattern &P) {  // src/utils/Parth_utils.cpp:40-85 numbering: ind = z + y n + x n^2
  P.N = n * n * n;
  P.col.assign((size_t)P.N, {});
  auto id = [n](int x, int y, int z) { return z + y * n + x * n * n; };
  for (int x = 0; x < n; x++)
    for (int y = 0; y < n; y++)
      for (int z = 0; z < n; z++) {
        auto &c = P.col[(size_t)id(x, y, z)];
        if (x > 0) c.push_back(id(x - 1, y, z));
        if (y > 0) c.push_back(id(x, y - 1, z));
        if (z > 0) c.push_back(id(x, y, z - 1));
        c.push_back(id(x, y, z));
        if (z + 1 < n) c.push_back(id(x, y, z + 1));
        if (y + 1 < n) c.push_back(id(x, y + 1, z));
        if (x + 1 < n) c.push_back(id(x + 1, y, z));
      }
}

// tests/example_creator.cpp:61-84
static void hmd_edges_to_matrix_edges(const std::vector<std::pair<int, int>> &hmd_edges,
                                      std::vector<std::pair<int, int>> &matrix_edges, PARTH::ParthAPI &parth,
                                      int num_edges, std::mt19937 &rng) {
  for (auto &e : hmd_edges) {
    if (e.first >= parth.hmd.num_HMD_nodes || e.second >= parth.hmd.num_HMD_nodes) continue;
    const std::vector<int> &a = parth.hmd.HMD_tree[(size_t)e.first].DOFs;
    const std::vector<int> &b = parth.hmd.HMD_tree[(size_t)e.second].DOFs;
    if (a.empty() || b.empty()) continue;
    for (int k = 0; k < num_edges; k++) {
      const int u = a[rng() % a.size()], v = b[rng() % b.size()];
      if (u != v) {
        matrix_edges.push_back({u, v});
        matrix_edges.push_back({v, u});
      }
    }
  }
}

int main(int argc, char **argv) {
  std::string input, out, save_tree, load_tree;
  int gridn = 24, num_edges = 1, seed = 1, levels = 8;
  for (int i = 1; i < argc; i++) {
    auto next = [&](const char *what) -> const char * {
      if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(1); }
      return argv[++i];
    };
    if (!std::strcmp(argv[i], "-i") || !std::strcmp(argv[i], "--input")) input = next("-i");
    else if (!std::strcmp(argv[i], "-o") || !std::strcmp(argv[i], "--output")) out = next("-o");
    else if (!std::strcmp(argv[i], "-g")) gridn = std::atoi(next("-g"));
    else if (!std::strcmp(argv[i], "-e")) num_edges = std::atoi(next("-e"));
    else if (!std::strcmp(argv[i], "-s")) seed = std::atoi(next("-s"));
    else if (!std::strcmp(argv[i], "-l")) levels = std::atoi(next("-l"));
    else if (!std::strcmp(argv[i], "--save-tree")) save_tree = next("--save-tree");
    else if (!std::strcmp(argv[i], "--load-tree")) load_tree = next("--load-tree");
    else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 1; }
  }
  if (out.empty()) {
    std::fprintf(stderr, "Error: Output folder not specified. Use -o or --output to specify the output folder.\n");
    return 1;
  }
  Pattern OL;
  if (!input.empty()) {
    if (!read_mtx(input, OL)) { std::fprintf(stderr, "Failed to read the matrix: %s\n", input.c_str()); return 1; }
  } else {
    grid(gridn, OL);
  }
  std::printf("Original matrix: %d x %d, %ld entries\nOutput folder: %s\n", OL.N, OL.N, OL.nnz(), out.c_str());
  if (!write_mtx(out + "/original_matrix.mtx", OL)) { std::fprintf(stderr, "cannot write into %s\n", out.c_str()); return 1; }

  // the reference's scenarios (tests/example_creator.cpp:115-128)
  const std::vector<std::vector<std::pair<int, int>>> test_hmd_edges = {
      {{2, 31}},                           // 0: no reuse
      {{0, 1}, {0, 5}, {4, 9}, {4, 10}},   // 1: full reuse (ancestor / descendant pairs)
      {{3, 4}},                            // 2: level-1 reuse
      {{7, 8}},                            // 3: level-2 reuse
      {{15, 16}},                          // 4: level-3 reuse
  };
  std::vector<int> Ap0, Ai0;
  OL.csc(Ap0, Ai0);
  for (size_t t = 0; t < test_hmd_edges.size(); t++) {
    std::mt19937 rng((unsigned)seed + (unsigned)t);
    PARTH::ParthAPI parth;
    parth.setVerbose(false);
    parth.setNDLevels(levels);
    parth.setReorderingType(PARTH::ReorderingType::AMD);
    std::vector<int> perm;
    if (!load_tree.empty()) {
      if (!parth.loadTree(load_tree)) { std::fprintf(stderr, "cannot load the tree fixture %s\n", load_tree.c_str()); return 1; }
    }
    parth.setMatrix(OL.N, Ap0.data(), Ai0.data(), 1);
    parth.computePermutation(perm, 1);  // cold start (or a no-change update of the loaded tree)
    parth.syncTreeMirror();
    if (t == 0 && !save_tree.empty() && !parth.saveTree(save_tree)) {
      std::fprintf(stderr, "cannot save the tree fixture %s\n", save_tree.c_str());
      return 1;
    }
    std::vector<std::pair<int, int>> matrix_edges;
    hmd_edges_to_matrix_edges(test_hmd_edges[t], matrix_edges, parth, num_edges, rng);
    Pattern L = OL;
    long added = 0;
    for (auto &e : matrix_edges) added += L.add(e.first, e.second);
    std::vector<int> Ap, Ai;
    L.csc(Ap, Ai);
    parth.setMatrix(L.N, Ap.data(), Ai.data(), 1);
    parth.computePermutation(perm, 1);
    parth.printTiming();
    std::vector<char> seen(perm.size(), 0);
    for (int v : perm) {
      if (v < 0 || (size_t)v >= perm.size() || seen[(size_t)v]) { std::printf("INVALID permutation\n"); return 2; }
      seen[(size_t)v] = 1;
    }
    if (!write_mtx(out + "/" + std::to_string(t) + "_matrix.mtx", L)) return 1;
    std::printf("test %zu: Number of added edges: %zu\n", t, matrix_edges.size());
    std::printf("test %zu: difference between original and modified matrix: %ld\n", t, L.nnz() - OL.nnz());
    std::printf("test %zu: changes %d reuse %.9f\n\n", t, parth.getNumChanges(), parth.getReuse());
    (void)added;
  }
  return 0;
}
