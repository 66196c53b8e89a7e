// examples/sharded_update.cpp -- ONE mesh over several GPUs through the drop-in C++ class, no Python and no torchrun:
// the collectives (NCCL all-gather / all-reduce over NVLink) run inside PARTH::ParthAPI::setMatrix / computePermutation.
//
//   sharded_update [--procs W | --threads W] [-g GRID] [-l LEVELS] [-s STEPS] [-e EDGES]
//
//   --procs W    W processes, rank r on GPU r (fork before any CUDA call; rank 0 writes the NCCL unique id to a
//                temporary file the others read -- any transport of 128 bytes would do, MPI_Bcast included)
//   --threads W  W threads of this process, all on GPU 0, in-process communicator (runs on a single-GPU box)
//
// Every rank also keeps a private single-GPU object holding the whole graph and checks, update by update, that the
// sharded permutation, the number of changed edges and the reuse ratio are bit-for-bit the single-GPU ones.  The
// separator tree comes from that object's cold start and is imported into the sharded one (importTree), so both update
// an identical tree -- the role the reference's first METIS computePermutation plays (ParthAPI.cpp:191-279).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include <parth/parth.h>

struct Csc { int N = 0; std::vector<int> Ap, Ai; };

// 7-point grid + extra symmetric couplings (a, b)
static Csc grid_matrix(int n, const std::vector<std::pair<int, int>> &extra) {
  Csc A;
  A.N = n * n * n;
  std::vector<std::vector<int>> col((size_t)A.N);
  auto id = [n](int x, int y, int z) { return z + y * n + x * n * n; };
  for (int x = 0; x < n; x++)
    for (int y = 0; y < n; y++)
      for (int z = 0; z < n; z++) {
        std::vector<int> &c = col[(size_t)id(x, y, z)];
        c.push_back(id(x, y, z));
        if (x > 0) c.push_back(id(x - 1, y, z));
        if (x + 1 < n) c.push_back(id(x + 1, y, z));
        if (y > 0) c.push_back(id(x, y - 1, z));
        if (y + 1 < n) c.push_back(id(x, y + 1, z));
        if (z > 0) c.push_back(id(x, y, z - 1));
        if (z + 1 < n) c.push_back(id(x, y, z + 1));
      }
  for (const auto &e : extra)
    if (e.first != e.second) { col[(size_t)e.first].push_back(e.second); col[(size_t)e.second].push_back(e.first); }
  A.Ap.assign(1, 0);
  for (auto &c : col) {
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    A.Ai.insert(A.Ai.end(), c.begin(), c.end());
    A.Ap.push_back((int)A.Ai.size());
  }
  return A;
}

struct Options { int world = 2, grid = 24, levels = 5, steps = 4, edges = 30; bool threads = false; };

// the update stream is a pure function of (step): every rank generates the same matrices
static Csc matrix_of_step(const Options &o, int step) {
  std::vector<std::pair<int, int>> extra;
  unsigned long long x = 88172645463325252ull;
  const int M = o.grid * o.grid * o.grid;
  for (int s = 0; s <= step; s++)
    for (int k = 0; k < o.edges; k++) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      const int a = (int)(x % (unsigned long long)M);
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      const int b = (int)(x % (unsigned long long)M);
      if (s > 0) extra.push_back({a, b});  // step 0 is the base grid
    }
  return grid_matrix(o.grid, extra);
}

static void configure(PARTH::ParthAPI &p, const Options &o) {
  p.setVerbose(false);
  p.setNDLevels(o.levels);
  p.setReorderingType(PARTH::ReorderingType::AMD);
  p.lagging = false;  // dirty coarse nodes are repaired and re-ordered (stage c runs)
  p.lazy_result_mirrors = true;
}

// one rank: returns 0 when every update matched the single-GPU object
static int run_rank(const Options &o, int rank, int device, PARTH::ParthAPI &sharded) {
  PARTH::ParthAPI whole(device);
  configure(whole, o);
  const Csc A0 = matrix_of_step(o, 0);
  std::vector<int> perm, perm_ref;
  whole.setMatrix(A0.N, const_cast<int *>(A0.Ap.data()), const_cast<int *>(A0.Ai.data()), 1);
  whole.computePermutation(perm_ref, 1);  // cold start: builds the separator tree
  whole.syncTreeMirror();
  sharded.importTree(whole.hmd);
  sharded.setMatrix(A0.N, const_cast<int *>(A0.Ap.data()), const_cast<int *>(A0.Ai.data()), 1);  // this rank's block only
  sharded.acceptSnapshot();
  int c0 = 0, c1 = 0;
  sharded.blockRange(A0.N, A0.Ap[(size_t)A0.N], c0, c1);
  std::printf("rank %d: columns [%d, %d) of %d\n", rank, c0, c1, A0.N);
  int bad = 0;
  for (int step = 1; step <= o.steps; step++) {
    const Csc A = matrix_of_step(o, step);
    whole.setMatrix(A.N, const_cast<int *>(A.Ap.data()), const_cast<int *>(A.Ai.data()), 1);
    whole.computePermutation(perm_ref, 1);
    const auto t0 = std::chrono::steady_clock::now();
    sharded.setMatrix(A.N, const_cast<int *>(A.Ap.data()), const_cast<int *>(A.Ai.data()), 1);
    sharded.computePermutation(perm, 1);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const bool same = perm == perm_ref && sharded.getNumChanges() == whole.getNumChanges() &&
                      sharded.getReuse() == whole.getReuse();
    if (!same) bad++;
    if (rank == 0)
      std::printf("update %d: changes %d reuse %.6f  sharded %.3f ms  %s\n", step, sharded.getNumChanges(),
                  sharded.getReuse(), ms, same ? "identical to the single-GPU permutation" : "MISMATCH");
  }
  return bad;
}

int main(int argc, char **argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() { return i + 1 < argc ? std::atoi(argv[++i]) : 0; };
    if (a == "--procs") { o.world = next(); o.threads = false; }
    else if (a == "--threads") { o.world = next(); o.threads = true; }
    else if (a == "-g") o.grid = next();
    else if (a == "-l") o.levels = next();
    else if (a == "-s") o.steps = next();
    else if (a == "-e") o.edges = next();
    else { std::fprintf(stderr, "usage: sharded_update [--procs W | --threads W] [-g GRID] [-l LEVELS] [-s STEPS] [-e EDGES]\n"); return 2; }
  }
  if (o.world < 1 || o.grid < 8) return 2;
  // the cout line of the reference ("number of bad edges") is kept by the drop-in class; silence it here
  std::cout.setstate(std::ios_base::failbit);

  if (o.threads) {
    std::vector<PARTH::ParthAPI *> ranks;
    for (int r = 0; r < o.world; r++) { ranks.push_back(new PARTH::ParthAPI(0)); configure(*ranks.back(), o); }
    PARTH::ParthAPI::commInitLocal(ranks);
    std::vector<int> bad((size_t)o.world, -1);
    std::vector<std::thread> th;
    for (int r = 0; r < o.world; r++)
      th.emplace_back([&, r]() {
        try { bad[(size_t)r] = run_rank(o, r, 0, *ranks[(size_t)r]); }
        catch (const std::exception &e) { std::fprintf(stderr, "rank %d: %s\n", r, e.what()); }
      });
    for (auto &t : th) t.join();
    for (auto *p : ranks) delete p;
    const bool ok = std::all_of(bad.begin(), bad.end(), [](int b) { return b == 0; });
    std::printf("%s (%d ranks as threads, in-process communicator)\n", ok ? "PASS" : "FAIL", o.world);
    return ok ? 0 : 1;
  }

  // one process per GPU: fork BEFORE the first CUDA call
  char path[] = "/tmp/parth_b200_ncclid_XXXXXX";
  const int fd = mkstemp(path);
  if (fd < 0) return 3;
  close(fd);
  std::vector<pid_t> kids;
  for (int r = 0; r < o.world; r++) {
    const pid_t pid = fork();
    if (pid == 0) {
      int rc = 1;
      try {
        std::vector<unsigned char> id(128);
        const std::string ready = std::string(path) + ".ready";
        if (r == 0) {
          id = PARTH::ParthAPI::commUniqueId();
          std::ofstream(path, std::ios::binary).write(reinterpret_cast<const char *>(id.data()), 128);
          std::ofstream(ready) << "1";
        } else {
          for (int spin = 0; spin < 6000 && access(ready.c_str(), F_OK) != 0; spin++) usleep(10000);
          std::ifstream(path, std::ios::binary).read(reinterpret_cast<char *>(id.data()), 128);
        }
        PARTH::ParthAPI p(r);
        configure(p, o);
        p.commInit(id.data(), r, o.world);
        rc = run_rank(o, r, r, p) == 0 ? 0 : 1;
      } catch (const std::exception &e) {
        std::fprintf(stderr, "rank %d: %s\n", r, e.what());
      }
      std::fflush(stdout);
      _exit(rc);
    }
    kids.push_back(pid);
  }
  bool ok = true;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    ok = ok && WIFEXITED(st) && WEXITSTATUS(st) == 0;
  }
  unlink(path);
  unlink((std::string(path) + ".ready").c_str());
  std::printf("%s (%d processes, NCCL)\n", ok ? "PASS" : "FAIL", o.world);
  return ok ? 0 : 1;
}
