// parth_b200/csrc/host_decompose.cu -- built-in HOST decomposer for the cold start (pure host code).
//
// The reference builds its separator tree with METIS_ComputeVertexSeparator inside
// HMD::Decompose (reference src/Parth/HMD.cpp:122-346); METIS is third-party, absent from the
// reference tree and outside the device path ("cold decomposition: host").
// A caller can plug any dissection through parth_b200_set_decomposer (e.g. its own METIS); this
// file is the self-contained default so that a cold computePermutation works out of the box:
// recursive bisection by BFS level structures (George's automatic nested dissection): root the
// level structure at a pseudo-peripheral vertex, take the level that splits the vertex count most
// evenly as the separator.  Every separator is exact (no edge joins the two sides), which is the
// only property the reuse path relies on.  Node layout and numbering are the reference's
// (heap order, children 2i+1 / 2i+2, post-order offsets, HMD.cpp:303-337).
#include <algorithm>
#include <queue>
#include <vector>

#include "handle.cuh"

namespace pb {

namespace {

struct Ctx {
  int M;
  const int *Mp, *Mi;
  int levels, nodes;
  std::vector<int> part;   // vertex -> id of the node region it currently belongs to
  std::vector<int> mark;   // BFS level, per call
  std::vector<std::vector<int>> dofs;
  int *meta;
};

// BFS inside the region `reg`; returns the vertices in visit order, level of each in ctx.mark
static void bfs(Ctx &c, int reg, int start, std::vector<int> &order, std::vector<int> &level_start) {
  order.clear();
  level_start.clear();
  c.mark[start] = 0;
  order.push_back(start);
  level_start.push_back(0);
  size_t head = 0;
  int cur = 0;
  while (head < order.size()) {
    const int v = order[head];
    if (c.mark[v] != cur) { cur = c.mark[v]; level_start.push_back((int)head); }
    head++;
    for (int p = c.Mp[v]; p < c.Mp[v + 1]; p++) {
      const int u = c.Mi[p];
      if (c.part[u] == reg && c.mark[u] == -1) { c.mark[u] = c.mark[v] + 1; order.push_back(u); }
    }
  }
  level_start.push_back((int)order.size());
}

static void split(Ctx &c, int node, int parent, int lvl, std::vector<int> &V) {
  int *m = c.meta + 7 * node;
  m[2] = node; m[3] = parent; m[5] = lvl; m[0] = m[1] = -1; m[4] = 0;
  if (lvl == c.levels || (int)V.size() < 3) {
    std::sort(V.begin(), V.end());
    m[6] = (int)V.size();
    c.dofs[node] = V;
    return;
  }
  // connected components of the region (vertices of V carry part == node)
  for (int v : V) c.mark[v] = -1;
  std::vector<int> order, ls;
  std::vector<std::vector<int>> comps;
  for (int v : V)
    if (c.mark[v] == -1) { bfs(c, node, v, order, ls); comps.push_back(order); }
  std::sort(comps.begin(), comps.end(), [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
  std::vector<int> left, right, sep;
  const size_t half = V.size() / 2;
  if (comps[0].size() > half + half / 4 || comps.size() == 1) {
    // dissect the dominant component: pseudo-peripheral root (two BFS sweeps), balanced level
    std::vector<int> &C = comps[0];
    for (int v : C) c.mark[v] = -1;
    bfs(c, node, C[0], order, ls);
    const int far = order.back();
    for (int v : C) c.mark[v] = -1;
    bfs(c, node, far, order, ls);
    const int nl = (int)ls.size() - 1;
    if (nl < 3) {
      // too shallow to cut (e.g. a clique): keep it whole on the left side
      left = C;
    } else {
      int best = 1;
      long long best_cost = -1;
      for (int L = 1; L < nl - 1; L++) {
        const long long a = ls[L], b = (long long)C.size() - ls[L + 1], s = ls[L + 1] - ls[L];
        const long long cost = std::max(a, b) + s;  // balance, lightly penalising thick levels
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = L; }
      }
      left.assign(order.begin(), order.begin() + ls[best]);
      sep.assign(order.begin() + ls[best], order.begin() + ls[best + 1]);
      right.assign(order.begin() + ls[best + 1], order.end());
    }
    for (size_t i = 1; i < comps.size(); i++) {
      std::vector<int> &dst = left.size() <= right.size() ? left : right;
      dst.insert(dst.end(), comps[i].begin(), comps[i].end());
    }
  } else {
    // several comparable components: bin-pack them, no separator needed
    for (auto &C : comps) {
      std::vector<int> &dst = left.size() <= right.size() ? left : right;
      dst.insert(dst.end(), C.begin(), C.end());
    }
  }
  std::sort(sep.begin(), sep.end());
  m[6] = (int)sep.size();
  c.dofs[node] = sep;
  const int l = 2 * node + 1, r = 2 * node + 2;
  for (int v : left) c.part[v] = l;
  for (int v : right) c.part[v] = r;
  for (int v : sep) c.part[v] = -2;
  m[0] = l; m[1] = r;
  V.clear(); V.shrink_to_fit();
  split(c, l, node, lvl + 1, left);
  split(c, r, node, lvl + 1, right);
}

}  // namespace

// same signature as parth_b200_decomposer
int builtin_decomposer(void *, int M_n, const int *Mp, const int *Mi, int levels, int nodes, int *meta,
                       int *node_ptr, int *dofs, int *labels) {
  Ctx c;
  c.M = M_n; c.Mp = Mp; c.Mi = Mi; c.levels = levels; c.nodes = nodes; c.meta = meta;
  c.part.assign((size_t)M_n, 0);
  c.mark.assign((size_t)M_n, -1);
  c.dofs.assign((size_t)nodes, {});
  for (int i = 0; i < nodes * 7; i++) meta[i] = -1;
  std::vector<int> V((size_t)M_n);
  for (int i = 0; i < M_n; i++) V[i] = i;
  split(c, 0, -1, 0, V);
  // flat arrays + identity labels
  int pos = 0;
  for (int x = 0; x < nodes; x++) {
    node_ptr[x] = pos;
    for (size_t k = 0; k < c.dofs[x].size(); k++) { dofs[pos] = c.dofs[x][k]; labels[pos] = (int)k; pos++; }
  }
  node_ptr[nodes] = pos;
  if (pos != M_n) return -1;
  // post-order offsets (HMD.cpp:303-337): slice of a node starts after both sub-trees
  std::vector<long long> sub((size_t)nodes, 0);
  for (int x = nodes - 1; x >= 0; x--) {
    if (meta[x * 7 + 2] == -1) continue;
    sub[x] = (long long)c.dofs[x].size();
    if (meta[x * 7] != -1) sub[x] += sub[meta[x * 7]];
    if (meta[x * 7 + 1] != -1) sub[x] += sub[meta[x * 7 + 1]];
  }
  std::vector<long long> first((size_t)nodes, 0);
  for (int x = 0; x < nodes; x++) {
    if (meta[x * 7 + 2] == -1) continue;
    const int l = meta[x * 7], r = meta[x * 7 + 1];
    const long long lsz = l != -1 ? sub[l] : 0, rsz = r != -1 ? sub[r] : 0;
    if (l != -1) first[l] = first[x];
    if (r != -1) first[r] = first[x] + lsz;
    meta[x * 7 + 4] = (int)(first[x] + lsz + rsz);
  }
  return 0;
}

}  // namespace pb
