/*
 * oracle/symbolic_oracle.c -- TEST INFRASTRUCTURE (CPU oracle, kind "port").
 *
 * Stage (d) of the path: elimination tree, column counts and nnz(L) of P*A*P'.
 *
 * Where the reference gets them: src/Parth itself never computes them; the reference ecosystem
 * calls CHOLMOD (cholmod_analyze_p with a given permutation, L_NNZ = cm.lnz*2 - N;
 * src/utils/ParthTestUtils.cpp:15-59, src/LinSysSolver/CHOLMODSolver.cpp:108-151).  CHOLMOD is a
 * third-party dependency ABSENT from /root/reference (SuiteSparse v7.11.0,
 * cmake/recipes/suitesparse.cmake:41-47).  Restated here from the published algorithms:
 *   - etree: Liu's algorithm with path compression, as the reference's own in-tree copy does it
 *     on the upper triangle (sym_lib::compute_etree, src/utils/etree.cpp:9-115, stype > 0 branch
 *     :72-84);
 *   - column counts: Gilbert, Ng, Peyton, "An efficient algorithm to compute row and column counts
 *     for sparse Cholesky factorization", SIMAX 15(4) 1994 (skeleton matrix, first descendants,
 *     least common ancestors by union-find), diagonal included.
 * nnz(L) is a mathematical function of (pattern, permutation); pinned by the known answers of
 * SURVEY.md Appendix B.3 and by brute-force symbolic elimination in tests/test_oracle_symbolic.py;
 * the etree is additionally compared with the reference's compute_etree through oracle/_ref.
 */
#include "cpu_parth.h"

#include <stdlib.h>
#include <string.h>

/* etree of P*A*P'; (Ap,Ai) full symmetric pattern, perm[new] = old */
int cpu_parth_etree(int n, const int *Ap, const int *Ai, const int *perm, int *parent) {
  int *inv = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *anc = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int k = 0; k < n; k++) inv[perm[k]] = k;
  for (int k = 0; k < n; k++) { parent[k] = -1; anc[k] = -1; }
  for (int j = 0; j < n; j++) {
    int old = perm[j];
    for (int p = Ap[old]; p < Ap[old + 1]; p++) {
      int k = inv[Ai[p]];
      if (k >= j) continue;
      /* walk from k to the root of its current tree, compressing onto j */
      for (;;) {
        int a = anc[k];
        if (a == j) break;
        anc[k] = j;
        if (a == -1) { parent[k] = j; break; }
        k = a;
      }
    }
  }
  free(inv); free(anc);
  return CPU_PARTH_OK;
}

static void postorder_forest(int n, const int *parent, int *post) {
  int *head = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *next = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *stack = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int j = 0; j < n; j++) head[j] = -1;
  for (int j = n - 1; j >= 0; j--) {
    if (parent[j] == -1) continue;
    next[j] = head[parent[j]];
    head[parent[j]] = j;
  }
  int k = 0;
  for (int root = 0; root < n; root++) {
    if (parent[root] != -1) continue;
    int top = 0;
    stack[0] = root;
    while (top >= 0) {
      int p = stack[top];
      int i = head[p];
      if (i == -1) { top--; post[k++] = p; }
      else { head[p] = next[i]; stack[++top] = i; }
    }
  }
  free(head); free(next); free(stack);
}

int64_t cpu_parth_colcount(int n, const int *Ap, const int *Ai, const int *perm,
                           const int *parent, int *colcount) {
  if (n == 0) return 0;
  int *inv = (int *)malloc(sizeof(int) * (size_t)n);
  int *post = (int *)malloc(sizeof(int) * (size_t)n);
  int *w = (int *)malloc(sizeof(int) * 4 * (size_t)n);
  int *ancestor = w, *maxfirst = w + n, *prevleaf = w + 2 * (size_t)n, *first = w + 3 * (size_t)n;
  int *delta = colcount;
  for (int k = 0; k < n; k++) inv[perm[k]] = k;
  postorder_forest(n, parent, post);
  for (size_t k = 0; k < 4 * (size_t)n; k++) w[k] = -1;
  for (int k = 0; k < n; k++) {
    int j = post[k];
    delta[j] = (first[j] == -1) ? 1 : 0;
    for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
  }
  for (int i = 0; i < n; i++) ancestor[i] = i;
  for (int k = 0; k < n; k++) {
    int j = post[k];
    if (parent[j] != -1) delta[parent[j]]--;
    int old = perm[j];
    for (int p = Ap[old]; p < Ap[old + 1]; p++) {
      int i = inv[Ai[p]];
      if (i <= j || first[j] <= maxfirst[i]) continue;
      maxfirst[i] = first[j];
      int jprev = prevleaf[i];
      prevleaf[i] = j;
      if (jprev == -1) { delta[j]++; continue; } /* first leaf of row subtree i */
      int q = jprev;
      while (q != ancestor[q]) q = ancestor[q];
      for (int s = jprev; s != q;) { int sp = ancestor[s]; ancestor[s] = q; s = sp; }
      delta[j]++;
      delta[q]--;
    }
    if (parent[j] != -1) ancestor[j] = parent[j];
  }
  for (int j = 0; j < n; j++)
    if (parent[j] != -1) colcount[parent[j]] += colcount[j];
  int64_t lnz = 0;
  for (int j = 0; j < n; j++) lnz += colcount[j];
  free(inv); free(post); free(w);
  return lnz;
}

/* (f4) fundamental supernodes of L from (parent, colcount), sequentially: the definition CHOLMOD's supernodal
 * analysis starts from (the step behind reference src/LinSysSolver/CHOLMODSolver.cpp:108-151; SuiteSparse itself is
 * absent from the reference tree -- PARITY UNPINNED against CHOLMOD, pinned by the brute-force symbolic
 * factorisation in tests/test_oracle_symbolic.py).  Column j > 0 continues the supernode of j-1 iff
 * parent[j-1] == j, colcount[j-1] == colcount[j] + 1 and j has exactly one child.  super has n+1 entries,
 * supermap n, sparent n (first nsuper used); sizes[0] = sum colcount[first], sizes[1] = sum ncols*colcount[first].
 * Returns nsuper. */
int cpu_parth_supernodes(int n, const int *parent, const int *colcount, int *super, int *supermap,
                         int *sparent, int64_t *sizes) {
  sizes[0] = sizes[1] = 0;
  if (n == 0) { super[0] = 0; return 0; }
  int *nchild = (int *)calloc((size_t)n, sizeof(int));
  for (int j = 0; j < n; j++)
    if (parent[j] >= 0 && parent[j] < n) nchild[parent[j]]++;
  int ns = 0;
  for (int j = 0; j < n; j++) {
    if (j == 0 || parent[j - 1] != j || colcount[j - 1] != colcount[j] + 1 || nchild[j] > 1) super[ns++] = j;
    supermap[j] = ns - 1;
  }
  super[ns] = n;
  for (int s = 0; s < ns; s++) {
    int last = super[s + 1] - 1, p = parent[last];
    sparent[s] = (p < 0 || p >= n) ? -1 : supermap[p];
    sizes[0] += colcount[super[s]];
    sizes[1] += (int64_t)colcount[super[s]] * (super[s + 1] - super[s]);
  }
  free(nchild);
  return ns;
}
