/*
 * oracle/amd_restated.c -- TEST INFRASTRUCTURE (CPU oracle), never on the product path.
 *
 * Restatement of the approximate-minimum-degree ordering that the reference
 * calls at src/Parth/HMD.cpp:74  (amd_order(N, Mp, Mi, perm, nullptr, nullptr)).
 *
 * The algorithm lives in a third-party dependency that is ABSENT from
 * /root/reference: SuiteSparse AMD, pinned at SuiteSparse v7.11.0
 * (cmake/recipes/suitesparse.cmake:41-47).  This file restates the published
 * algorithm (Amestoy, Davis, Duff, "An approximate minimum degree ordering
 * algorithm", SIMAX 17(4) 1996; ACM TOMS Algorithm 837) with the default
 * control values the reference selects by passing Control == NULL
 * (dense = 10.0, aggressive absorption on), int32 indices:
 *
 *   order  -> validity check -> (jumbled input: sorted duplicate-free transpose)
 *          -> symmetric pattern sizes of A+A' -> build A+A' adjacency in the
 *             sweep order of the original (entry order inside each list is
 *             part of the tie-breaking) -> quotient-graph elimination with
 *             approximate external degrees, element absorption, mass
 *             elimination, supervariable detection by hashing -> assembly
 *             tree post-order with the largest child last.
 *
 * PARITY STATUS: "parity unpinned" versus SuiteSparse v7.11.0 binaries (no
 * SuiteSparse source or library exists in this container, and the reference
 * holds no golden vector at this boundary).  It is
 * pinned only by (a) brute-force invariants in tests/test_oracle_amd.py
 * (valid permutation, exact-minimum-degree fill comparison on small graphs)
 * and (b) an optional cross-check against cusolverSpXcsrsymamdHost on the GPU
 * box.  The CUDA kernel is held bit-exact to THIS restatement.
 *
 * Exported symbol: amd_order (so the reference build in oracle/_ref links it)
 * and oracle_amd_order (same function, unambiguous name for ctypes).
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AMD_OK 0
#define AMD_OUT_OF_MEMORY (-1)
#define AMD_INVALID (-2)
#define AMD_OK_BUT_JUMBLED 1

#define NONE (-1)
#define FLIPI(x) (-(x)-2)

static int amd_check(int n_row, int n_col, const int *Ap, const int *Ai) {
  int result = AMD_OK;
  if (n_row < 0 || n_col < 0 || !Ap || !Ai) return AMD_INVALID;
  if (Ap[0] != 0 || Ap[n_col] < 0) return AMD_INVALID;
  for (int j = 0; j < n_col; j++) {
    int p1 = Ap[j], p2 = Ap[j + 1];
    if (p1 > p2) return AMD_INVALID;
    int ilast = NONE;
    for (int p = p1; p < p2; p++) {
      int i = Ai[p];
      if (i < 0 || i >= n_row) return AMD_INVALID;
      if (i <= ilast) result = AMD_OK_BUT_JUMBLED;
      ilast = i;
    }
  }
  return result;
}

/* R = pattern of A', rows sorted, duplicates dropped */
static void amd_sorted_transpose(int n, const int *Ap, const int *Ai, int *Rp,
                                 int *Ri, int *W, int *Flag) {
  for (int i = 0; i < n; i++) { W[i] = 0; Flag[i] = NONE; }
  for (int j = 0; j < n; j++)
    for (int p = Ap[j]; p < Ap[j + 1]; p++) {
      int i = Ai[p];
      if (Flag[i] != j) { W[i]++; Flag[i] = j; }
    }
  Rp[0] = 0;
  for (int i = 0; i < n; i++) Rp[i + 1] = Rp[i] + W[i];
  for (int i = 0; i < n; i++) { W[i] = Rp[i]; Flag[i] = NONE; }
  for (int j = 0; j < n; j++)
    for (int p = Ap[j]; p < Ap[j + 1]; p++) {
      int i = Ai[p];
      if (Flag[i] != j) { Ri[W[i]++] = j; Flag[i] = j; }
    }
}

/* Len[i] = off-diagonal entries of row/col i in A+A'; returns their sum. */
static long amd_sym_sizes(int n, const int *Ap, const int *Ai, int *Len, int *Tp) {
  for (int k = 0; k < n; k++) Len[k] = 0;
  for (int k = 0; k < n; k++) {
    int p1 = Ap[k], p2 = Ap[k + 1], p;
    for (p = p1; p < p2;) {
      int j = Ai[p];
      if (j < k) { Len[j]++; Len[k]++; p++; }
      else if (j == k) { p++; break; }
      else break;
      int pj2 = Ap[j + 1], pj;
      for (pj = Tp[j]; pj < pj2;) {
        int i = Ai[pj];
        if (i < k) { Len[i]++; Len[j]++; pj++; }
        else if (i == k) { pj++; break; }
        else break;
      }
      Tp[j] = pj;
    }
    Tp[k] = p;
  }
  for (int j = 0; j < n; j++)
    for (int pj = Tp[j]; pj < Ap[j + 1]; pj++) { Len[Ai[pj]]++; Len[j]++; }
  long nz = 0;
  for (int k = 0; k < n; k++) nz += Len[k];
  return nz;
}

static int reset_stamp(int wflg, int wbig, int *W, int n) {
  if (wflg < 2 || wflg >= wbig) {
    for (int x = 0; x < n; x++) if (W[x] != 0) W[x] = 1;
    wflg = 2;
  }
  return wflg;
}

static int post_tree(int root, int k, int *Child, const int *Sibling, int *Order, int *Stack) {
  int head = 0;
  Stack[0] = root;
  while (head >= 0) {
    int i = Stack[head];
    if (Child[i] != NONE) {
      for (int f = Child[i]; f != NONE; f = Sibling[f]) head++;
      int h = head;
      for (int f = Child[i]; f != NONE; f = Sibling[f]) Stack[h--] = f;
      Child[i] = NONE;
    } else {
      head--;
      Order[i] = k++;
    }
  }
  return k;
}

static void postorder(int nn, const int *Parent, const int *Nv, const int *Fsize,
                      int *Order, int *Child, int *Sibling, int *Stack) {
  for (int j = 0; j < nn; j++) { Child[j] = NONE; Sibling[j] = NONE; }
  for (int j = nn - 1; j >= 0; j--)
    if (Nv[j] > 0) {
      int parent = Parent[j];
      if (parent != NONE) { Sibling[j] = Child[parent]; Child[parent] = j; }
    }
  for (int i = 0; i < nn; i++)
    if (Nv[i] > 0 && Child[i] != NONE) {
      int fprev = NONE, maxfr = NONE, bigfprev = NONE, bigf = NONE;
      for (int f = Child[i]; f != NONE; f = Sibling[f]) {
        int fr = Fsize[f];
        if (fr >= maxfr) { maxfr = fr; bigfprev = fprev; bigf = f; }
        fprev = f;
      }
      int fnext = Sibling[bigf];
      if (fnext != NONE) {
        if (bigfprev == NONE) Child[i] = fnext; else Sibling[bigfprev] = fnext;
        Sibling[bigf] = NONE;
        Sibling[fprev] = bigf;
      }
    }
  for (int i = 0; i < nn; i++) Order[i] = NONE;
  int k = 0;
  for (int i = 0; i < nn; i++)
    if (Parent[i] == NONE && Nv[i] > 0) k = post_tree(i, k, Child, Sibling, Order, Stack);
}

/* Quotient-graph elimination.  Pe/Len describe the lists in Iw. */
static void amd_eliminate(int n, int *Pe, int *Iw, int *Len, int iwlen, int pfree,
                          int *Nv, int *Next, int *Last, int *Head, int *Elen,
                          int *Degree, int *W) {
  int me = NONE, mindeg = 0, nel = 0, lemax = 0, ndense = 0;
  const double alpha = 10.0;
  const int aggressive = 1;
  int dense = (alpha < 0) ? n - 2 : (int)(alpha * sqrt((double)n));
  if (dense < 16) dense = 16;
  if (dense > n) dense = n;

  for (int i = 0; i < n; i++) {
    Last[i] = NONE; Head[i] = NONE; Next[i] = NONE;
    Nv[i] = 1; W[i] = 1; Elen[i] = 0; Degree[i] = Len[i];
  }
  int wbig = INT_MAX - n;
  int wflg = reset_stamp(0, wbig, W, n);

  for (int i = 0; i < n; i++) {
    int deg = Degree[i];
    if (deg == 0) {
      Elen[i] = FLIPI(1); nel++; Pe[i] = NONE; W[i] = 0;
    } else if (deg > dense) {
      ndense++; Nv[i] = 0; Elen[i] = NONE; nel++; Pe[i] = NONE;
    } else {
      int inext = Head[deg];
      if (inext != NONE) Last[inext] = i;
      Next[i] = inext;
      Head[deg] = i;
    }
  }
  (void)ndense;

  while (nel < n) {
    int deg;
    for (deg = mindeg; deg < n; deg++) { me = Head[deg]; if (me != NONE) break; }
    mindeg = deg;
    int inext = Next[me];
    if (inext != NONE) Last[inext] = NONE;
    Head[deg] = inext;

    int elenme = Elen[me];
    int nvpiv = Nv[me];
    nel += nvpiv;
    Nv[me] = -nvpiv;
    int degme = 0;
    int pme1, pme2;

    if (elenme == 0) {
      pme1 = Pe[me];
      pme2 = pme1 - 1;
      for (int p = pme1; p <= pme1 + Len[me] - 1; p++) {
        int i = Iw[p];
        int nvi = Nv[i];
        if (nvi > 0) {
          degme += nvi;
          Nv[i] = -nvi;
          Iw[++pme2] = i;
          int ilast = Last[i]; inext = Next[i];
          if (inext != NONE) Last[inext] = ilast;
          if (ilast != NONE) Next[ilast] = inext; else Head[Degree[i]] = inext;
        }
      }
    } else {
      int p = Pe[me];
      pme1 = pfree;
      int slenme = Len[me] - elenme;
      for (int knt1 = 1; knt1 <= elenme + 1; knt1++) {
        int e, pj, ln;
        if (knt1 > elenme) { e = me; pj = p; ln = slenme; }
        else { e = Iw[p++]; pj = Pe[e]; ln = Len[e]; }
        for (int knt2 = 1; knt2 <= ln; knt2++) {
          int i = Iw[pj++];
          int nvi = Nv[i];
          if (nvi > 0) {
            if (pfree >= iwlen) {
              /* compact the workspace, keeping list order */
              Pe[me] = p;
              Len[me] -= knt1;
              if (Len[me] == 0) Pe[me] = NONE;
              Pe[e] = pj;
              Len[e] = ln - knt2;
              if (Len[e] == 0) Pe[e] = NONE;
              for (int j = 0; j < n; j++) {
                int pn = Pe[j];
                if (pn >= 0) { Pe[j] = Iw[pn]; Iw[pn] = FLIPI(j); }
              }
              int psrc = 0, pdst = 0, pend = pme1 - 1;
              while (psrc <= pend) {
                int j = FLIPI(Iw[psrc++]);
                if (j >= 0) {
                  Iw[pdst] = Pe[j];
                  Pe[j] = pdst++;
                  int lenj = Len[j];
                  for (int knt3 = 0; knt3 <= lenj - 2; knt3++) Iw[pdst++] = Iw[psrc++];
                }
              }
              int p1 = pdst;
              for (psrc = pme1; psrc <= pfree - 1; psrc++) Iw[pdst++] = Iw[psrc];
              pme1 = p1;
              pfree = pdst;
              pj = Pe[e];
              p = Pe[me];
            }
            degme += nvi;
            Nv[i] = -nvi;
            Iw[pfree++] = i;
            int ilast = Last[i]; inext = Next[i];
            if (inext != NONE) Last[inext] = ilast;
            if (ilast != NONE) Next[ilast] = inext; else Head[Degree[i]] = inext;
          }
        }
        if (e != me) { Pe[e] = FLIPI(me); W[e] = 0; }
      }
      pme2 = pfree - 1;
    }

    Degree[me] = degme;
    Pe[me] = pme1;
    Len[me] = pme2 - pme1 + 1;
    Elen[me] = FLIPI(nvpiv + degme);

    wflg = reset_stamp(wflg, wbig, W, n);

    /* scan 1: W[e]-wflg = |Le \ Lme| */
    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      int eln = Elen[i];
      if (eln > 0) {
        int nvi = -Nv[i];
        int wnvi = wflg - nvi;
        for (int p = Pe[i]; p <= Pe[i] + eln - 1; p++) {
          int e = Iw[p];
          int we = W[e];
          if (we >= wflg) we -= nvi;
          else if (we != 0) we = Degree[e] + wnvi;
          W[e] = we;
        }
      }
    }

    /* scan 2: degree update, absorption, hashing */
    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      int p1 = Pe[i];
      int p2 = p1 + Elen[i] - 1;
      int pn = p1;
      unsigned int hash = 0;
      deg = 0;
      if (aggressive) {
        for (int p = p1; p <= p2; p++) {
          int e = Iw[p];
          int we = W[e];
          if (we != 0) {
            int dext = we - wflg;
            if (dext > 0) { deg += dext; Iw[pn++] = e; hash += (unsigned int)e; }
            else { Pe[e] = FLIPI(me); W[e] = 0; }
          }
        }
      } else {
        for (int p = p1; p <= p2; p++) {
          int e = Iw[p];
          int we = W[e];
          if (we != 0) { deg += we - wflg; Iw[pn++] = e; hash += (unsigned int)e; }
        }
      }
      Elen[i] = pn - p1 + 1;
      int p3 = pn;
      int p4 = p1 + Len[i];
      for (int p = p2 + 1; p < p4; p++) {
        int j = Iw[p];
        int nvj = Nv[j];
        if (nvj > 0) { deg += nvj; Iw[pn++] = j; hash += (unsigned int)j; }
      }
      if (Elen[i] == 1 && p3 == pn) {
        /* mass elimination */
        Pe[i] = FLIPI(me);
        int nvi = -Nv[i];
        degme -= nvi;
        nvpiv += nvi;
        nel += nvi;
        Nv[i] = 0;
        Elen[i] = NONE;
      } else {
        if (deg < Degree[i]) Degree[i] = deg;
        Iw[pn] = Iw[p3];
        Iw[p3] = Iw[p1];
        Iw[p1] = me;
        Len[i] = pn - p1 + 1;
        hash = hash % (unsigned int)n;
        int j = Head[hash];
        if (j <= NONE) { Next[i] = FLIPI(j); Head[hash] = FLIPI(i); }
        else { Next[i] = Last[j]; Last[j] = i; }
        Last[i] = (int)hash;
      }
    }

    Degree[me] = degme;
    if (degme > lemax) lemax = degme;
    wflg += lemax;
    wflg = reset_stamp(wflg, wbig, W, n);

    /* supervariable detection */
    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      if (Nv[i] < 0) {
        int hash = Last[i];
        int j = Head[hash];
        if (j == NONE) i = NONE;
        else if (j < NONE) { i = FLIPI(j); Head[hash] = NONE; }
        else { i = Last[j]; Last[j] = NONE; }
        while (i != NONE && Next[i] != NONE) {
          int ln = Len[i];
          int eln = Elen[i];
          for (int p = Pe[i] + 1; p <= Pe[i] + ln - 1; p++) W[Iw[p]] = wflg;
          int jlast = i;
          j = Next[i];
          while (j != NONE) {
            int ok = (Len[j] == ln) && (Elen[j] == eln);
            for (int p = Pe[j] + 1; ok && p <= Pe[j] + ln - 1; p++)
              if (W[Iw[p]] != wflg) ok = 0;
            if (ok) {
              Pe[j] = FLIPI(i);
              Nv[i] += Nv[j];
              Nv[j] = 0;
              Elen[j] = NONE;
              j = Next[j];
              Next[jlast] = j;
            } else {
              jlast = j;
              j = Next[j];
            }
          }
          wflg++;
          i = Next[i];
        }
      }
    }

    /* restore degree lists, drop non-principal variables from the element */
    int p = pme1;
    int nleft = n - nel;
    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      int nvi = -Nv[i];
      if (nvi > 0) {
        Nv[i] = nvi;
        deg = Degree[i] + degme - nvi;
        if (deg > nleft - nvi) deg = nleft - nvi;
        inext = Head[deg];
        if (inext != NONE) Last[inext] = i;
        Next[i] = inext;
        Last[i] = NONE;
        Head[deg] = i;
        if (deg < mindeg) mindeg = deg;
        Degree[i] = deg;
        Iw[p++] = i;
      }
    }
    Nv[me] = nvpiv;
    Len[me] = p - pme1;
    if (Len[me] == 0) { Pe[me] = NONE; W[me] = 0; }
    if (elenme != 0) pfree = p;
  }

  /* assembly tree -> output order */
  for (int i = 0; i < n; i++) Pe[i] = FLIPI(Pe[i]);
  for (int i = 0; i < n; i++) Elen[i] = FLIPI(Elen[i]);
  for (int i = 0; i < n; i++) {
    if (Nv[i] == 0) {
      int j = Pe[i];
      if (j == NONE) continue;
      while (Nv[j] == 0) j = Pe[j];
      int e = j;
      j = i;
      while (Nv[j] == 0) { int jnext = Pe[j]; Pe[j] = e; j = jnext; }
    }
  }
  postorder(n, Pe, Nv, Elen, W, Head, Next, Last);
  for (int k = 0; k < n; k++) { Head[k] = NONE; Next[k] = NONE; }
  for (int e = 0; e < n; e++) { int k = W[e]; if (k != NONE) Head[k] = e; }
  nel = 0;
  for (int k = 0; k < n; k++) {
    int e = Head[k];
    if (e == NONE) break;
    Next[e] = nel;
    nel += Nv[e];
  }
  for (int i = 0; i < n; i++) {
    if (Nv[i] == 0) {
      int e = Pe[i];
      if (e != NONE) { Next[i] = Next[e]; Next[e]++; }
      else Next[i] = nel++;
    }
  }
  for (int i = 0; i < n; i++) Last[Next[i]] = i;
}

int oracle_amd_order(int n, const int *Ap, const int *Ai, int *P) {
  if (!Ap || !Ai || !P || n < 0) return AMD_INVALID;
  if (n == 0) return AMD_OK;
  int nz = Ap[n];
  if (nz < 0) return AMD_INVALID;
  int status = amd_check(n, n, Ap, Ai);
  if (status == AMD_INVALID) return AMD_INVALID;

  int *Len = (int *)malloc(sizeof(int) * (size_t)n);
  int *Pinv = (int *)malloc(sizeof(int) * (size_t)n);
  int *Rp = NULL, *Ri = NULL;
  const int *Cp = Ap, *Ci = Ai;
  if (!Len || !Pinv) { free(Len); free(Pinv); return AMD_OUT_OF_MEMORY; }
  if (status == AMD_OK_BUT_JUMBLED) {
    Rp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    Ri = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    if (!Rp || !Ri) { free(Rp); free(Ri); free(Len); free(Pinv); return AMD_OUT_OF_MEMORY; }
    amd_sorted_transpose(n, Ap, Ai, Rp, Ri, Len, Pinv);
    Cp = Rp; Ci = Ri;
  }
  long nzaat = amd_sym_sizes(n, Cp, Ci, Len, P);
  size_t slen = (size_t)nzaat + (size_t)(nzaat / 5) + 7 * (size_t)n;
  int *S = (int *)malloc(sizeof(int) * slen);
  if (!S) { free(Rp); free(Ri); free(Len); free(Pinv); return AMD_OUT_OF_MEMORY; }

  int iwlen = (int)(slen - 6 * (size_t)n);
  int *Pe = S, *Nv = S + n, *Head = S + 2 * (size_t)n, *Elen = S + 3 * (size_t)n;
  int *Degree = S + 4 * (size_t)n, *W = S + 5 * (size_t)n, *Iw = S + 6 * (size_t)n;
  int *Sp = Nv, *Tp = W;
  int pfree = 0;
  for (int j = 0; j < n; j++) { Pe[j] = pfree; Sp[j] = pfree; pfree += Len[j]; }
  for (int k = 0; k < n; k++) {
    int p1 = Cp[k], p2 = Cp[k + 1], p;
    for (p = p1; p < p2;) {
      int j = Ci[p];
      if (j < k) { Iw[Sp[j]++] = k; Iw[Sp[k]++] = j; p++; }
      else if (j == k) { p++; break; }
      else break;
      int pj2 = Cp[j + 1], pj;
      for (pj = Tp[j]; pj < pj2;) {
        int i = Ci[pj];
        if (i < k) { Iw[Sp[i]++] = j; Iw[Sp[j]++] = i; pj++; }
        else if (i == k) { pj++; break; }
        else break;
      }
      Tp[j] = pj;
    }
    Tp[k] = p;
  }
  for (int j = 0; j < n; j++)
    for (int pj = Tp[j]; pj < Cp[j + 1]; pj++) {
      int i = Ci[pj];
      Iw[Sp[i]++] = j; Iw[Sp[j]++] = i;
    }
  amd_eliminate(n, Pe, Iw, Len, iwlen, pfree, Nv, Pinv, P, Head, Elen, Degree, W);
  free(S); free(Rp); free(Ri); free(Len); free(Pinv);
  return status;
}

/* symbol the reference links against (HMD.cpp:74) */
int amd_order(int n, const int *Ap, const int *Ai, int *P, double *Control, double *Info) {
  (void)Control; (void)Info;
  return oracle_amd_order(n, Ap, Ai, P);
}
