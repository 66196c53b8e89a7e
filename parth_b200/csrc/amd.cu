// parth_b200/csrc/amd.cu -- stage (c), ordering: batched approximate-minimum-degree ordering of
// the extracted sub-meshes of dirty nodes, one CTA per sub-mesh.
//
// Replaces HMD::AMDNodeNDPermutation -> amd_order(N, Mp, Mi, perm, NULL, NULL)
// (reference src/Parth/HMD.cpp:58-75).  amd_order is SuiteSparse AMD (third party, absent from the
// reference tree, pinned v7.11.0); the algorithm here is the published one (Amestoy, Davis, Duff
// 1996 / ACM TOMS 837) with the default controls that Control == NULL selects (dense = 10,
// aggressive absorption), and it is held bit-exact to the CPU restatement oracle/amd_restated.c:
// same A+A' construction sweep (the entry order of every adjacency list is part of the
// tie-breaking), same degree lists with head insertion, element absorption, mass elimination,
// supervariable hashing, workspace compaction and assembly-tree post-order.
//
// Minimum degree is a sequential pivot recurrence: inside one sub-mesh the pivots are taken by
// lane 0 of the CTA's warp (bit-exact order); parallelism comes from the batch (one CTA per dirty
// node, many nodes in flight) and the work arrays live in an L2-resident slab per CTA.  This stage
// is latency bound and is reported as nodes/s and DOFs/s, outside the HBM-roofline claim
// (SURVEY.md section 8d).
#include <climits>

#include "reorder.cuh"

namespace pb {

#define AMD_NONE (-1)
#define AMD_FLIP(x) (-(x)-2)

__device__ static int amd_reset_stamp(int wflg, int wbig, int *W, int n) {
  if (wflg < 2 || wflg >= wbig) {
    for (int x = 0; x < n; x++) if (W[x] != 0) W[x] = 1;
    wflg = 2;
  }
  return wflg;
}

__device__ static int amd_post_tree(int root, int k, int *Child, const int *Sibling, int *Order, int *Stack) {
  int head = 0;
  Stack[0] = root;
  while (head >= 0) {
    int i = Stack[head];
    if (Child[i] != AMD_NONE) {
      for (int f = Child[i]; f != AMD_NONE; f = Sibling[f]) head++;
      int h = head;
      for (int f = Child[i]; f != AMD_NONE; f = Sibling[f]) Stack[h--] = f;
      Child[i] = AMD_NONE;
    } else {
      head--;
      Order[i] = k++;
    }
  }
  return k;
}

__device__ static void amd_postorder(int nn, const int *Parent, const int *Nv, const int *Fsize,
                                     int *Order, int *Child, int *Sibling, int *Stack) {
  for (int j = 0; j < nn; j++) { Child[j] = AMD_NONE; Sibling[j] = AMD_NONE; }
  for (int j = nn - 1; j >= 0; j--)
    if (Nv[j] > 0) {
      int parent = Parent[j];
      if (parent != AMD_NONE) { Sibling[j] = Child[parent]; Child[parent] = j; }
    }
  for (int i = 0; i < nn; i++)
    if (Nv[i] > 0 && Child[i] != AMD_NONE) {
      int fprev = AMD_NONE, maxfr = AMD_NONE, bigfprev = AMD_NONE, bigf = AMD_NONE;
      for (int f = Child[i]; f != AMD_NONE; f = Sibling[f]) {
        int fr = Fsize[f];
        if (fr >= maxfr) { maxfr = fr; bigfprev = fprev; bigf = f; }
        fprev = f;
      }
      int fnext = Sibling[bigf];
      if (fnext != AMD_NONE) {
        if (bigfprev == AMD_NONE) Child[i] = fnext; else Sibling[bigfprev] = fnext;
        Sibling[bigf] = AMD_NONE;
        Sibling[fprev] = bigf;
      }
    }
  for (int i = 0; i < nn; i++) Order[i] = AMD_NONE;
  int k = 0;
  for (int i = 0; i < nn; i++)
    if (Parent[i] == AMD_NONE && Nv[i] > 0) k = amd_post_tree(i, k, Child, Sibling, Order, Stack);
}

// quotient-graph elimination; Pe/Len describe the lists stored in Iw
__device__ static void amd_eliminate(int n, int *Pe, int *Iw, int *Len, int iwlen, int pfree, int *Nv,
                                     int *Next, int *Last, int *Head, int *Elen, int *Degree, int *W) {
  int me = AMD_NONE, mindeg = 0, nel = 0, lemax = 0;
  int dense = (int)(10.0 * sqrt((double)n));
  if (dense < 16) dense = 16;
  if (dense > n) dense = n;
  for (int i = 0; i < n; i++) {
    Last[i] = AMD_NONE; Head[i] = AMD_NONE; Next[i] = AMD_NONE;
    Nv[i] = 1; W[i] = 1; Elen[i] = 0; Degree[i] = Len[i];
  }
  const int wbig = INT_MAX - n;
  int wflg = amd_reset_stamp(0, wbig, W, n);
  for (int i = 0; i < n; i++) {
    int deg = Degree[i];
    if (deg == 0) {
      Elen[i] = AMD_FLIP(1); nel++; Pe[i] = AMD_NONE; W[i] = 0;
    } else if (deg > dense) {
      Nv[i] = 0; Elen[i] = AMD_NONE; nel++; Pe[i] = AMD_NONE;
    } else {
      int inext = Head[deg];
      if (inext != AMD_NONE) Last[inext] = i;
      Next[i] = inext;
      Head[deg] = i;
    }
  }
  while (nel < n) {
    int deg;
    for (deg = mindeg; deg < n; deg++) { me = Head[deg]; if (me != AMD_NONE) break; }
    mindeg = deg;
    int inext = Next[me];
    if (inext != AMD_NONE) Last[inext] = AMD_NONE;
    Head[deg] = inext;
    const int elenme = Elen[me];
    int nvpiv = Nv[me];
    nel += nvpiv;
    Nv[me] = -nvpiv;
    int degme = 0, pme1, pme2;
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
          if (inext != AMD_NONE) Last[inext] = ilast;
          if (ilast != AMD_NONE) Next[ilast] = inext; else Head[Degree[i]] = inext;
        }
      }
    } else {
      int p = Pe[me];
      pme1 = pfree;
      const int slenme = Len[me] - elenme;
      for (int knt1 = 1; knt1 <= elenme + 1; knt1++) {
        int e, pj, ln;
        if (knt1 > elenme) { e = me; pj = p; ln = slenme; }
        else { e = Iw[p++]; pj = Pe[e]; ln = Len[e]; }
        for (int knt2 = 1; knt2 <= ln; knt2++) {
          int i = Iw[pj++];
          int nvi = Nv[i];
          if (nvi > 0) {
            if (pfree >= iwlen) {
              // compact the workspace, keeping list order
              Pe[me] = p;
              Len[me] -= knt1;
              if (Len[me] == 0) Pe[me] = AMD_NONE;
              Pe[e] = pj;
              Len[e] = ln - knt2;
              if (Len[e] == 0) Pe[e] = AMD_NONE;
              for (int j = 0; j < n; j++) {
                int pn = Pe[j];
                if (pn >= 0) { Pe[j] = Iw[pn]; Iw[pn] = AMD_FLIP(j); }
              }
              int psrc = 0, pdst = 0;
              const int pend = pme1 - 1;
              while (psrc <= pend) {
                int j = AMD_FLIP(Iw[psrc++]);
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
            if (inext != AMD_NONE) Last[inext] = ilast;
            if (ilast != AMD_NONE) Next[ilast] = inext; else Head[Degree[i]] = inext;
          }
        }
        if (e != me) { Pe[e] = AMD_FLIP(me); W[e] = 0; }
      }
      pme2 = pfree - 1;
    }
    Degree[me] = degme;
    Pe[me] = pme1;
    Len[me] = pme2 - pme1 + 1;
    Elen[me] = AMD_FLIP(nvpiv + degme);
    wflg = amd_reset_stamp(wflg, wbig, W, n);

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

    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      int p1 = Pe[i];
      int p2 = p1 + Elen[i] - 1;
      int pn = p1;
      unsigned int hash = 0;
      deg = 0;
      for (int p = p1; p <= p2; p++) {
        int e = Iw[p];
        int we = W[e];
        if (we != 0) {
          int dext = we - wflg;
          if (dext > 0) { deg += dext; Iw[pn++] = e; hash += (unsigned int)e; }
          else { Pe[e] = AMD_FLIP(me); W[e] = 0; }
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
        Pe[i] = AMD_FLIP(me);
        int nvi = -Nv[i];
        degme -= nvi;
        nvpiv += nvi;
        nel += nvi;
        Nv[i] = 0;
        Elen[i] = AMD_NONE;
      } else {
        if (deg < Degree[i]) Degree[i] = deg;
        Iw[pn] = Iw[p3];
        Iw[p3] = Iw[p1];
        Iw[p1] = me;
        Len[i] = pn - p1 + 1;
        hash = hash % (unsigned int)n;
        int j = Head[hash];
        if (j <= AMD_NONE) { Next[i] = AMD_FLIP(j); Head[hash] = AMD_FLIP(i); }
        else { Next[i] = Last[j]; Last[j] = i; }
        Last[i] = (int)hash;
      }
    }
    Degree[me] = degme;
    if (degme > lemax) lemax = degme;
    wflg += lemax;
    wflg = amd_reset_stamp(wflg, wbig, W, n);

    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      if (Nv[i] < 0) {
        int hash = Last[i];
        int j = Head[hash];
        if (j == AMD_NONE) i = AMD_NONE;
        else if (j < AMD_NONE) { i = AMD_FLIP(j); Head[hash] = AMD_NONE; }
        else { i = Last[j]; Last[j] = AMD_NONE; }
        while (i != AMD_NONE && Next[i] != AMD_NONE) {
          int ln = Len[i];
          int eln = Elen[i];
          for (int p = Pe[i] + 1; p <= Pe[i] + ln - 1; p++) W[Iw[p]] = wflg;
          int jlast = i;
          j = Next[i];
          while (j != AMD_NONE) {
            int ok = (Len[j] == ln) && (Elen[j] == eln);
            for (int p = Pe[j] + 1; ok && p <= Pe[j] + ln - 1; p++)
              if (W[Iw[p]] != wflg) ok = 0;
            if (ok) {
              Pe[j] = AMD_FLIP(i);
              Nv[i] += Nv[j];
              Nv[j] = 0;
              Elen[j] = AMD_NONE;
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

    int p = pme1;
    const int nleft = n - nel;
    for (int pme = pme1; pme <= pme2; pme++) {
      int i = Iw[pme];
      int nvi = -Nv[i];
      if (nvi > 0) {
        Nv[i] = nvi;
        deg = Degree[i] + degme - nvi;
        if (deg > nleft - nvi) deg = nleft - nvi;
        inext = Head[deg];
        if (inext != AMD_NONE) Last[inext] = i;
        Next[i] = inext;
        Last[i] = AMD_NONE;
        Head[deg] = i;
        if (deg < mindeg) mindeg = deg;
        Degree[i] = deg;
        Iw[p++] = i;
      }
    }
    Nv[me] = nvpiv;
    Len[me] = p - pme1;
    if (Len[me] == 0) { Pe[me] = AMD_NONE; W[me] = 0; }
    if (elenme != 0) pfree = p;
  }

  for (int i = 0; i < n; i++) Pe[i] = AMD_FLIP(Pe[i]);
  for (int i = 0; i < n; i++) Elen[i] = AMD_FLIP(Elen[i]);
  for (int i = 0; i < n; i++) {
    if (Nv[i] == 0) {
      int j = Pe[i];
      if (j == AMD_NONE) continue;
      while (Nv[j] == 0) j = Pe[j];
      int e = j;
      j = i;
      while (Nv[j] == 0) { int jnext = Pe[j]; Pe[j] = e; j = jnext; }
    }
  }
  amd_postorder(n, Pe, Nv, Elen, W, Head, Next, Last);
  for (int k = 0; k < n; k++) { Head[k] = AMD_NONE; Next[k] = AMD_NONE; }
  for (int e = 0; e < n; e++) { int k = W[e]; if (k != AMD_NONE) Head[k] = e; }
  nel = 0;
  for (int k = 0; k < n; k++) {
    int e = Head[k];
    if (e == AMD_NONE) break;
    Next[e] = nel;
    nel += Nv[e];
  }
  for (int i = 0; i < n; i++) {
    if (Nv[i] == 0) {
      int e = Pe[i];
      if (e != AMD_NONE) { Next[i] = Next[e]; Next[e]++; }
      else Next[i] = nel++;
    }
  }
  for (int i = 0; i < n; i++) Last[Next[i]] = i;
}

// One CTA (one warp) per graph g.  Graph g has vertices [vtx_ptr[g], vtx_ptr[g+1]) of the
// concatenated vertex space; its CSR is (G + vtx_ptr[g]) with global entry offsets (subtract
// G[vtx_ptr[g]]) into sub_Mi.  The input is sorted, duplicate- and diagonal-free and structurally
// symmetric (what the extraction kernels produce), so nzaat = nnz and no preprocessing applies.
// Workspace slab of graph g: ws + ws_ptr[g], 8n + slen ints with slen = nnz + nnz/5 + 7n.
// perm_out[vtx_ptr[g] + k] = k-th pivot (local id).  Zero-edge graphs get the identity
// (HMD.cpp:62-72).
__global__ void amd_batch_kernel(int count, const int *__restrict__ vtx_ptr, const int *__restrict__ G,
                                 const int *__restrict__ sub_Mi, const long long *__restrict__ ws_ptr,
                                 int *__restrict__ ws, int *__restrict__ perm_out, int smem_ints) {
  const int g = blockIdx.x;
  if (g >= count || threadIdx.x != 0) return;
  const int v0 = vtx_ptr[g], n = vtx_ptr[g + 1] - v0;
  if (n <= 0) return;
  const int *Cp = G + v0;
  const int e0 = Cp[0];
  const int nz = Cp[n] - e0;
  int *P = perm_out + v0;
  if (nz == 0) { for (int i = 0; i < n; i++) P[i] = i; return; }
  const int *Ci = sub_Mi + e0;
  // quotient-graph staging: the whole work-space of a graph lives in shared memory when it fits
  // (separators and small leaves: ~10x lower latency per dependent access), else in its slab
  extern __shared__ int amd_smem[];
  const long long need = 2 * (long long)n + (long long)nz + nz / 5 + 7 * (long long)n + 16;
  const bool in_smem = need + n <= (long long)smem_ints;
  int *slab = in_smem ? amd_smem : ws + ws_ptr[g];
  int *const Pout = P;
  if (in_smem) P = amd_smem + need;  // the pivot / Last array too; copied out at the end
  int *Len = slab, *Pinv = slab + n, *S = slab + 2 * (size_t)n;
  // symmetric pattern sizes (amd_aat) with P as the Tp workspace
  int *Tp = P;
  for (int k = 0; k < n; k++) Len[k] = 0;
  for (int k = 0; k < n; k++) {
    const int p1 = Cp[k] - e0, p2 = Cp[k + 1] - e0;
    int p;
    for (p = p1; p < p2;) {
      int j = Ci[p];
      if (j < k) { Len[j]++; Len[k]++; p++; }
      else if (j == k) { p++; break; }
      else break;
      const int pj2 = Cp[j + 1] - e0;
      int pj;
      for (pj = Tp[j]; pj < pj2;) {
        int i = Ci[pj];
        if (i < k) { Len[i]++; Len[j]++; pj++; }
        else if (i == k) { pj++; break; }
        else break;
      }
      Tp[j] = pj;
    }
    Tp[k] = p;
  }
  for (int j = 0; j < n; j++)
    for (int pj = Tp[j]; pj < Cp[j + 1] - e0; pj++) { Len[Ci[pj]]++; Len[j]++; }
  long long nzaat = 0;
  for (int k = 0; k < n; k++) nzaat += Len[k];
  const long long slen = nzaat + nzaat / 5 + 7 * (long long)n;
  const int iwlen = (int)(slen - 6 * (long long)n);
  int *Pe = S, *Nv = S + n, *Head = S + 2 * (size_t)n, *Elen = S + 3 * (size_t)n;
  int *Degree = S + 4 * (size_t)n, *W = S + 5 * (size_t)n, *Iw = S + 6 * (size_t)n;
  int *Sp = Nv;
  Tp = W;
  int pfree = 0;
  for (int j = 0; j < n; j++) { Pe[j] = pfree; Sp[j] = pfree; pfree += Len[j]; }
  for (int k = 0; k < n; k++) {
    const int p1 = Cp[k] - e0, p2 = Cp[k + 1] - e0;
    int p;
    for (p = p1; p < p2;) {
      int j = Ci[p];
      if (j < k) { Iw[Sp[j]++] = k; Iw[Sp[k]++] = j; p++; }
      else if (j == k) { p++; break; }
      else break;
      const int pj2 = Cp[j + 1] - e0;
      int pj;
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
    for (int pj = Tp[j]; pj < Cp[j + 1] - e0; pj++) {
      int i = Ci[pj];
      Iw[Sp[i]++] = j; Iw[Sp[j]++] = i;
    }
  amd_eliminate(n, Pe, Iw, Len, iwlen, pfree, Nv, Pinv, P, Head, Elen, Degree, W);
  if (in_smem) for (int i = 0; i < n; i++) Pout[i] = P[i];
}

int launch_amd_batch(int count, const int *vtx_ptr, const int *G, const int *sub_Mi,
                     const long long *ws_ptr, int *ws, int *perm_out, cudaStream_t s) {
  if (count <= 0) return 0;
  int dev = 0, max_smem = 0;
  PB_CUDA(cudaGetDevice(&dev));
  PB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int smem_bytes = max_smem - 1024;
  PB_CUDA(cudaFuncSetAttribute(amd_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  amd_batch_kernel<<<count, 32, smem_bytes, s>>>(count, vtx_ptr, G, sub_Mi, ws_ptr, ws, perm_out, smem_bytes / 4);
  PB_CUDA(cudaGetLastError());
  return 1;
}

// host-side slab size of one graph (ints)
long long amd_slab_ints(int n, int nnz) {
  long long nz = nnz;
  return 2 * (long long)n + nz + nz / 5 + 7 * (long long)n + 16;
}

}  // namespace pb
