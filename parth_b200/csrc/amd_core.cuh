// parth_b200/csrc/amd_core.cuh -- warp-cooperative approximate minimum degree, bit-exact to
// oracle/amd_restated.c (= amd_order of SuiteSparse AMD 7.11.0 with default controls, the call of
// HMD::AMDNodeNDPermutation, reference src/Parth/HMD.cpp:58-75).
//
// Minimum degree is a pivot recurrence: pivot k+1 is chosen from the degree lists pivot k leaves
// behind, so the pivots of ONE graph stay sequential.  Everything inside a pivot is spread over the
// lanes of a warp, in a form whose result does not depend on the lane schedule:
//   P0 pick     lanes probe 32 consecutive degree-list heads at once
//   P1 Lme      every list that feeds the new element is read 32 entries at a time, compacted
//               with a ballot (first occurrence wins: Nv[i] is negated chunk by chunk), degree-list
//               removals run in parallel unless two list neighbours leave in the same chunk
//   P2 |Le\Lme| one lane per variable of Lme; W[e] = Degree[e] + wflg - sum nv_i by atomic
//               subtraction (integer sums: order-free)
//   P3 degrees  one lane per variable: in-place compaction of its own element / variable lists,
//               approximate degree, hash; aggressive absorption writes are idempotent; mass
//               elimination totals are warp sums; hash buckets are chained per chunk with
//               match_any so that each bucket keeps the sequential (LIFO) order
//   P4 supervariables  buckets with >= 2 members are walked in Lme order (the principal variable
//               is decided by that order); list marking and comparison are lane-parallel
//   P5 re-insert  ballot compaction of Lme, degree lists pushed per chunk with match_any (the
//               head-insertion order inside one degree list IS the tie-breaking)
// The stamp values (wflg) and the moment of workspace compaction differ from a sequential run;
// neither can change the ordering (stamps are only compared for equality / order against wflg,
// compaction moves lists without reordering them).
//
// Written against lanes.cuh so that tests/amd_model can run the same source on the CPU.
#pragma once
#include <climits>
#include <cmath>

namespace pb {
namespace amdk {

using namespace lanes;

constexpr int NONE = -1;
PB_DEV int flip(int x) { return -x - 2; }

struct Ws {
  int n, iwlen;
  int *Pe, *Nv, *Head, *Elen, *Degree, *W, *Len, *Next, *Last, *Iw;
};

// phase timers of scripts/amd_prof.cu (compiled out of the product)
#ifdef PB_AMD_PROF
__device__ long long g_amd_prof[16];
#define PB_T0() long long t_last_ = clock64()
#define PB_T0R() t_last_ = clock64()
#define PB_T(k) do { const long long t_ = clock64(); if (thread_id() == 0) g_amd_prof[k] += t_ - t_last_; t_last_ = t_; } while (0)
#define PB_COUNT(k, v) do { if (thread_id() == 0) g_amd_prof[k] += (v); } while (0)
#else
#define PB_T0() do {} while (0)
#define PB_T0R() do {} while (0)
#define PB_T(k) do {} while (0)
#define PB_COUNT(k, v) do {} while (0)
#endif

// ------------------------------------------------------------------ A + A' (amd_aat / amd_1)
// For a sorted, duplicate-free, structurally symmetric pattern without diagonal (what the
// extraction kernels emit) amd_1's sweep appends to the list of v first every j < v (when column v
// is visited) and then every k > v in ascending order: the list IS the stored row.  Checked here
// entry by entry (binary search of the mirror); anything else goes through build_general below.
PB_DEVFN bool rows_are_symmetric(int n, const int *Cp, int e0, const int *Ci, int *cta_flag) {
  const int tid = thread_id(), nt = num_threads();
  if (tid == 0) *cta_flag = 0;
  sync_cta();
  int bad = 0;
  for (int k = tid; k < n; k += nt) {
    int prev = -1;
    for (int p = Cp[k] - e0; p < Cp[k + 1] - e0; p++) {
      const int j = Ci[p];
      if (j <= prev || j == k || j < 0 || j >= n) { bad = 1; break; }
      prev = j;
      int lo = Cp[j] - e0, hi = Cp[j + 1] - e0 - 1, found = 0;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1, v = Ci[mid];
        if (v == k) { found = 1; break; }
        if (v < k) lo = mid + 1; else hi = mid - 1;
      }
      if (!found) { bad = 1; break; }
    }
  }
  if (bad) atomic_add(cta_flag, 1);
  sync_cta();
  const bool ok = *cta_flag == 0;
  sync_cta();
  return ok;
}

PB_DEVFN int build_symmetric(const Ws &w, const int *Cp, int e0, const int *Ci) {
  const int tid = thread_id(), nt = num_threads(), n = w.n;
  for (int k = tid; k < n; k += nt) {
    w.Len[k] = Cp[k + 1] - Cp[k];
    w.Pe[k] = Cp[k] - e0;
  }
  const int nz = Cp[n] - e0;
  for (int p = tid; p < nz; p += nt) w.Iw[p] = Ci[p];
  sync_cta();
  return nz;
}

// general input (one thread; never taken by the product path, kept for amd_order's full contract)
PB_DEVFN long long aat_count_serial(int n, const int *Cp, int e0, const int *Ci, int *Len, int *Tp) {
  for (int k = 0; k < n; k++) Len[k] = 0;
  for (int k = 0; k < n; k++) {
    const int p1 = Cp[k] - e0, p2 = Cp[k + 1] - e0;
    int p;
    for (p = p1; p < p2;) {
      const int j = Ci[p];
      if (j < k) { Len[j]++; Len[k]++; p++; }
      else if (j == k) { p++; break; }
      else break;
      const int pj2 = Cp[j + 1] - e0;
      int pj;
      for (pj = Tp[j]; pj < pj2;) {
        const int i = Ci[pj];
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
  return nzaat;
}

PB_DEVFN int build_general_serial(const Ws &w, const int *Cp, int e0, const int *Ci, int *Sp, int *Tp) {
  const int n = w.n;
  int pfree = 0;
  for (int j = 0; j < n; j++) { w.Pe[j] = pfree; Sp[j] = pfree; pfree += w.Len[j]; }
  for (int k = 0; k < n; k++) {
    const int p1 = Cp[k] - e0, p2 = Cp[k + 1] - e0;
    int p;
    for (p = p1; p < p2;) {
      const int j = Ci[p];
      if (j < k) { w.Iw[Sp[j]++] = k; w.Iw[Sp[k]++] = j; p++; }
      else if (j == k) { p++; break; }
      else break;
      const int pj2 = Cp[j + 1] - e0;
      int pj;
      for (pj = Tp[j]; pj < pj2;) {
        const int i = Ci[pj];
        if (i < k) { w.Iw[Sp[i]++] = j; w.Iw[Sp[j]++] = i; pj++; }
        else if (i == k) { pj++; break; }
        else break;
      }
      Tp[j] = pj;
    }
    Tp[k] = p;
  }
  for (int j = 0; j < n; j++)
    for (int pj = Tp[j]; pj < Cp[j + 1] - e0; pj++) {
      const int i = Ci[pj];
      w.Iw[Sp[i]++] = j; w.Iw[Sp[j]++] = i;
    }
  return pfree;
}

// ------------------------------------------------------------------ degree lists
// Pushes the active lanes' variables at the heads of their degree lists in lane order (what a
// sequential loop over the chunk does): lanes of one degree chain to each other, the lowest one
// links to the old head, the highest one becomes the head.
PB_DEV void deglist_push_chunk(const Ws &w, int i, int deg, bool act) {
  const int lane = lane_id();
  const unsigned grp = match_any(act ? deg : NONE);
  int old = NONE;
  if (act) old = w.Head[deg];
  const unsigned lower = grp & lanemask_lt(), higher = grp & lanemask_gt();
  const int pred = (act && lower) ? highest(lower) : -1;
  const int succ = (act && higher) ? lowest(higher) : -1;
  const int i_pred = shfl(i, pred >= 0 ? pred : lane);
  const int i_succ = shfl(i, succ >= 0 ? succ : lane);
  sync_warp();  // every old head is read before any list is written
  if (act) {
    w.Next[i] = pred >= 0 ? i_pred : old;
    w.Last[i] = succ >= 0 ? i_succ : NONE;
    if (pred < 0 && old != NONE) w.Last[old] = i;
    if (succ < 0) w.Head[deg] = i;
  }
  sync_warp();
}

// Unlinks the kept lanes' variables (their Nv is already negated, which is the mark "leaving")
// from their degree lists.  Variables that are list neighbours leave as one run: the lane at the
// front of a run walks to the first staying successor and splices the whole run out.
PB_DEV void deglist_remove_chunk(const Ws &w, int i, bool keep) {
  sync_warp();  // the Nv marks of this chunk are visible
  if (keep) {
    const int ilast = w.Last[i];
    if (ilast == NONE || w.Nv[ilast] > 0) {
      int s = w.Next[i], guard = 0;
      while (s != NONE && w.Nv[s] < 0 && ++guard <= WARP) s = w.Next[s];
      // written: fields of staying variables only; read above: fields of leaving ones only
      if (s != NONE) w.Last[s] = ilast;
      if (ilast != NONE) w.Next[ilast] = s; else w.Head[w.Degree[i]] = s;
    }
  }
  sync_warp();
}

PB_DEV int reset_stamp_warp(int wflg, int wbig, int *W, int n) {
  if (wflg < 2 || wflg >= wbig) {
    for (int x = lane_id(); x < n; x += WARP) if (W[x] != 0) W[x] = 1;
    sync_warp();
    wflg = 2;
  }
  return wflg;
}

// moves a block of Iw down (dst <= src), 32 entries at a time in ascending order
PB_DEV void move_down(int *Iw, int dst, int src, int len) {
  const int lane = lane_id();
  for (int k = 0; k < len; k += WARP) {
    int v = 0;
    if (k + lane < len) v = Iw[src + k + lane];
    sync_warp();
    if (k + lane < len) Iw[dst + k + lane] = v;
    sync_warp();
  }
}

// Workspace compaction (the "garbage collection" of amd_2): live lists keep their order and their
// relative position; the part of Lme built so far ([pme1, pfree)) moves behind them.
PB_DEV void compact_workspace(const Ws &w, int &pme1, int &pfree) {
  const int lane = lane_id(), n = w.n;
  for (int j = lane; j < n; j += WARP) {
    const int pn = w.Pe[j];
    if (pn >= 0) { w.Pe[j] = w.Iw[pn]; w.Iw[pn] = flip(j); }
  }
  sync_warp();
  int psrc = 0, pdst = 0;
  while (psrc < pme1) {
    int v = 0;
    if (psrc + lane < pme1) v = w.Iw[psrc + lane];
    const unsigned b = ballot(v <= -2);
    if (!b) { psrc += WARP; continue; }
    const int first = lowest(b);
    const int j = flip(shfl(v, first));
    psrc += first;
    const int lenj = w.Len[j];
    sync_warp();
    if (lane == 0) { w.Iw[pdst] = w.Pe[j]; w.Pe[j] = pdst; }
    sync_warp();
    move_down(w.Iw, pdst + 1, psrc + 1, lenj - 1);
    pdst += lenj;
    psrc += lenj;
  }
  const int p1 = pdst, lme = pfree - pme1;
  move_down(w.Iw, pdst, pme1, lme);
  pme1 = p1;
  pfree = p1 + lme;
}

// ------------------------------------------------------------------ initial state
// All threads of the CTA.  *nel_shared receives the number of variables eliminated up front
// (empty rows, dense rows).
PB_DEVFN void init_state(const Ws &w, int *nel_shared) {
  const int tid = thread_id(), nt = num_threads(), n = w.n;
  int dense = (int)(10.0 * sqrt((double)n));
  if (dense < 16) dense = 16;
  if (dense > n) dense = n;
  if (tid == 0) *nel_shared = 0;
  for (int i = tid; i < n; i += nt) {
    w.Last[i] = NONE; w.Head[i] = NONE; w.Next[i] = NONE;
    w.Nv[i] = 1; w.W[i] = 1; w.Elen[i] = 0; w.Degree[i] = w.Len[i];
  }
  sync_cta();
  int cnt = 0;
  for (int i = tid; i < n; i += nt) {
    const int deg = w.Degree[i];
    if (deg == 0) { w.Elen[i] = flip(1); cnt++; w.Pe[i] = NONE; w.W[i] = 0; }
    else if (deg > dense) { w.Nv[i] = 0; w.Elen[i] = NONE; cnt++; w.Pe[i] = NONE; }
  }
  if (cnt) atomic_add(nel_shared, cnt);
  sync_cta();
  if (tid < WARP) {
    const int lane = lane_id();
    for (int base = 0; base < n; base += WARP) {
      const int i = base + lane;
      int deg = 0;
      if (i < n) deg = w.Degree[i];
      deglist_push_chunk(w, i, deg, i < n && deg > 0 && deg <= dense);
    }
  }
  sync_cta();
}

// ------------------------------------------------------------------ elimination (ONE warp)
// per-lane pieces of the phases; `valid` lanes own one variable i of Lme

// P2a: an element touched for the first time in this pivot gets W[e] = Degree[e] + wflg
PB_DEV void p2_seed(const Ws &w, bool valid, int p1, int eln, int wflg) {
  if (!valid) return;
  for (int p = p1; p < p1 + eln; p += 4) {
    int e[4], we[4], dg[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e[k] = p + k < p1 + eln ? w.Iw[p + k] : NONE;
#pragma unroll
    for (int k = 0; k < 4; k++) { we[k] = 0; dg[k] = 0; if (e[k] != NONE) { we[k] = w.W[e[k]]; dg[k] = w.Degree[e[k]]; } }
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (e[k] != NONE && we[k] != 0 && we[k] < wflg) w.W[e[k]] = dg[k] + wflg;  // same value from every writer
  }
}

// P2b: ... minus nv_i for every variable i of Lme that e is adjacent to
PB_DEV void p2_subtract(const Ws &w, bool valid, int p1, int eln, int nvi) {
  if (!valid) return;
  for (int p = p1; p < p1 + eln; p += 4) {
    int e[4], we[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e[k] = p + k < p1 + eln ? w.Iw[p + k] : NONE;
#pragma unroll
    for (int k = 0; k < 4; k++) we[k] = e[k] != NONE ? w.W[e[k]] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++) if (we[k] != 0) atomic_add(&w.W[e[k]], -nvi);
  }
}

// P3: compaction of i's lists, approximate external degree, hash.  Returns true if i stays alive
// (false: mass-eliminated, its weight is returned in massnv).
PB_DEV bool p3_degree(const Ws &w, int i, int p1, int eln, int ln, int nvi, int me, int wflg, int &massnv, int &hash_out) {
  int *const Iw = w.Iw;
  const int p2 = p1 + eln;  // one past the elements
  int pn = p1, d = 0, first_e = NONE, first_v = NONE;
  unsigned int hash = 0;
  for (int p = p1; p < p2; p += 4) {
    int e[4], we[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e[k] = p + k < p2 ? Iw[p + k] : NONE;
#pragma unroll
    for (int k = 0; k < 4; k++) we[k] = e[k] != NONE ? w.W[e[k]] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (we[k] != 0) {
        const int dext = we[k] - wflg;
        if (dext > 0) {
          d += dext;
          if (first_e == NONE) first_e = e[k];
          Iw[pn++] = e[k];
          hash += (unsigned int)e[k];
        } else {
          w.Pe[e[k]] = flip(me);  // aggressive absorption (idempotent across lanes)
          w.W[e[k]] = 0;
        }
      }
  }
  const int p3 = pn, p4 = p1 + ln;
  for (int p = p2; p < p4; p += 4) {
    int j[4], nvj[4];
#pragma unroll
    for (int k = 0; k < 4; k++) j[k] = p + k < p4 ? Iw[p + k] : NONE;
#pragma unroll
    for (int k = 0; k < 4; k++) nvj[k] = j[k] != NONE ? w.Nv[j[k]] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (nvj[k] > 0) {
        d += nvj[k];
        if (first_v == NONE) first_v = j[k];
        Iw[pn++] = j[k];
        hash += (unsigned int)j[k];
      }
  }
  if (p3 == p1 && pn == p3) {
    // mass elimination: only the element me is left
    w.Pe[i] = flip(me);
    massnv = nvi;
    w.Nv[i] = 0;
    w.Elen[i] = NONE;
    return false;
  }
  w.Elen[i] = p3 - p1 + 1;
  if (d < w.Degree[i]) w.Degree[i] = d;
  // me goes first, the first element moves behind the elements, the first variable to the end
  if (pn > p3) Iw[pn] = first_v;
  if (p3 > p1) Iw[p3] = first_e;
  Iw[p1] = me;
  w.Len[i] = pn - p1 + 1;
  hash_out = (int)(hash % (unsigned int)w.n);
  w.Last[i] = hash_out;
  return true;
}

// hash buckets, chained in lane order.  Bucket head of hash h: flipped in Head[h] when degree
// list h is empty, else in Last[] of that list's head variable.
PB_DEV void hash_push_chunk(const Ws &w, int i, int hash_i, bool alive) {
  const int lane = lane_id();
  const unsigned grp = match_any(alive ? hash_i : NONE);
  int hj = NONE;
  if (alive) hj = w.Head[hash_i];
  const unsigned lower = grp & lanemask_lt(), higher = grp & lanemask_gt();
  const int pred = (alive && lower) ? highest(lower) : -1;
  const int i_pred = shfl(i, pred >= 0 ? pred : lane);
  int oldb = NONE;
  if (alive && pred < 0) oldb = hj <= NONE ? flip(hj) : w.Last[hj];
  sync_warp();
  if (alive) {
    w.Next[i] = pred >= 0 ? i_pred : oldb;
    if (!higher) { if (hj <= NONE) w.Head[hash_i] = flip(i); else w.Last[hj] = i; }
  }
  sync_warp();
}

// The whole element fits one chunk: the buckets are empty before (they are emptied at the end of
// every pivot), so a hash that only one lane holds is a bucket of one -- nothing to compare, nothing
// to store, nothing to clear.  Only hashes shared by several lanes are chained; returns the lanes
// that have company (P4 walks those buckets).
PB_DEV unsigned hash_push_single_chunk(const Ws &w, int i, int hash_i, bool alive) {
  const int lane = lane_id();
  const unsigned grp = match_any(alive ? hash_i : NONE);
  const bool shared = alive && (grp & (grp - 1)) != 0;
  const unsigned any_shared = ballot(shared);
  if (!any_shared) return 0u;
  int hj = NONE;
  if (shared) hj = w.Head[hash_i];
  const unsigned lower = grp & lanemask_lt(), higher = grp & lanemask_gt();
  const int pred = (shared && lower) ? highest(lower) : -1;
  const int i_pred = shfl(i, pred >= 0 ? pred : lane);
  int oldb = NONE;
  if (shared && pred < 0) oldb = hj <= NONE ? flip(hj) : w.Last[hj];
  sync_warp();
  if (shared) {
    w.Next[i] = pred >= 0 ? i_pred : oldb;
    if (!higher) { if (hj <= NONE) w.Head[hash_i] = flip(i); else w.Last[hj] = i; }
  }
  sync_warp();
  return any_shared;
}

// P4 for one chunk: buckets with one member are emptied in parallel; the others are walked one
// after the other in lane (= Lme) order, marking and comparing lists with all lanes
PB_DEV void p4_supervariables(const Ws &w, int i, bool alive, int hash_i, int &wflg, bool known, unsigned known_mm) {
  const int lane = lane_id(), n = w.n;
  int *const Pe = w.Pe, *const Nv = w.Nv, *const Head = w.Head, *const Elen = w.Elen;
  int *const W = w.W, *const Len = w.Len, *const Next = w.Next, *const Last = w.Last, *const Iw = w.Iw;
  unsigned mm;
  if (known) {
    // single-chunk mode: only buckets shared by several lanes exist (hash_push_single_chunk)
    mm = known_mm;
    if (!mm) return;
  } else {
    bool multi = false;
    int j0 = NONE;
    if (alive) {
      j0 = Head[hash_i];
      const int bh = j0 == NONE ? NONE : (j0 < NONE ? flip(j0) : Last[j0]);
      multi = bh != NONE && Next[bh] != NONE;
    }
    mm = ballot(multi);
    sync_warp();
    if (alive && !multi) { if (j0 < NONE) Head[hash_i] = NONE; else if (j0 > NONE) Last[j0] = NONE; }
    if (!mm) return;
    sync_warp();
  }
  while (mm) {
    const int src = lowest(mm);
    mm &= mm - 1;
    const int h = shfl(hash_i, src);
    const int j = Head[h];
    int cur;
    if (j == NONE) cur = NONE;
    else if (j < NONE) cur = flip(j);
    else cur = Last[j];
    sync_warp();
    if (lane == 0) { if (j < NONE) Head[h] = NONE; else if (j > NONE) Last[j] = NONE; }
    sync_warp();
    int guard = 0;
    while (cur != NONE && Next[cur] != NONE && ++guard <= n) {
      const int ln = Len[cur], eln = Elen[cur], pc = Pe[cur];
      for (int p = pc + 1 + lane; p <= pc + ln - 1; p += WARP) W[Iw[p]] = wflg;
      sync_warp();
      int jlast = cur, jj = Next[cur], guard2 = 0;
      while (jj != NONE && ++guard2 <= n) {
        bool ok = Len[jj] == ln && Elen[jj] == eln;
        if (ok) {
          const int pjj = Pe[jj];
          bool bad = false;
          for (int p = pjj + 1 + lane; p <= pjj + ln - 1; p += WARP)
            if (W[Iw[p]] != wflg) bad = true;
          ok = ballot(bad) == 0;
        }
        const int jn = Next[jj];
        sync_warp();
        if (ok) {
          if (lane == 0) {
            Pe[jj] = flip(cur);
            Nv[cur] += Nv[jj];
            Nv[jj] = 0;
            Elen[jj] = NONE;
            Next[jlast] = jn;
          }
        } else {
          jlast = jj;
        }
        jj = jn;
        sync_warp();
      }
      wflg++;
      cur = Next[cur];
    }
  }
}

// the flattened read of P1: lane q < nlists describes list q (start pe_k, length len_k, exclusive
// prefix excl_k); returns the Iw position of stream entry t, or -1
PB_DEV int stream_source(int t, int nlists, int excl_k, int len_k, int pe_k, int base, int &q0) {
  int src = -1;
  for (int q = q0; q < nlists; q++) {
    const int sq = shfl(excl_k, q), lq = shfl(len_k, q), pq = shfl(pe_k, q);
    if (sq >= base + WARP) break;           // uniform: the rest starts behind this chunk
    if (sq + lq <= base) { q0 = q + 1; continue; }  // uniform: entirely before this chunk
    if (t >= sq && t < sq + lq) src = pq + (t - sq);
  }
  return src;
}

// one chunk of the flattened P1 stream: src = Iw position of this lane's entry (or -1)
PB_DEV void p1_take_chunk(const Ws &w, int src, int &pfree, int &degme) {
  const int lane = lane_id();
  int i = 0, nvi = 0;
  if (src >= 0) { i = w.Iw[src]; nvi = w.Nv[i]; }
  const bool cand = nvi > 0;
  const unsigned grp = match_any(cand ? i : NONE);
  const bool keep = cand && (grp & lanemask_lt()) == 0;
  const unsigned b = ballot(keep);
  if (keep) { w.Nv[i] = -nvi; w.Iw[pfree + popc(b & lanemask_lt())] = i; }
  degme += warp_sum(keep ? nvi : 0);
  pfree += popc(b);
  deglist_remove_chunk(w, i, keep);
}

PB_DEVFN void eliminate_warp(const Ws &w, int pfree, int nel) {
  const int lane = lane_id(), n = w.n, iwlen = w.iwlen;
  int *const Pe = w.Pe, *const Nv = w.Nv, *const Head = w.Head, *const Elen = w.Elen;
  int *const Degree = w.Degree, *const W = w.W, *const Len = w.Len, *const Next = w.Next;
  int *const Last = w.Last, *const Iw = w.Iw;
  int mindeg = 0, lemax = 0;
  const int wbig = INT_MAX - n;
  int wflg = 2;  // W is all 0 / 1 after init_state

  PB_T0();
  while (nel < n) {
    // ---- P0: the head of the first non-empty degree list
    int me = NONE, deg = mindeg;
    for (;;) {
      const int d = deg + lane;
      const int hd = d < n ? Head[d] : NONE;
      const unsigned b = ballot(hd != NONE);
      if (b) { const int src = lowest(b); me = shfl(hd, src); deg += src; break; }
      deg += WARP;
      if (deg >= n) break;  // cannot happen while nel < n
    }
    if (me == NONE) break;
    mindeg = deg;
    const int elenme = Elen[me];
    int nvpiv = Nv[me];
    const int me_next = Next[me];
    const int pe_me = Pe[me], len_me = Len[me];
    sync_warp();
    if (lane == 0) {
      if (me_next != NONE) Last[me_next] = NONE;
      Head[deg] = me_next;
      Nv[me] = -nvpiv;
    }
    nel += nvpiv;
    int degme = 0, pme1, pme2;
    sync_warp();
    PB_T(0);

    // ---- P1: the new element Lme
    if (elenme == 0) {
      pme1 = pe_me;
      pme2 = pme1 - 1;
      for (int base = 0; base < len_me; base += WARP) {
        const bool valid = base + lane < len_me;
        int i = 0, nvi = 0;
        if (valid) { i = Iw[pme1 + base + lane]; nvi = Nv[i]; }
        const bool keep = nvi > 0;
        const unsigned b = ballot(keep);
        if (keep) { Nv[i] = -nvi; Iw[pme2 + 1 + popc(b & lanemask_lt())] = i; }
        degme += warp_sum(keep ? nvi : 0);
        pme2 += popc(b);
        deglist_remove_chunk(w, i, keep);
      }
    } else {
      int p = pe_me;
      pme1 = pfree;
      const int slenme = len_me - elenme;
      const int nlists = elenme + 1;
      bool flat = nlists <= WARP, done = false;
      int e_k = NONE, pe_k = 0, len_k = 0, incl = 0, total = 0;
      if (elenme == 1) {
        // the common case (one adjacent element): both list descriptors are read by every lane,
        // no scan and no shuffles to find a lane's source
        const int e0 = Iw[p];
        const int pe0 = Pe[e0], len0 = Len[e0];
        const int tot = len0 + slenme;
        if ((long long)pfree + tot <= (long long)iwlen) {
          for (int base = 0; base < tot; base += WARP) {
            const int t = base + lane;
            p1_take_chunk(w, t < len0 ? pe0 + t : (t < tot ? p + 1 + (t - len0) : -1), pfree, degme);
          }
          if (lane == 0) { Pe[e0] = flip(me); W[e0] = 0; }
          done = true;
        }
      }
      if (!done && flat) {
        if (lane < elenme) { e_k = Iw[p + lane]; pe_k = Pe[e_k]; len_k = Len[e_k]; }
        else if (lane == elenme) { pe_k = p + elenme; len_k = slenme; }
        incl = warp_incl_scan(len_k);
        total = shfl(incl, WARP - 1);
        flat = (long long)pfree + total <= (long long)iwlen;
      }
      if (done) {
        // Lme is complete
      } else if (flat) {
        // every feeding list at once: the concatenation of the lists is read 32 entries at a
        // time; a variable met twice keeps its first place (lowest lane, or an earlier chunk)
        const int excl = incl - len_k;
        int q0 = 0;
        for (int base = 0; base < total; base += WARP)
          p1_take_chunk(w, stream_source(base + lane, nlists, excl, len_k, pe_k, base, q0), pfree, degme);
        if (lane < elenme) { Pe[e_k] = flip(me); W[e_k] = 0; }
      } else {
        // list by list, with workspace compaction when the room runs out
        for (int knt1 = 1; knt1 <= elenme + 1; knt1++) {
          int e, pj, ln;
          if (knt1 > elenme) { e = me; pj = p; ln = slenme; }
          else { e = Iw[p++]; pj = Pe[e]; ln = Len[e]; }
          while (ln > 0) {
            const int chunk = ln < WARP ? ln : WARP;
            const bool valid = lane < chunk;
            int i = 0, nvi = 0;
            if (valid) { i = Iw[pj + lane]; nvi = Nv[i]; }
            const bool keep = nvi > 0;
            const unsigned b = ballot(keep);
            const int cnt = popc(b);
            if (pfree + cnt > iwlen) {
              // compact, with the consumed parts of me's and e's lists trimmed away
              sync_warp();
              if (lane == 0) {
                Pe[me] = p;
                Len[me] -= knt1;
                if (Len[me] == 0) Pe[me] = NONE;
                Pe[e] = pj;
                Len[e] = ln;
              }
              sync_warp();
              compact_workspace(w, pme1, pfree);
              sync_warp();
              pj = Pe[e];
              p = Pe[me];
              continue;  // re-read the chunk at its new place (room is guaranteed now)
            }
            if (keep) { Nv[i] = -nvi; Iw[pfree + popc(b & lanemask_lt())] = i; }
            degme += warp_sum(keep ? nvi : 0);
            pfree += cnt;
            deglist_remove_chunk(w, i, keep);
            pj += chunk;
            ln -= chunk;
          }
          if (e != me && lane == 0) { Pe[e] = flip(me); W[e] = 0; }
        }
      }
      pme2 = pfree - 1;
    }
    sync_warp();
    if (lane == 0) {
      Degree[me] = degme;
      Pe[me] = pme1;
      Len[me] = pme2 - pme1 + 1;
      Elen[me] = flip(nvpiv + degme);
    }
    wflg = reset_stamp_warp(wflg, wbig, W, n);
    sync_warp();
    const int lme = pme2 - pme1 + 1;
    PB_T(1);
    PB_COUNT(8, 1);
    PB_COUNT(9, lme);
    PB_COUNT(10, elenme);

    int pw = pme1;
    if (lme <= WARP) {
      // ---- P2..P5 with every variable of Lme held by one lane for the whole pivot
      const bool valid = lane < lme;
      int i = 0, p1 = 0, eln = 0, ln = 0, nvi = 0;
      if (valid) { i = Iw[pme1 + lane]; p1 = Pe[i]; eln = Elen[i]; ln = Len[i]; nvi = -Nv[i]; }
      p2_seed(w, valid, p1, eln, wflg);
      sync_warp();
      p2_subtract(w, valid, p1, eln, nvi);
      sync_warp();
      PB_T(2);
      int massnv = 0, hash_i = 0;
      bool alive = false;
      if (valid) alive = p3_degree(w, i, p1, eln, ln, nvi, me, wflg, massnv, hash_i);
      const int dm = warp_sum(massnv);
      degme -= dm;
      nvpiv += dm;
      nel += dm;
      const unsigned shared_mm = hash_push_single_chunk(w, i, hash_i, alive);
      if (degme > lemax) lemax = degme;
      wflg += lemax;
      wflg = reset_stamp_warp(wflg, wbig, W, n);
      sync_warp();
      PB_T(3);
      p4_supervariables(w, i, alive, hash_i, wflg, true, shared_mm);
      sync_warp();
      PB_T(4);
      const int nleft = n - nel;
      int d = 0;
      nvi = alive ? -Nv[i] : 0;  // merged weight, 0 if absorbed by P4
      alive = nvi > 0;
      if (alive) {
        Nv[i] = nvi;
        d = Degree[i] + degme - nvi;
        if (d > nleft - nvi) d = nleft - nvi;
        Degree[i] = d;
      }
      const unsigned b = ballot(alive);
      if (alive) Iw[pw + popc(b & lanemask_lt())] = i;
      pw += popc(b);
      const int dmin = warp_min(alive ? d : INT_MAX);
      if (dmin < mindeg) mindeg = dmin;
      deglist_push_chunk(w, i, d, alive);
    } else {
      // ---- the same phases, chunk by chunk
      for (int base = 0; base < lme; base += WARP) {
        const bool valid = base + lane < lme;
        int p1 = 0, eln = 0;
        if (valid) { const int i = Iw[pme1 + base + lane]; p1 = Pe[i]; eln = Elen[i]; }
        p2_seed(w, valid, p1, eln, wflg);
      }
      sync_warp();
      for (int base = 0; base < lme; base += WARP) {
        const bool valid = base + lane < lme;
        int p1 = 0, eln = 0, nvi = 0;
        if (valid) { const int i = Iw[pme1 + base + lane]; p1 = Pe[i]; eln = Elen[i]; nvi = -Nv[i]; }
        p2_subtract(w, valid, p1, eln, nvi);
      }
      sync_warp();
      PB_T(2);
      for (int base = 0; base < lme; base += WARP) {
        const bool valid = base + lane < lme;
        int i = 0, massnv = 0, hash_i = 0;
        bool alive = false;
        if (valid) {
          i = Iw[pme1 + base + lane];
          alive = p3_degree(w, i, Pe[i], Elen[i], Len[i], -Nv[i], me, wflg, massnv, hash_i);
        }
        const int dm = warp_sum(massnv);
        degme -= dm;
        nvpiv += dm;
        nel += dm;
        hash_push_chunk(w, i, hash_i, alive);
      }
      if (degme > lemax) lemax = degme;
      wflg += lemax;
      wflg = reset_stamp_warp(wflg, wbig, W, n);
      sync_warp();
      PB_T(3);
      for (int base = 0; base < lme; base += WARP) {
        const bool valid = base + lane < lme;
        int i = 0, hash_i = 0;
        bool alive = false;
        if (valid) {
          i = Iw[pme1 + base + lane];
          alive = Nv[i] < 0;
          if (alive) hash_i = Last[i];
        }
        p4_supervariables(w, i, alive, hash_i, wflg, false, 0u);
        sync_warp();
      }
      PB_T(4);
      const int nleft = n - nel;
      for (int base = 0; base < lme; base += WARP) {
        const bool valid = base + lane < lme;
        int i = 0, nvi = 0, d = 0;
        if (valid) { i = Iw[pme1 + base + lane]; nvi = -Nv[i]; }
        const bool alive = nvi > 0;
        if (alive) {
          Nv[i] = nvi;
          d = Degree[i] + degme - nvi;
          if (d > nleft - nvi) d = nleft - nvi;
          Degree[i] = d;
        }
        const unsigned b = ballot(alive);
        if (alive) Iw[pw + popc(b & lanemask_lt())] = i;
        pw += popc(b);
        const int dmin = warp_min(alive ? d : INT_MAX);
        if (dmin < mindeg) mindeg = dmin;
        deglist_push_chunk(w, i, d, alive);
      }
    }
    if (lane == 0) {
      Degree[me] = degme;
      Nv[me] = nvpiv;
      Len[me] = pw - pme1;
      if (pw == pme1) { Pe[me] = NONE; W[me] = 0; }
    }
    if (elenme != 0) pfree = pw;
    sync_warp();
    PB_T(5);
  }
}

// ------------------------------------------------------------------ assembly-tree post-order
// amd_post_tree: depth-first numbering of one tree of the assembly forest (one thread; the rest of
// amd_2's epilogue and of amd_postorder runs on the whole CTA, finish_cta below)
PB_DEVFN int post_tree_serial(int root, int k, int *Child, const int *Sibling, int *Order, int *Stack) {
  int head = 0;
  Stack[0] = root;
  while (head >= 0) {
    const int i = Stack[head];
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

// ------------------------------------------------------------------ epilogue, whole CTA
// amd_2's epilogue + amd_postorder.  Per-vertex passes run on all threads; the steps whose outcome
// depends on the visiting order (children lists in ascending id, slots of absorbed variables in
// ascending id) advance one warp-wide chunk at a time on warp 0; only the depth-first numbering of
// the assembly tree (one node per pivot) stays on one thread.  sh: >= 4 + 32 ints shared by the CTA.
PB_DEVFN void finish_cta(const Ws &w, int *sh) {
  const int n = w.n, tid = thread_id(), nt = num_threads(), lane = lane_id();
  int *const Pe = w.Pe, *const Nv = w.Nv, *const Head = w.Head, *const Elen = w.Elen;
  int *const W = w.W, *const Next = w.Next, *const Last = w.Last, *const Root = w.Degree, *const Tail = w.Len;
  for (int i = tid; i < n; i += nt) { Pe[i] = flip(Pe[i]); Elen[i] = flip(Elen[i]); }
  sync_cta();
  // absorbed variables point at the element that finally holds them
  for (int i = tid; i < n; i += nt) {
    int r = NONE;
    if (Nv[i] == 0) {
      r = Pe[i];
      int guard = 0;
      while (r != NONE && Nv[r] == 0 && ++guard <= n) r = Pe[r];
    }
    Root[i] = r;
  }
  sync_cta();
  int *const Child = Head, *const Sibling = Next, *const Order = W, *const Stack = Last;
  for (int i = tid; i < n; i += nt) {
    if (Nv[i] == 0) Pe[i] = Root[i];
    Child[i] = NONE; Sibling[i] = NONE; Tail[i] = NONE; Order[i] = NONE;
  }
  sync_cta();
  // children of every element in ascending id (amd_postorder pushes them in descending order)
  if (tid < WARP) {
    for (int base = 0; base < n; base += WARP) {
      const int j = base + lane;
      int par = NONE;
      bool act = false;
      if (j < n && Nv[j] > 0) { par = Pe[j]; act = par != NONE; }
      const unsigned grp = match_any(act ? par : NONE);
      const unsigned lower = grp & lanemask_lt(), higher = grp & lanemask_gt();
      const int succ = (act && higher) ? lowest(higher) : -1;
      const int j_succ = shfl(j, succ >= 0 ? succ : lane);
      int oldtail = NONE;
      if (act && !lower) oldtail = Tail[par];
      sync_warp();
      if (act) {
        if (!lower) { if (oldtail == NONE) Child[par] = j; else Sibling[oldtail] = j; }
        if (succ >= 0) Sibling[j] = j_succ;
        if (!higher) Tail[par] = j;
      }
      sync_warp();
    }
  }
  sync_cta();
  // the largest child (last one among equals) goes to the end of its list
  for (int i = tid; i < n; i += nt)
    if (Nv[i] > 0 && Child[i] != NONE) {
      int fprev = NONE, maxfr = NONE, bigfprev = NONE, bigf = NONE;
      for (int f = Child[i]; f != NONE; f = Sibling[f]) {
        const int fr = Elen[f];
        if (fr >= maxfr) { maxfr = fr; bigfprev = fprev; bigf = f; }
        fprev = f;
      }
      const int fnext = Sibling[bigf];
      if (fnext != NONE) {
        if (bigfprev == NONE) Child[i] = fnext; else Sibling[bigfprev] = fnext;
        Sibling[bigf] = NONE;
        Sibling[fprev] = bigf;
      }
    }
  sync_cta();
  // post-order of the forest, roots in ascending id
  if (tid < WARP) {
    int k = 0;
    for (int base = 0; base < n; base += WARP) {
      const int i = base + lane;
      unsigned b = ballot(i < n && Nv[i] > 0 && Pe[i] == NONE);
      while (b) {
        const int r = base + lowest(b);
        b &= b - 1;
        if (lane == 0) k = post_tree_serial(r, k, Child, Sibling, Order, Stack);
        k = shfl(k, 0);
      }
    }
  }
  sync_cta();
  // slots: elements in post-order, each behind the variables it absorbed (ascending id)
  for (int q = tid; q < n; q += nt) { Head[q] = NONE; Next[q] = NONE; }
  sync_cta();
  for (int e = tid; e < n; e += nt) { const int q = Order[e]; if (q != NONE) Head[q] = e; }
  sync_cta();
  const int nwarps = (nt + WARP - 1) / WARP, warp = tid / WARP;
  int running = 0;
  for (int base = 0; base < n; base += nt) {
    const int q = base + tid;
    int e = NONE, v = 0;
    if (q < n) { e = Head[q]; if (e != NONE) v = Nv[e]; }
    const int incl = warp_incl_scan(v);
    if (lane == WARP - 1) sh[4 + warp] = incl;
    sync_cta();
    int off = 0, tot = 0;
    for (int x = 0; x < nwarps; x++) { const int sx = sh[4 + x]; if (x < warp) off += sx; tot += sx; }
    if (e != NONE) Next[e] = running + off + incl - v;
    running += tot;
    sync_cta();
  }
  if (tid < WARP) {
    int nel = running;
    for (int base = 0; base < n; base += WARP) {
      const int i = base + lane;
      int e = NONE;
      bool absorbed = false, dense = false;
      if (i < n && Nv[i] == 0) { e = Pe[i]; absorbed = e != NONE; dense = e == NONE; }
      const unsigned grp = match_any(absorbed ? e : NONE);
      int slot = 0;
      if (absorbed) slot = Next[e];
      sync_warp();
      if (absorbed) {
        Next[i] = slot + popc(grp & lanemask_lt());
        if ((grp & lanemask_lt()) == 0) Next[e] = slot + popc(grp);
      }
      const unsigned bd = ballot(dense);
      if (dense) Next[i] = nel + popc(bd & lanemask_lt());
      nel += popc(bd);
      sync_warp();
    }
  }
  sync_cta();
  for (int i = tid; i < n; i += nt) Last[Next[i]] = i;
  sync_cta();
}

// ------------------------------------------------------------------ one graph, one CTA
// slab: 8n + iwlen ints (Len, Next, Pe, Nv, Head, Elen, Degree, W, Iw); P: n ints, receives the
// permutation (P[k] = k-th pivot); sh: 4 + 32 ints shared by the CTA.  iwlen >= nzaat + n is the
// caller's job (nzaat = nz for symmetric rows, <= 2 nz otherwise).
PB_DEVFN void order_graph(int n, const int *Cp, int e0, const int *Ci, bool symmetric, int *slab,
                          long long slab_ints, int *P, int *sh) {
  Ws w;
  w.n = n;
  w.Len = slab;
  w.Next = slab + n;
  w.Pe = slab + 2 * (size_t)n;
  w.Nv = slab + 3 * (size_t)n;
  w.Head = slab + 4 * (size_t)n;
  w.Elen = slab + 5 * (size_t)n;
  w.Degree = slab + 6 * (size_t)n;
  w.W = slab + 7 * (size_t)n;
  w.Iw = slab + 8 * (size_t)n;
  w.Last = P;
  const long long room = slab_ints - 8 * (long long)n;
  w.iwlen = room > INT_MAX ? INT_MAX : (int)room;
  int pfree;
  if (symmetric) {
    pfree = build_symmetric(w, Cp, e0, Ci);
  } else {
    if (thread_id() == 0) {
      aat_count_serial(n, Cp, e0, Ci, w.Len, P);
      sh[1] = build_general_serial(w, Cp, e0, Ci, w.Nv, w.W);
    }
    sync_cta();
    pfree = sh[1];
  }
  PB_T0();
  init_state(w, &sh[2]);
  PB_T(6);
  const int nel = sh[2];
  if (thread_id() < WARP) eliminate_warp(w, pfree, nel);
  sync_cta();
  PB_T0R();
  finish_cta(w, sh);
  PB_T(7);
}

}  // namespace amdk
}  // namespace pb
