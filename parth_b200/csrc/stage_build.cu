// parth_b200/csrc/stage_build.cu -- host driver of stage (a): picks the filter path, checks its
// proof flags, falls back to the general kernels (graph_build.cu).
#include <algorithm>
#include <cstring>

#include "build.cuh"
#include "scan.cuh"
#include "stages.cuh"
#include "tree.cuh"

namespace pb {

void read_ints(parth_b200 &h, const int *dsrc, int *dst, size_t n) {
  h.pin.reserve(n);
  PB_CUDA(cudaMemcpyAsync(h.pin.p, dsrc, sizeof(int) * n, cudaMemcpyDeviceToHost, h.stream));
  PB_CUDA(cudaStreamSynchronize(h.stream));
  for (size_t i = 0; i < n; i++) dst[i] = h.pin.p[i];
  h.stats.d2h_bytes += (long long)(sizeof(int) * n);
  collect_timers(h);  // everything recorded so far has completed: free
}

// Enqueue (no read-back) the row diff of the freshly built graph (Mp_own / Mi_own) against the
// snapshot, the ordered list of changed rows and the symmetry check over them.  The results land
// in the mailbox (DetectCounters: changed_rows, dirty_tiles, pad0 = asymmetric) and are cached for
// stage (b).  Rows beyond `cap` are not listed: the caller then falls back to the streaming proof.
static int changed_rows_cap(int M) { return std::max(1024, M / 20); }

static bool spec_ready(const parth_b200 &h, int M) {
  return h.tree_init && h.num_nodes > 0 && h.tree_M == M && h.map_len == 0;
}

// fused_cols > 0: the pipeline build kernel has already written the bitmap, the counters and the
// per-tile counts (h.tile_dirty); three launches finish stages (a) and (b): the row list from the
// tile counts, the symmetry check with the per-row edge counts fused in, and one CTA that scans the
// counts, emits the changed edges in order and marks the dirty nodes.  Returns true when the
// stage (b) result (w5) was enqueued as well.
static bool enqueue_changed_row_proof(parth_b200 &h, int M, int fused_cols, int num_tiles) {
  cudaStream_t s = h.stream;
  const int n_words = (M + 31) / 32;
  const int cap = changed_rows_cap(M);
  h.chg_bits.reserve((size_t)n_words + 1, s);
  h.chg_pre.reserve((size_t)n_words + 2, s);
  h.chg_rows.reserve((size_t)cap + 1, s);
  h.scan_ws.reserve(scan_ws_words((size_t)std::max(n_words, M) + 1), s);
  unsigned *bits = reinterpret_cast<unsigned *>(h.chg_bits.p);
  if (fused_cols > 0) {
    const bool spec = spec_ready(h, M);
    const int nn = h.num_nodes;
    if (spec) {
      h.w3.reserve((size_t)kSpecRows + 2, s);
      if (h.d_changed.cap < 2 * 65536) h.d_changed.reserve(2 * 65536, s);
    }
    int *spec_dev = h.mail.p + 16;  // the result sits behind the mailbox and travels with it in one copy
    h.tile_off.reserve((size_t)num_tiles + 2, s);
    h.stats.kernels_launched += launch_rows_from_tiles(h.tile_dirty.p, h.tile_off.p, h.scan_ws.p, num_tiles, fused_cols, bits,
                                                       M, h.chg_rows.p, cap, s);
    h.stats.kernels_launched += launch_sym_check_rows(h.chg_rows.p, &h.dcnt_p()->changed_rows, cap, h.Mp_own.p,
                                                      h.Mi_own.p, h.Mp_prev.p, h.Mi_prev.p, &h.dcnt_p()->pad0,
                                                      spec ? h.d_dof2node.p : nullptr, spec ? h.w3.p : nullptr,
                                                      kSpecRows, s, spec ? spec_dev : nullptr, spec ? nn * 2 + 2 : 0);
    if (spec)
      h.stats.kernels_launched += launch_emit_mark(h.chg_rows.p, h.dcnt_p(), h.w3.p, h.Mp_own.p, h.Mi_own.p,
                                                   h.d_dof2node.p, h.d_meta.p, h.d_changed.p,
                                                   (int)(h.d_changed.cap / 2), nn, spec_dev, s);
    return spec;
  }
  {
    StageTimer td(h, &h.stats.diff_kernel_ms);
    h.stats.kernels_launched += launch_diff_rows(h.Mp_own.p, h.Mi_own.p, h.Mp_prev.p, h.Mi_prev.p, M, bits, h.dcnt_p(), s);
    td.stop();
  }
  h.stats.kernels_launched += exclusive_scan<1>(h.chg_bits.p, h.chg_pre.p, n_words, h.scan_ws.p, s);
  h.stats.kernels_launched += launch_list_changed_rows(bits, h.chg_pre.p, n_words, M, h.chg_rows.p, cap, s);
  h.stats.kernels_launched += launch_sym_check_rows(h.chg_rows.p, &h.dcnt_p()->changed_rows, cap, h.Mp_own.p, h.Mi_own.p,
                                                    h.Mp_prev.p, h.Mi_prev.p, &h.dcnt_p()->pad0, nullptr, nullptr, 0, s);
  return false;
}

// Stage (b)'s edge kernels (computeTheAddedRemovedEdgesForDOF, findRegionConnections), enqueued
// speculatively behind the changed-row list so that their result (edge count + node flags in w5)
// travels with the build's mailbox: computePermutation then needs no read-back of its own.  The row
// count is still on the device; up to kSpecRows changed rows are covered, more fall back to
// stage_detect's normal path.
static bool enqueue_speculative_detection(parth_b200 &h, int M) {
  if (!spec_ready(h, M)) return false;
  cudaStream_t s = h.stream;
  const int nn = h.num_nodes;
  h.w3.reserve((size_t)kSpecRows + 2, s);
  h.w4.reserve((size_t)kSpecRows + 2, s);
  h.w5.reserve((size_t)nn * 2 + 2, s);
  if (h.d_changed.cap < 2 * 65536) h.d_changed.reserve(2 * 65536, s);
  h.scan_ws.reserve(scan_ws_words((size_t)kSpecRows + 1), s);
  const int *n_rows_dev = &h.dcnt_p()->changed_rows;
  h.stats.kernels_launched += launch_count_row_edges(h.chg_rows.p, kSpecRows, n_rows_dev, h.Mp_own.p, h.Mi_own.p,
                                                     h.d_dof2node.p, h.w3.p, s);
  h.stats.kernels_launched += exclusive_scan<0>(h.w3.p, h.w4.p, kSpecRows, h.scan_ws.p, s);
  PB_CUDA(cudaMemsetAsync(h.w5.p, 0, sizeof(int) * ((size_t)nn * 2 + 2), s));
  const int cap = (int)(h.d_changed.cap / 2);
  h.stats.kernels_launched += launch_emit_row_edges(h.chg_rows.p, kSpecRows, n_rows_dev, h.Mp_own.p, h.Mi_own.p,
                                                    h.d_dof2node.p, h.w4.p, h.d_changed.p, cap, s);
  h.stats.kernels_launched += launch_mark_dirty_nodes_dev(h.d_changed.p, h.w4.p + kSpecRows, cap, h.d_dof2node.p,
                                                          h.d_meta.p, h.w5.p + 2, h.w5.p + 2 + nn, h.w5.p, s);
  return true;
}

void collect_timers(parth_b200 &h) {
  for (auto &t : h.pending_timers) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) {  // a bracket whose shared start event was never recorded
      ms = 0;
      cudaGetLastError();
    }
    if (t.accumulate) *t.dst += ms; else *t.dst = ms;
    if (t.dst2) *t.dst2 = ms;
    h.free_pairs.push_back(t.pair);
  }
  h.pending_timers.clear();
  for (int i = 0; i < 5; i++) h.timers[i] = h.timer_ms[i] * 1e-3;
}

void resolve_timers(parth_b200 &h) {
  PB_CUDA(cudaStreamSynchronize(h.stream));
  collect_timers(h);
}

// what a proven build leaves in the handle (shared by the synchronous path and stage_build_settle)
static void accept_spec(parth_b200 &h, const std::vector<int> &spec_out, int changed_rows) {
  if (h.diff_cached && changed_rows <= kSpecRows && (int)spec_out.size() == h.num_nodes * 2 + 2 && spec_out[0] >= 0 &&
      spec_out[0] <= (int)(h.d_changed.cap / 2)) {
    h.spec_valid = true;
    h.spec_n_changed = spec_out[0];
    h.spec_flags.assign(spec_out.begin() + 2, spec_out.end());
  }
}

static void publish_graph(parth_b200 &h, int M) {
  h.Mp = h.Mp_own.p;
  h.Mi = h.Mi_own.p;
  h.M_n = M;
  h.mesh_owned = true;
  h.map_len = 0;
}

// Completion of a build that stage_build left in flight (may_defer).  Waits for the mailbox copy
// only -- work enqueued behind it (the speculative expansion of computePermutation) keeps running.
// The common outcome (no flag raised, every tile compared in the build kernel, every changed row
// mirrored) is accepted here; anything else re-runs the synchronous state machine on the same
// arrays (they must stay valid until then, see include/parth_b200.h).
void stage_build_settle(parth_b200 &h) {
  if (!h.pend.active) return;
  const parth_b200::PendingBuild p = h.pend;
  h.pend.active = false;
  PB_CUDA(cudaEventSynchronize(h.ev_mail));
  BuildStatus st{};
  DetectCounters dc{};
  std::memcpy(&st, h.pin_mail.p, sizeof(st));
  std::memcpy(&dc, h.pin_mail.p + 8, sizeof(dc));
  h.stats.d2h_bytes += (long long)(sizeof(int) * (12 + p.n_spec));
  {
    // the timers stopped before the mailbox copy have completed with it; later ones stay pending
    std::vector<parth_b200::PendingTimer> later(h.pending_timers.begin() + std::min(p.timers_before, h.pending_timers.size()),
                                                h.pending_timers.end());
    h.pending_timers.resize(std::min(p.timers_before, h.pending_timers.size()));
    collect_timers(h);
    h.pending_timers.swap(later);
  }
  const int M = p.M;
  const bool ok = st.flags == 0 && dc.pad1 == 0 && dc.changed_rows <= changed_rows_cap(M) &&
                  (dc.changed_rows == 0 || dc.pad0 == 0);
  if (!ok) {
    h.stats.build_deferred = 0;
    stage_build(h, p.N, p.dAp, p.dAi, p.nnz, p.dim, false);
    return;
  }
  float t = 0;
  cudaEventElapsedTime(&t, h.ev2, h.ev3);
  h.stats.build_kernel_ms = t;
  h.n_chg_rows = dc.changed_rows;
  h.diff_cached = true;
  std::vector<int> spec_out(h.pin_mail.p + 16, h.pin_mail.p + 16 + p.n_spec);
  accept_spec(h, spec_out, dc.changed_rows);
  h.nnzM = st.nnz_out;
  h.stats.build_path = 1;
  h.stats.build_fused_detect = 1;
  h.stats.build_deferred = 1;
  h.cur_sym_proven = true;
  publish_graph(h, M);
}

void stage_build(parth_b200 &h, int N, const int *dAp, const int *dAi, int nnz, int dim, bool may_defer) {
  cudaStream_t s = h.stream;
  if (h.block_mode)
    throw StatusError(PARTH_B200_ERR_UNSUPPORTED, "the handle holds a row block of a sharded mesh (parth_b200_clear resets it)");
  const int M = N / dim;
  // dim == 1: nothing is enqueued before the filter kernel, so the event recorded right in front of it (ev2) opens
  // the build bracket as well -- one stream operation less while the device is idle between two updates
  StageTimer total(h, &h.stats.build_ms, false, nullptr, dim == 1 ? h.ev2 : nullptr);
  h.stats.proof_kernel_ms = 0;
  h.stats.diff_kernel_ms = 0;

  h.scan_ws.reserve(scan_ws_words((size_t)M + 1), s);
  int Q = nnz;
  if (dim != 1) {
    h.V.reserve((size_t)M + 2, s);
    h.ns_tmp.reserve((size_t)M + 1, s);
    h.stats.kernels_launched += build_sample_prefix_launch(dAp, M, dim, h.ns_tmp.p, h.V.p, h.scan_ws.p, s);
    read_ints(h, h.V.p + M, &Q, 1);
  }

  BuildArgs a{};
  a.Ap = dAp; a.Ai = dAi; a.V = h.V.p; a.dim = dim; a.M = M; a.Q = Q;
  a.pipe_cfg = h.pipe_cfg;
  // Incremental proof: when the snapshot graph is proven symmetric and has the same vertex set,
  // symmetry of the new graph follows from the rows that changed (sym_check_rows_kernel), so the
  // filter kernel skips its mirror probes; the row diff it needs is the one change detection wants
  // anyway and is cached for it.
  h.diff_cached = false;
  h.cur_sym_proven = false;
  h.spec_valid = false;
  const bool incremental = !h.assume_symmetric && h.prev_sym_proven && M == h.M_n_prev &&
                           h.Mp_prev.p && h.Mi_prev.p;
  a.check_mirror = (h.assume_symmetric || incremental) ? 0 : 1;
  const int tiles = build_filter_tiles(M, Q);
  h.tile_c.reserve((size_t)tiles + 2, s);
  h.desc.reserve((size_t)tiles + 1, s);
  // after a snapshot-by-swap the "own" buffers hold stale data and are free to be overwritten
  h.Mp_own.reserve((size_t)M + 1, s);
  h.Mi_own.reserve((size_t)(Q > 0 ? Q : 1), s);
  a.tile_c = h.tile_c.p; a.desc = h.desc.p; a.status = h.bstatus_p();
  a.Mp = h.Mp_own.p; a.Mi = h.Mi_own.p;

  BuildStatus st{};
  DetectCounters dc{};
  a.ev_a = h.ev2; a.ev_b = h.ev3;
  a.status = h.bstatus_p();
  static_assert(sizeof(BuildStatus) == 24 && sizeof(DetectCounters) == 16, "mailbox layout");
  bool spec_enqueued = false;
  std::vector<int> spec_out;
  int fused_cols = 0;  // > 0: the pipeline kernel compares each tile with the snapshot on the fly (fused change detection)
  auto read_mail = [&]() {
    // ONE synchronisation: the mailbox and, when the edge kernels ran speculatively, their result
    int m[12];
    const size_t n_spec = spec_enqueued ? (size_t)h.num_nodes * 2 + 2 : 0;
    h.pin.reserve(16 + n_spec);
    PB_CUDA(cudaMemcpyAsync(h.pin.p, h.mail.p, sizeof(int) * 12, cudaMemcpyDeviceToHost, s));
    if (n_spec && fused_cols > 0)  // the fused tail wrote its result behind the mailbox: same copy, extended
      PB_CUDA(cudaMemcpyAsync(h.pin.p + 12, h.mail.p + 12, sizeof(int) * (4 + n_spec), cudaMemcpyDeviceToHost, s));
    else if (n_spec)
      PB_CUDA(cudaMemcpyAsync(h.pin.p + 16, h.w5.p, sizeof(int) * n_spec, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 12; i++) m[i] = h.pin.p[i];
    spec_out.assign(h.pin.p + 16, h.pin.p + 16 + n_spec);
    h.stats.d2h_bytes += (long long)(sizeof(int) * (12 + n_spec));
    collect_timers(h);
    std::memcpy(&st, m, sizeof(st));
    std::memcpy(&dc, m + 8, sizeof(dc));
  };
  auto launch_filter = [&]() {
    fused_cols = 0;
    const bool was_clean = h.mail_clean;
    h.mail_clean = false;
    a.Mp_prev = a.Mi_prev = nullptr;
    a.tile_dirty = nullptr;
    a.row_bits = nullptr;
    a.dcnt = nullptr;
    if (dim == 1 && h.build_mode <= 1) {
      if (h.build_mode == 0 && build_pipe_supported(a)) {
        if (incremental && !a.check_mirror) {
          h.tile_dirty.reserve((size_t)build_pipe_num_tiles(a) + 1, s);
          a.Mp_prev = h.Mp_prev.p; a.Mi_prev = h.Mi_prev.p; a.nnz_prev = h.nnz_prev; a.tile_dirty = h.tile_dirty.p;
          h.chg_bits.reserve((size_t)(M + 31) / 32 + 1, s);
          a.row_bits = reinterpret_cast<unsigned *>(h.chg_bits.p);
          a.dcnt = h.dcnt_p();
          fused_cols = build_pipe_tile_cols(a);
        }
        a.mail_prezeroed = (fused_cols > 0 && was_clean) ? 1 : 0;
        h.stats.kernels_launched += build_pipe_launch(a, h.num_sms, s);
      }
      else h.stats.kernels_launched += build_cols_launch(a, h.build_mode == 0, s);
    } else if (dim > 1 && h.build_mode_n == 0) {
      // closed-form offsets, strided samples; the full symmetry proof runs on the compact output
      const int want_proof = a.check_mirror;
      a.check_mirror = 0;
      h.stats.kernels_launched += build_cols_launch(a, true, s);
      a.check_mirror = want_proof;
      if (want_proof) {
        StageTimer tp(h, &h.stats.proof_kernel_ms);
        h.stats.kernels_launched += sym_proof_csr_launch(a.Mp, a.Mi, M, Q, a.status, s);
        tp.stop();
      }
    } else {
      h.stats.kernels_launched += build_filter_launch(a, s);
    }
  };
  auto cols_family = [&]() { return (dim == 1 && h.build_mode <= 1) || (dim > 1 && h.build_mode_n == 0); };
  auto diff_ok = [&]() { return cols_family() ? st.diff_sum == 0 : st.diff_mod == 0; };
  {
    const int tiles7 = build_cols_tiles(Q);
    h.tile_c.reserve((size_t)std::max(tiles, tiles7) + 2, s);
    h.desc.reserve((size_t)std::max(tiles, tiles7) + 1, s);
    a.tile_c = h.tile_c.p; a.desc = h.desc.p;
  }
  bool proven = false;
  for (;;) {
    // the filter kernel, then (incremental mode) the changed-row proof, then ONE read-back
    launch_filter();
    spec_enqueued = false;
    if (incremental) {
      spec_enqueued = enqueue_changed_row_proof(h, M, fused_cols, fused_cols ? build_pipe_num_tiles(a) : 0);
      if (!fused_cols) spec_enqueued = enqueue_speculative_detection(h, M);
    }
    if (may_defer && incremental && fused_cols > 0 && spec_enqueued && h.build_mode == 0) {
      // steady state of the device-pointer API: leave everything in flight (stage_build_settle)
      const size_t n_spec = (size_t)h.num_nodes * 2 + 2;
      h.pin_mail.reserve(16 + n_spec);
      PB_CUDA(cudaMemcpyAsync(h.pin_mail.p, h.mail.p, sizeof(int) * (16 + n_spec), cudaMemcpyDeviceToHost, s));
      total.stop(h.ev_mail);  // the mailbox event doubles as the end of the build bracket
      PB_CUDA(cudaEventRecord(h.ev_mail, s));
      // status + counters for the NEXT build, zeroed here where the stream is busy anyway
      PB_CUDA(cudaMemsetAsync(h.mail.p, 0, sizeof(int) * 12, s));
      h.mail_clean = true;
      h.pend.active = true;
      h.pend.N = N; h.pend.nnz = nnz; h.pend.dim = dim; h.pend.M = M; h.pend.fused_cols = fused_cols;
      h.pend.dAp = dAp; h.pend.dAi = dAi;
      h.pend.n_spec = n_spec;
      h.pend.timers_before = h.pending_timers.size();
      return;
    }
    may_defer = false;
    read_mail();
    if (incremental && fused_cols && st.flags == 0 && dc.pad1 > 0) {
      // some tiles could not be compared inside the build kernel (over-long columns, snapshot tile
      // larger than the staging buffer): the streaming row diff takes over for this update
      fused_cols = 0;
      enqueue_changed_row_proof(h, M, 0, 0);
      spec_enqueued = enqueue_speculative_detection(h, M);
      read_mail();
    }
    if (dim == 1 && h.build_mode <= 1) {
      // remember the variant that works so that steady-state updates launch it directly
      const unsigned only = st.flags & ~(BUILD_NODIAG | BUILD_WIDE);
      if (h.build_mode == 0 && (st.flags & BUILD_NODIAG) && !only) { h.build_mode = 1; continue; }
      if ((st.flags & BUILD_WIDE) && !only && h.build_mode == 0 && h.pipe_cfg == 0 && build_pipe_supported(a)) {
        h.pipe_cfg = a.pipe_cfg = 1;  // a tile overflowed the compact shape: the roomier one takes over for this handle
        continue;
      }
      if ((st.flags & BUILD_WIDE) && !only) { h.build_mode = 2; continue; }
    } else if (dim > 1 && h.build_mode_n == 0) {
      const unsigned only = st.flags & ~(BUILD_NODIAG | BUILD_WIDE);
      if ((st.flags & (BUILD_NODIAG | BUILD_WIDE)) && !only) { h.build_mode_n = 1; continue; }  // sample-tiled, chained scan
    }
    break;
  }
  h.stats.build_kernel_ms = 0;
  if (Q > 0) {  // the filter kernel of the last launch alone (events recorded around it)
    float t = 0;
    cudaEventElapsedTime(&t, h.ev2, h.ev3);
    h.stats.build_kernel_ms = t;
  }
  if (st.flags & BUILD_RANGE)
    throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: row index out of range");
  if (st.flags == 0) {
    if (h.assume_symmetric) proven = true;
    else if (!incremental) proven = diff_ok();
    else {
      h.n_chg_rows = dc.changed_rows;
      h.diff_cached = dc.changed_rows <= changed_rows_cap(M);
      if (dc.changed_rows == 0) proven = true;                       // identical to a proven graph
      else if (h.diff_cached && dc.pad0 == 0) proven = true;         // every changed row is mirrored
      else if (h.diff_cached && dc.pad0 != 0) proven = false;        // a mirror is missing: symmetrise below
      else {
        // more than 5 % of the rows changed: the streaming proof is cheaper than probing them
        a.check_mirror = 1;
        launch_filter();
        read_mail();
        proven = st.flags == 0 && diff_ok();
        h.diff_cached = false;
      }
    }
  }
  if (proven && incremental && spec_enqueued) accept_spec(h, spec_out, dc.changed_rows);
  h.stats.build_deferred = 0;
  if (proven) {
    h.nnzM = st.nnz_out;
    h.stats.build_path = 1;
    h.stats.build_fused_detect = fused_cols > 0 ? 1 : 0;
    h.cur_sym_proven = !h.assume_symmetric;
  } else {
    // general path: any triangle, unsorted columns, duplicates
    h.cnt.reserve((size_t)M + 1, s);
    h.ptr.reserve((size_t)M + 2, s);
    h.uniq.reserve((size_t)M + 1, s);
    h.stats.kernels_launched += build_general_count_launch(a, h.cnt.p, h.ptr.p, h.scan_ws.p, s);
    int total_pairs = 0;
    read_ints(h, h.ptr.p + M, &total_pairs, 1);
    read_mail();
    if (st.flags & BUILD_RANGE)
      throw StatusError(PARTH_B200_ERR_INPUT, "set_matrix: row index out of range");
    h.tmp.reserve((size_t)(total_pairs > 0 ? total_pairs : 1), s);
    h.Mi_own.reserve((size_t)(total_pairs > 0 ? total_pairs : 1), s);
    a.Mi = h.Mi_own.p;
    h.stats.kernels_launched +=
        build_general_finish_launch(a, h.cnt.p, h.ptr.p, h.tmp.p, h.uniq.p, h.scan_ws.p, s);
    int nnzM = 0;
    read_ints(h, h.Mp_own.p + M, &nnzM, 1);
    h.stats.kernels_launched += build_general_gather_launch(a, h.ptr.p, h.tmp.p, s);
    h.nnzM = nnzM;
    h.stats.build_path = 2;
    h.stats.build_fused_detect = 0;
    h.cur_sym_proven = true;  // the general path symmetrises by construction
    h.diff_cached = false;
  }
  total.stop();
  publish_graph(h, M);
}

}  // namespace pb
