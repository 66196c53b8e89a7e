// parth_b200/csrc/handle.cuh -- the opaque handle behind include/parth_b200.h.
//
// Device state is flat SoA int32 (no per-node std::vectors as in reference HMD.h:18-104):
//   graph     Mp/Mi (current), Mp_prev/Mi_prev (snapshot, swapped not copied when owned)
//   tree      meta on host AND device (<= 8191 nodes), node_ptr, dofs, labels, dof2node, mesh_perm
//   scratch   grow-only buffers, so a steady-state update performs no allocation
#pragma once
#include <string>
#include <vector>

#include "../../include/parth_b200.h"
#include "build.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "hostcopy.cuh"
#include "hosttrace.cuh"
#include "tree.cuh"

struct parth_b200 {
  int device = 0, num_sms = pb::kNumSMsB200;
  cudaStream_t stream = nullptr;
  // side stream of the sparse relocation: the streaming copy of the untouched slices runs beside the joiner ranking and
  // the regroup kernel of the changed nodes (forked from and joined back into `stream` inside the stage)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // deferred stage timers: events are recorded while the work is enqueued and read once, after the
  // final synchronisation of the API call (no intermediate cudaEventSynchronize on the hot path)
  static constexpr int kTimerPairs = 48;
  cudaEvent_t tev[2 * kTimerPairs] = {};
  struct PendingTimer { double *dst; double *dst2; int pair; bool accumulate; cudaEvent_t a, b; };
  std::vector<PendingTimer> pending_timers;  // stopped, not yet read
  std::vector<int> free_pairs;               // event pairs nobody holds (filled at creation)
  std::string err;

  // ---- options (reference ParthAPI.h:19-25, Integrator.h:24-28)
  int reorder_type = PARTH_B200_METIS, nd_levels = 8, lagging = 1, aggressive = 0;
  int relocate_lvl = 2, lag_lvl = 5, verbose = 1, sim_dim = 1;
  int coarse_policy = PARTH_B200_COARSE_REPAIR, assume_symmetric = 0;
  int build_mode_n = 0;  // dim > 1: 0 closed-form offsets (column-aligned), 1 sample-tiled chained scan
  int pipe_cfg = pb::build_pipe_default_cfg();  // pipeline kernel tile shape: 0 compact (3 CTAs / SM), 1 after a tile overflowed it
  int build_mode = 0;  // dim == 1 filter kernel that proved itself last time: 0 closed-form offsets, 1 chained scan, 2 sample-tiled
  int rank = 0, world = 1;
  std::vector<int> xchg_nodes, xchg_owner;  // sharded mode: plan of the last update
  bool xchg_pending = false;
  bool builtin_decomposer = false, force_amd = false;
  parth_b200_decomposer decomposer = nullptr;
  void *decomposer_user = nullptr;

  // ---- matrix staging for the host-pointer API
  pb::DevBuf<int> dAp, dAi;

  // ---- mesh graph
  int M_n = 0, nnzM = 0;
  const int *Mp = nullptr, *Mi = nullptr;  // device pointers of the current graph
  bool mesh_owned = false;                 // Mp/Mi live in Mp_own/Mi_own (or were swapped into *_prev)
  pb::DevBuf<int> Mp_own, Mi_own;
  int M_n_prev = 0, nnz_prev = 0;
  pb::DevBuf<int> Mp_prev, Mi_prev;

  // ---- row-block mode: ONE mesh over the ranks of a communicator (stage_sharded.cu).  Mp_own / Mi_own / Mp_prev /
  // Mi_prev then hold only the rows [blk_c0, blk_c1) of the two graphs; h.Mp / h.Mi and the pointers handed to the
  // kernels are BIASED (Mp_own.p - blk_c0, Mi_own.p + own_off - own_q0) so that global row ids and global entry
  // offsets index them; own_q0 / own_q1 = global Mp[blk_c0] / Mp[blk_c1], own_off = own_q0 & 3 keeps the biased
  // Mi pointer 16-byte aligned for the bulk copies.  The separator tree and every per-DOF array stay replicated.
  pb::Comm *comm = nullptr;
  bool block_mode = false;
  int blk_TC = 0, blk_t0 = 0, blk_t1 = 0, blk_tiles = 0, blk_c0 = 0, blk_c1 = 0;
  int own_off = 0, prev_off = 0, own_q0 = 0, own_q1 = 0, prev_q0 = 0, prev_q1 = 0;
  const int *Mp_own_b() const { return Mp_own.p - blk_c0; }
  const int *Mi_own_b() const { return Mi_own.p + own_off - own_q0; }
  const int *Mp_prev_b() const { return Mp_prev.p - blk_c0; }
  const int *Mi_prev_b() const { return Mi_prev.p + prev_off - prev_q0; }
  pb::DevBuf<int> x_hdr, x_buf, x_off, x_glist;  // exchange: own slot, gathered slots, reduced mailbox, graph share
  int x_cap = 4096;            // edge / diff pairs a slot holds (grown when a rank overflows it)
  bool blk_tiles_zeroed = false;
  int n_chg_rows_total = 0;

  // ---- symmetry bookkeeping (stage a): a proven-symmetric snapshot lets the next build prove its
  // graph by checking only the rows that changed; the row diff is then reused by change detection
  bool cur_sym_proven = false, prev_sym_proven = false;
  bool diff_cached = false;   // chg_bits / chg_rows describe (current graph vs snapshot)
  int n_chg_rows = 0;
  pb::DevBuf<int> chg_bits, chg_pre, chg_rows;

  // ---- DOF maps (reference ParthAPI.h:79-84)
  int map_len = 0;  // 0: no map
  pb::DevBuf<int> new_to_old, old_to_new;
  int n_added = 0, n_removed = 0;
  pb::DevBuf<int> added, removed;

  // ---- separator tree
  bool tree_init = false;
  bool meta_heap = false;  // stored parent / level fields agree with heap arithmetic (checked at import)
  int num_levels = 0, num_nodes = 0, tree_M = 0;
  std::vector<int> meta;      // host mirror, nodes*7
  std::vector<int> node_ptr;  // host mirror, nodes+1
  pb::DevBuf<int> d_meta, d_node_ptr, d_visit, d_dofs, d_labels, d_dof2node, d_mesh_perm;
  pb::DevBuf<int> d_dofs2, d_labels2;  // double buffers for integration / relocation
  // sparse relocation (stage_maps.cu): per-DOF occurrence counts / relocation targets, clean between calls
  pb::DevBuf<int> rel_occ, rel_target, rank_seg;
  // sticky error flags of the relocation's regroup kernels ([0] unsorted survivors, [1] size mismatch).  The
  // host-pointer entry points read them before they return; the stream-ordered *_dev entry points leave them to the
  // next call that synchronises anyway (relocation read-back, synchronize, get_stats, any host-pointer call)
  pb::DevBuf<int> reloc_sticky;
  bool defer_reloc_check = false, reloc_check_pending = false;
  int rel_M = 0;
  pb::PinBuf<int> pin_rel;

  // ---- results of the last computePermutation
  double reuse = 0.0;
  int n_changed = 0;
  pb::DevBuf<int> d_changed;  // pairs
  std::vector<int> dirty_fine, dirty_coarse, dirty_fine_saved, dirty_coarse_saved;
  double timers[5] = {0, 0, 0, 0, 0};     // seconds, reference order (ParthAPI.h:87-91)
  double timer_ms[5] = {0, 0, 0, 0, 0};   // the same spans in device milliseconds

  // ---- scratch
  pb::DevBuf<int> V, ns_tmp, tile_c, cnt, ptr, tmp, uniq, tile_dirty, tile_off;
  pb::DevBuf<unsigned long long> desc, scan_ws;
  // device mailbox, read back with ONE copy: [0..5] BuildStatus, [8..11] DetectCounters
  pb::DevBuf<int> mail;
  pb::BuildStatus *bstatus_p() { return reinterpret_cast<pb::BuildStatus *>(mail.p); }
  pb::DetectCounters *dcnt_p() { return reinterpret_cast<pb::DetectCounters *>(mail.p + 8); }
  // the first 12 ints of the mailbox were zeroed on the stream right behind the last deferred build's mailbox copy and
  // nothing has been enqueued since that writes them: the next fused build skips its memset (one stream operation
  // less between two updates, where the device is idle)
  bool mail_clean = false;
  // node flags fetched together with the changed-edge count (stage b)
  bool node_flags_valid = false;
  std::vector<int> node_flags;
  // stage (b) results computed speculatively behind the build (steady state): edge count + node flags
  bool spec_valid = false;
  int spec_n_changed = 0;
  std::vector<int> spec_flags;
  pb::DevBuf<int> w0, w1, w2, w3, w4, w5;  // general-purpose int scratch
  // stage (c): node list, vertex offsets, per-vertex counts, sub-mesh CSR, pivots, AMD slabs
  pb::DevBuf<int> r_list, r_vtx, r_cnt, r_G, r_sub, r_perm, r_ws, r_chosen, r_g2l, r_l2g;
  pb::DevBuf<long long> r_wsptr;
  // ReorderingType MORTON_CODE (morton.cu): key / point-id pairs (double buffered), global rank, the same for a node batch
  int morton_n = 0, morton_bits = 21, morton_cur = 0, morton_passes = 0;
  bool morton_rank_valid = false;
  pb::DevBuf<unsigned long long> m_k0, m_k1, m_nk0, m_nk1, m_box, m_desc;
  pb::DevBuf<int> m_v0, m_v1, m_nv0, m_nv1, m_rank, m_hist;
  pb::DevBuf<double> m_xyz;                // staging of host coordinates
  pb::PinBuf<unsigned long long> pin_u64;
  pb::DevBuf<int> d_out;                   // staging for host-pointer outputs
  pb::PinBuf<int> pin;                     // small pinned read-backs
  pb::PinBuf<int> pin_mail;                // mailbox read-back of a deferred build (own buffer: nothing may move it)

  // ---- deferred completion of stage (a): the stream-ordered device-pointer set_matrix enqueues the
  // build, the changed-row proof, stage (b)'s edge kernels and the mailbox copy, records ev_mail and
  // returns; the next call that needs the result waits for ev_mail only (stage_build_settle), so the
  // device never idles while the host reads flags in the steady state
  struct PendingBuild {
    bool active = false;
    int N = 0, nnz = 0, dim = 1, M = 0, fused_cols = 0;
    const int *dAp = nullptr, *dAi = nullptr;
    size_t n_spec = 0, timers_before = 0;
  } pend;
  cudaEvent_t ev_mail = nullptr;
  pb::PinBuf<int> pin_big;                 // pinned staging for big host transfers
  pb::HostCopier hostcopy;                 // caller arrays <-> device (staged through page-locked chunks when unregistered)

  parth_b200_stats stats = {};
};
