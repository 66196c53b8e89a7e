/*
 * include/parth_b200.h -- C ABI of the B200-native dynamic-reuse ordering path.
 *
 * The reference has no C ABI: its boundary is the C++ class PARTH::ParthAPI
 * (reference include/parth/parth.h:25 -> src/Parth/include/ParthAPI.h:16-145).  The drop-in class
 * in include/parth/ParthAPI.h keeps that surface and forwards every call to the entry points
 * below; each entry point cites the reference member it replaces.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns an int status: 0 = PARTH_B200_OK, negative = PARTH_B200_ERR_*,
 *     positive = a cudaError_t passed through.  parth_b200_last_error() gives the text.
 *   - pointers are HOST pointers unless the function name ends in _dev.
 *   - one handle owns one CUDA stream and all device state; like the reference class it is not
 *     thread-safe and not re-entrant.
 *   - there is NO CPU fallback: if no CUDA device is usable, parth_b200_create fails.
 *
 * Separator-tree fixture layout (flat int32; the reference keeps per-node std::vectors,
 * HMD.h:18-104):
 *   meta[node*7 + k], k = 0..6 : left, right, node_id, parent, offset, level, org_size
 *                                heap order, children of i are 2i+1 / 2i+2, empty node <=> node_id -1
 *   node_ptr[nodes+1]          : start of each node's slice in dofs[] / labels[]
 *   dofs[M_n]                  : HMDNode::DOFs concatenated in heap order
 *   labels[M_n]                : HMDNode::permuted_new_label, same indexing
 */
#ifndef PARTH_B200_H
#define PARTH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct parth_b200 parth_b200;

enum {
  PARTH_B200_OK = 0,
  PARTH_B200_ERR_ARG = -1,        /* null pointer / inconsistent sizes */
  PARTH_B200_ERR_NO_DEVICE = -2,  /* no usable CUDA device: the library has no CPU path */
  PARTH_B200_ERR_NO_TREE = -3,    /* cold start requested but no tree imported and no decomposer set */
  PARTH_B200_ERR_INPUT = -4,      /* matrix indices out of range */
  PARTH_B200_ERR_INTERNAL = -5,
  PARTH_B200_ERR_UNSUPPORTED = -6 /* e.g. METIS leaf ordering requested on the device */
};

/* ReorderingType, reference ParthTypes.h:11 */
enum { PARTH_B200_METIS = 0, PARTH_B200_AMD = 1, PARTH_B200_MORTON_CODE = 2 };

/* what to do with dirty COARSE nodes (the reference re-dissects them with METIS,
 * Integrator.cpp:812-849, which is outside the device path -- see DESIGN.md) */
enum {
  PARTH_B200_COARSE_REPAIR = 0,   /* device: relocate one endpoint of every broken edge into the
                                     common separator, re-order the touched nodes (default) */
  PARTH_B200_COARSE_CALLBACK = 1, /* host callback re-dissects the extracted sub-mesh */
  PARTH_B200_COARSE_REPORT = 2,   /* leave the permutation, report PARTH_B200_NEEDS_REDISSECT */
  PARTH_B200_COARSE_RELOCATE = 3  /* device: the relocation of REPAIR without its re-ordering -- a separator that
                                     receives DOFs keeps the order of the DOFs it had and takes the new ones at its
                                     end (ascending id), a node that loses DOFs keeps the order of the rest.  The
                                     columns of a separator form a nearly dense block of L whatever their order, so
                                     fill stays within a fraction of a percent of REPAIR (profiles/) at none of the
                                     ordering cost */
};
#define PARTH_B200_NEEDS_REDISSECT 100 /* positive, non-error status of compute_permutation */
#define PARTH_B200_NEEDS_EXCHANGE 101  /* sharded mode: gather the re-ordered nodes, then call finish */

const char *parth_b200_version(void);
const char *parth_b200_last_error(const parth_b200 *h);

/* ParthAPI::ParthAPI / ~ParthAPI, ParthAPI.cpp:47-60.  device < 0: current device. */
int parth_b200_create(parth_b200 **out, int device);
int parth_b200_destroy(parth_b200 *h);
/* ParthAPI::clearParth, ParthAPI.cpp:62-74 */
int parth_b200_clear(parth_b200 *h);

/* setReorderingType/NDLevels/Verbose + public fields lagging, activate_aggressive_reuse,
 * integrator.relocate_lvl / lag_lvl; ParthAPI.cpp:13-45, ParthAPI.h:19-25, Integrator.h:24-28 */
int parth_b200_set_options(parth_b200 *h, int reorder_type, int nd_levels, int lagging,
                           int aggressive_reuse, int relocate_lvl, int lag_lvl, int verbose);
int parth_b200_set_coarse_policy(parth_b200 *h, int policy);
/* optional hint: the caller guarantees a structurally symmetric matrix with sorted columns, so
 * the build skips its symmetry proof (default 0 = prove it on the device, fall back to the
 * general sort/dedupe kernels when the proof fails) */
int parth_b200_set_assume_symmetric(parth_b200 *h, int on);

/* ParthAPI::setMatrix -> computeMeshFromMatrix, ParthAPI.cpp:100-124,366-418 */
int parth_b200_set_matrix(parth_b200 *h, int N, const int *Ap, const int *Ai, int dim);
/* device-pointer variant: STREAM-ORDERED on the handle's stream.  In the steady state of the reuse
 * path (same vertex set as a proven-symmetric snapshot, tree present, no DOF map) it returns with the
 * build, the changed-row proof and the edge kernels of change detection in flight; the next call on
 * the handle that needs their outcome waits for the flags alone.  Consequences for the caller:
 * dAp / dAi must stay valid and unchanged until that next call has returned (normally
 * parth_b200_compute_permutation_dev, or parth_b200_synchronize), and an input error found on the
 * device (PARTH_B200_ERR_INPUT) is reported by that call instead of this one. */
int parth_b200_set_matrix_dev(parth_b200 *h, int N, const int *dAp, const int *dAi, int nnz, int dim);
/* ParthAPI::setMesh, ParthAPI.cpp:76-98 (host arrays are copied to the device) */
int parth_b200_set_mesh(parth_b200 *h, int n, const int *Mp, const int *Mi);
int parth_b200_set_mesh_dev(parth_b200 *h, int n, const int *dMp, const int *dMi, int nnz);
/* ParthAPI::setNewToOldDOFMap -> computeOldToNewDOFMap, ParthAPI.cpp:126-189; len 0 clears */
int parth_b200_set_new_to_old_map(parth_b200 *h, const int *map, int len);
/* same with the map in device memory (a remesher that runs on the GPU) */
int parth_b200_set_new_to_old_map_dev(parth_b200 *h, const int *dmap, int len);

/* ParthAPI::old_to_new_map / added_dofs / removed_dofs (ParthAPI.h:79-84) as computeOldToNewDOFMap
 * (ParthAPI.cpp:131-189) left them; two-call protocol: array pointers may be NULL to get the sizes.  All sizes are 0
 * when no map is active (none set, dropped for lack of a previous mesh, or cleared by the 0.4*M_n cold-start rule). */
int parth_b200_get_dof_maps(parth_b200 *h, int *len_old, int *old_to_new, int *n_added, int *added, int *n_removed,
                            int *removed);
/* ParthAPI::Mp_prev / Mi_prev / M_n_prev (ParthAPI.h:71-73): the snapshot graph; array pointers may be NULL */
int parth_b200_get_snapshot(parth_b200 *h, int *M_n_prev, int *nnz_prev, int *Mp_prev, int *Mi_prev);

/* current mesh graph (ParthAPI::Mp / Mi / M_n, ParthAPI.h:66-71) */
int parth_b200_mesh_dims(parth_b200 *h, int *M_n, int *nnz);
int parth_b200_get_mesh(parth_b200 *h, int *Mp, int *Mi);
int parth_b200_mesh_dev(parth_b200 *h, const int **dMp, const int **dMi);

/* tree fixture in / out (ParthAPI::hmd public members, HMD.h:94-104) */
int parth_b200_import_tree(parth_b200 *h, int M_n, int levels, int nodes, const int *meta,
                           const int *node_ptr, const int *dofs, const int *labels,
                           const int *Mp_prev, const int *Mi_prev);
int parth_b200_tree_dims(parth_b200 *h, int *M_n, int *levels, int *nodes);
int parth_b200_export_tree(parth_b200 *h, int *meta, int *node_ptr, int *dofs, int *labels,
                           int *dof2node, int *mesh_perm);

/* cold decomposition hook (HMD::initHMD, HMD.cpp:89-120): called with the host mesh when
 * computePermutation needs a tree and none exists; must fill the fixture arrays. */
typedef int (*parth_b200_decomposer)(void *user, int M_n, const int *Mp, const int *Mi, int levels,
                                     int nodes, int *meta, int *node_ptr, int *dofs, int *labels);
int parth_b200_set_decomposer(parth_b200 *h, parth_b200_decomposer fn, void *user);
/* self-contained default for the cold start when no callback is set: host-side nested dissection
 * by BFS level structures, every node then ordered by the device AMD kernel.  Off by default on
 * the raw C ABI (a missing tree is an error); the C++ drop-in class switches it on. */
int parth_b200_use_builtin_decomposer(parth_b200 *h, int on);

/* ParthAPI::computePermutation, ParthAPI.cpp:191-279.  perm must hold M_n*dim ints. */
int parth_b200_compute_permutation(parth_b200 *h, int *perm, int dim);
/* device-pointer variant: STREAM-ORDERED.  dperm is written by work enqueued on the handle's stream
 * (parth_b200_stream); it is valid for later work on that stream, or on the host after
 * parth_b200_synchronize.  No synchronisation is done just to return. */
int parth_b200_compute_permutation_dev(parth_b200 *h, int *dperm, int dim);
/* After PARTH_B200_NEEDS_REDISSECT (COARSE_REPORT policy) perm holds the tree's CURRENT permutation (valid, but
 * nothing was re-ordered) and the snapshot has NOT advanced: the next update reports the unresolved edges again.
 * A caller that has dealt with the report -- it re-dissected and imported a new tree, or it accepts the tree as it
 * is -- advances the snapshot to the current graph with this call (the copy at the end of ParthAPI::computePermutation,
 * ParthAPI.cpp:270-278). */
int parth_b200_accept_snapshot(parth_b200 *h);
/* ParthAPI::mapMeshPermToMatrixPerm, ParthAPI.cpp:285-337 (dim == -1: sim_dim) */
int parth_b200_map_mesh_perm_to_matrix_perm(parth_b200 *h, const int *mesh_perm, int n, int *out,
                                            int dim);
int parth_b200_map_mesh_perm_to_matrix_perm_dev(parth_b200 *h, const int *dmesh_perm, int n,
                                                int *dout, int dim);

/* ParthAPI::getReuse / getNumChanges, ParthAPI.cpp:339-341; changed_dof_edges, ParthAPI.h:77;
 * integrator.dirty_*_HMD_nodes_saved, Integrator.h:20-21; timers ParthAPI.h:87-91 in the order
 * init, dof_change_integrator, change_computation, compute_permutation, map_mesh_to_matrix */
double parth_b200_get_reuse(parth_b200 *h);
int parth_b200_get_num_changes(parth_b200 *h);
int parth_b200_get_changed_edges(parth_b200 *h, int *pairs, int cap_pairs);
int parth_b200_get_dirty_nodes(parth_b200 *h, int *fine, int *nfine, int *coarse, int *ncoarse);
int parth_b200_get_timers(parth_b200 *h, double t[5]);
int parth_b200_get_mesh_perm(parth_b200 *h, int *mesh_perm, int cap);
int parth_b200_mesh_perm_dev(parth_b200 *h, const int **dmesh_perm);

/* stage-wise entry points (parity tests drive the same public Integrator/HMD members in the
 * reference: Integrator.h:40-144, HMD.h:112-252) */
int parth_b200_stage_integrate(parth_b200 *h);             /* Integrator::integrateDOFChanges */
int parth_b200_stage_detect(parth_b200 *h, int *nchanged); /* Integrator::computeChangedEdges */
int parth_b200_stage_permute_fine(parth_b200 *h, const int *nodes, int n); /* permuteFineHMDNodes */
int parth_b200_stage_submesh(parth_b200 *h, int node, int coarse, int *n, int *nnz, int *Mp,
                             int *Mi, int *local_to_global); /* createMultiSubMesh / getCoarseMesh */
/* approximate-minimum-degree ordering of one graph on the device (HMD::AMDNodeNDPermutation,
 * HMD.cpp:58-75); batch of `count` graphs concatenated CSR-of-CSR style */
int parth_b200_amd_batch(parth_b200 *h, int count, const int *graph_ptr /*count+1, in vertices*/,
                         const int *Mp /*per graph n+1, concatenated*/, const int *Mi, int *perm);

/* ReorderingType MORTON_CODE.  The reference declares the type (ParthTypes.h:11) but HMD::Permute only prints
 * "Morton code is not implemented yet" (HMD.cpp:82-83) and ParthAPI has no coordinate input, so these entry
 * points are an extension: set_coordinates takes one 3-D point per current DOF (point i, axis a at
 * xyz[i * point_stride + a * comp_stride]: 3, 1 for interleaved points; 1, n for an Eigen column-major V),
 * quantises each axis of the bounding box to `bits` (1..21) bits, interleaves (x most significant) and sorts
 * ascending by (key, point id).  With reorder_type MORTON_CODE the fine nodes of computePermutation /
 * stage_permute_fine are then ordered by that global rank; without coordinates the request fails with
 * PARTH_B200_ERR_ARG.  morton_order returns the global order (perm[new] = old), morton_keys the sorted keys. */
int parth_b200_set_coordinates(parth_b200 *h, const double *xyz, int n, long long point_stride,
                               long long comp_stride, int bits);
int parth_b200_set_coordinates_dev(parth_b200 *h, const double *dxyz, int n, long long point_stride,
                                   long long comp_stride, int bits);
int parth_b200_morton_order(parth_b200 *h, int *perm, int cap);
int parth_b200_morton_order_dev(parth_b200 *h, const int **dperm, int *n);
int parth_b200_morton_keys(parth_b200 *h, unsigned long long *keys_sorted, int cap);

/* stage (d): elimination tree, column counts and nnz(L) of P*A*P' for the current mesh graph and
 * permutation perm[new]=old (NULL: current mesh_perm).  Replaces cholmod_analyze_p as used in
 * reference src/utils/ParthTestUtils.cpp:15-59; etree as src/utils/etree.cpp:39-115. */
int parth_b200_symbolic(parth_b200 *h, const int *perm, int *parent, int *colcount, int64_t *lnz);
/* same with device pointers: the etree and the column counts stay on the device for the consumer of the
 * ordering (the symbolic phase of a device sparse Cholesky: reference src/LinSysSolver/LinSysSolver.hpp:92-113,
 * CHOLMODSolver.cpp:108-151 recompute them on the host).  dperm NULL: current mesh_perm. */
int parth_b200_symbolic_dev(parth_b200 *h, const int *dperm, int *dparent, int *dcolcount, int64_t *lnz);
/* the step after stage (d): fundamental supernodes of L from the elimination tree and the column
 * counts, on the device -- what the supernodal analysis behind reference src/LinSysSolver/CHOLMODSolver.cpp:108-151
 * (cholmod_analyze_p, cm.supernodal = CHOLMOD_SUPERNODAL) derives first, after recomputing both arrays on the host.
 * Column j > 0 continues the supernode of j-1 iff parent[j-1] == j, colcount[j-1] == colcount[j] + 1 and j has
 * exactly one child.  Outputs (each may be NULL): super[0..nsuper] first column of every supernode (super[nsuper] = n;
 * capacity n + 1), supermap[n] supernode of every column, sparent[nsuper] supernodal elimination tree (-1 = root;
 * capacity n); *nsuper, *ssize = sum of colcount[first column] (row indices of the supernodal structure),
 * *xsize = sum of ncols * colcount[first column] (numerical entries).  Relaxed amalgamation is not applied. */
int parth_b200_supernodes(parth_b200 *h, int n, const int *parent, const int *colcount, int *super, int *supermap,
                          int *sparent, int *nsuper, int64_t *ssize, int64_t *xsize);
/* same with device pointers for the five arrays (stream-ordered behind parth_b200_symbolic_dev) */
int parth_b200_supernodes_dev(parth_b200 *h, int n, const int *dparent, const int *dcolcount, int *dsuper,
                              int *dsupermap, int *dsparent, int *nsuper, int64_t *ssize, int64_t *xsize);

/* instrumentation: device time of the last call's stages (CUDA events on the handle's stream)
 * and the number of kernels launched by the library since creation */
typedef struct {
  double build_ms, detect_ms, integrate_ms, dirty_ms, reorder_ms, assemble_ms, expand_ms, symbolic_ms;
  double build_kernel_ms;       /* dominant build kernel only */
  long long kernels_launched;
  int build_path;               /* 1 = proven-symmetric filter path, 2 = general sort/dedupe path */
  long long h2d_bytes, d2h_bytes; /* bytes moved by the last set_* / compute_* call */
  int integrate_passes, integrate_sweeps; /* added-DOF placement: FIFO passes / wavefront sweeps */
  double symbolic_etree_ms;               /* etree part of symbolic_ms */
  double diff_kernel_ms;                  /* diff_rows_kernel alone (last build or detection that ran it) */
  int symbolic_path;                      /* 2 = per-node shared-memory kernel, 1 = wave-scheduled walk, 0 = one range */
  int build_fused_detect;                 /* 1 = the last build kernel also compared its tiles with the snapshot (stage b fused into stage a) */
  double proof_kernel_ms;                 /* dim > 1: full symmetry proof on the compacted graph (0 when the proof was incremental or in-kernel) */
  int build_deferred;                     /* 1 = the last set_matrix_dev returned with its work in flight and was completed by the next call */
  int expand_speculative;                 /* 1 = the last compute_permutation_dev expanded before the flags were read and the result stood */
  int morton_passes;                      /* radix passes the last set_coordinates ran (digits that are constant over all keys are skipped) */
  double morton_ms;                       /* last set_coordinates: bounding box + keys + radix sort + rank */
  double supernodes_ms;                   /* last parth_b200_supernodes* call */
} parth_b200_stats;
int parth_b200_get_stats(parth_b200 *h, parth_b200_stats *out);
int parth_b200_synchronize(parth_b200 *h);
/* Page-locks / releases a host array of the CALLER (cudaHostRegister): the host-pointer entry points then move it at the
 * PCIe rate (255 MB in 4.8 ms instead of ~20 ms from pageable memory at 200^3).  For callers that keep their Ap / Ai /
 * perm arrays across updates -- what the reference's users do (LinSysSolver.hpp keeps the matrix arrays alive) -- and
 * do not link the CUDA runtime themselves.  The array must stay allocated until it is unpinned. */
int parth_b200_pin_host(void *ptr, size_t bytes);
int parth_b200_unpin_host(void *ptr);
/* the stream all work of this handle runs on (cudaStream_t) */
void *parth_b200_stream(parth_b200 *h);

/* Multi-GPU, one process per GPU, every rank holding the same graph and tree.  Independent dirty
 * separator-tree nodes are the unit that shards: after
 * parth_b200_set_shard(h, rank, world) with world > 1, the nodes an update has to re-order
 * (Integrator::permuteFineHMDNodes, Integrator.cpp:851-893) are dealt to the ranks largest-first
 * onto the least loaded rank -- the same plan on every rank -- and compute_permutation orders only
 * this rank's share, then returns PARTH_B200_NEEDS_EXCHANGE without expanding.  The caller gathers
 * the per-node label slices over NCCL (pack -> all-gather -> unpack; parth_b200/sharded.py) and
 * calls parth_b200_finish_permutation*.  Nothing else of the update needs a collective: build,
 * detection and the dirty sets are computed redundantly and identically on every rank. */
int parth_b200_set_shard(parth_b200 *h, int rank, int world);
/* two-call protocol: nodes == NULL returns the count.  owner[i] = rank that ordered nodes[i],
 * sizes[i] = its number of DOFs (= length of its label slice) */
int parth_b200_exchange_plan(parth_b200 *h, int *n_nodes, int *nodes, int *owner, int *sizes);
/* labels of the listed nodes, concatenated in list order, to / from a device buffer; unpack also
 * re-scatters the nodes' slices of mesh_perm (mesh_perm[label + offset] = DOF) */
int parth_b200_pack_labels_dev(parth_b200 *h, const int *nodes, int n, int *dbuf);
int parth_b200_unpack_labels_dev(parth_b200 *h, const int *nodes, int n, const int *dbuf);
/* expansion + snapshot that compute_permutation skipped; perm / dperm hold M_n*dim ints */
int parth_b200_finish_permutation(parth_b200 *h, int *perm, int dim);
int parth_b200_finish_permutation_dev(parth_b200 *h, int *dperm, int dim);

/* ONE mesh over several GPUs: row blocks of the graph build and of change detection (SURVEY.md section 8e rows 1-2;
 * reference sites ParthAPI.cpp:366-418, Integrator.cpp:259-338), one process per GPU, the collectives inside the
 * library (NCCL, loaded with dlopen when a communicator is created -- a single-GPU caller never needs it).
 *
 *   rank 0: parth_b200_comm_unique_id(id)  -> send the 128 bytes to every rank (file, MPI, a socket ...)
 *   all   : parth_b200_comm_init(h, id, rank, world)
 *           parth_b200_import_tree(h, ..., NULL, NULL)      the replicated separator tree (a fixture)
 *           parth_b200_set_matrix*(h, ...)                  first matrix: builds this rank's rows, proves symmetry
 *           parth_b200_accept_snapshot(h)                   ... which the tree belongs to
 *   per update, all ranks: parth_b200_set_matrix*(h, ...); parth_b200_compute_permutation*(h, perm, dim)
 *
 * With a communicator set, set_matrix / set_matrix_dev / set_matrix_block* read ONLY this rank's block of the
 * arrays -- Ap[c0 .. c1] and Ai[Ap[c0] .. Ap[c1]), global indexing, [c0, c1) from parth_b200_block_range -- so a
 * host caller uploads 1/world of the matrix over its own PCIe link and a caller that holds only its block passes
 * pointers biased by -c0 / -Ap[c0].  Detection exchanges one small slot per rank (header, dirty-node flags, changed
 * edges, directed row diffs: one all-gather); dirty sub-meshes are assembled by all-reduce, the graphs of an update
 * are dealt to the ranks and their permutations all-reduced.  Every rank returns the same full permutation.
 * Restrictions of the block path: dim == 1, matrices the pipeline build kernel accepts (sorted symmetric columns
 * with their diagonal), no DOF map, at most 16384 changed rows per rank and update; the stage-wise entry points,
 * get_mesh, symbolic and the coarse sub-mesh extraction need the whole graph and are refused.
 * parth_b200_comm_init_local: the ranks are `world` handles of ONE process, each driven by its own host thread
 * (test plumbing for the sharded logic on a single GPU, or several GPUs of one process). */
int parth_b200_comm_unique_id(void *id128);
int parth_b200_comm_init(parth_b200 *h, const void *id128, int rank, int world);
int parth_b200_comm_init_local(parth_b200 **handles, int world);
int parth_b200_comm_info(parth_b200 *h, int *rank, int *world, const char **kind);
/* columns [c0, c1) this rank owns of an N x N matrix with nnz entries (fixes the partition of the handle) */
int parth_b200_block_range(parth_b200 *h, int N, int nnz, int *c0, int *c1);
/* host arrays, global indexing, only the block is dereferenced; nnz = Ap[N] of the whole matrix */
int parth_b200_set_matrix_block(parth_b200 *h, int N, int nnz, const int *Ap, const int *Ai, int dim);
/* device arrays, global indexing, only the block is dereferenced; q0 = Ap[c0], q1 = Ap[c1] (no read-back) */
int parth_b200_set_matrix_block_dev(parth_b200 *h, int N, int nnz, const int *dAp, const int *dAi, int dim, int q0,
                                    int q1);
/* the update of a sharded mesh with a distributed result: as parth_b200_compute_permutation, but this rank's host
 * receives only its block of the expanded permutation, perm[c0 * dim .. c1 * dim) with [c0, c1) from
 * parth_b200_block_range (first / count return that range; a caller that wants the permutation whole on one host
 * gathers the blocks).  The device keeps the whole permutation on every rank. */
int parth_b200_compute_permutation_block(parth_b200 *h, int *perm_block, int dim, int *first, int *count);

#ifdef __cplusplus
}
#endif
#endif
