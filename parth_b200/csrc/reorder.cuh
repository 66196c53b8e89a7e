// parth_b200/csrc/reorder.cuh -- launchers of extract.cu / amd.cu (stage c).
#pragma once
#include "common.cuh"

namespace pb {

// ---- extract.cu: sub-mesh extraction
// Fine nodes (HMD::createMultiSubMesh, HMD.cpp:538-612), node-list driven: the concatenated
// vertex space holds the DOFs of list[0], list[1], ... in node order; vtx_ptr[g] delimits them.
int launch_extract_count(const int *list, const int *vtx_ptr, int count, int total,
                         const int *node_ptr, const int *dofs, const int *Mp, const int *Mi,
                         const int *dof2node, int *cnt, cudaStream_t s, int row0 = 0, int row1 = 0x7fffffff);
int launch_extract_fill(const int *list, const int *vtx_ptr, int count, int total,
                        const int *node_ptr, const int *dofs, const int *Mp, const int *Mi,
                        const int *dof2node, const int *G, int *sub_Mi, cudaStream_t s, int row0 = 0,
                        int row1 = 0x7fffffff);
// (row0, row1: row-block mode -- only the rows [row0, row1) of the graph are local; the other vertices of the
// batch get count 0 / no entries and arrive by all-reduce, stage_sharded.cu)
// Marker driven (HMD::createSubMesh via getCoarseMesh, HMD.cpp:465-523): chosen[] is 0/1 per DOF,
// g2l its exclusive scan, l2g the ascending list of chosen DOFs.
int launch_mark_nodes(const int *list, int n_list, const int *node_ptr, const int *dofs, int *chosen,
                      cudaStream_t s);
int launch_compact_chosen(const int *chosen, const int *g2l, int M, int *l2g, cudaStream_t s);
int launch_marker_count(const int *l2g, int n, const int *Mp, const int *Mi, const int *chosen,
                        int *cnt, cudaStream_t s);
int launch_marker_fill(const int *l2g, int n, const int *Mp, const int *Mi, const int *chosen,
                       const int *g2l, const int *G, int *sub_Mi, cudaStream_t s);
// HMDNode::setPermutedNewLabel + the scatter of Integrator::permuteFineHMDNodes
// (HMD.h:70-79, Integrator.cpp:886-891): labels[perm[j]] = j; mesh_perm[j + offset] = dofs[perm[j]]
int launch_apply_order(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                       const int *meta, const int *dofs, const int *perm, int *labels, int *mesh_perm,
                       cudaStream_t s);

// label slices of the listed nodes <-> one concatenated buffer (multi-GPU exchange)
int launch_pack_labels(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                       const int *labels, int *buf, cudaStream_t s);
int launch_unpack_labels(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                         const int *meta, const int *dofs, const int *buf, int *labels, int *mesh_perm,
                         cudaStream_t s);

// ---- amd.cu
// glist != nullptr: only the n_list graphs glist[0..n_list) of the batch are ordered (this rank's share)
int launch_amd_batch(int count, const int *vtx_ptr, const int *G, const int *sub_Mi,
                     const long long *ws_ptr, int *ws, int *perm_out, long long max_stage_ints, bool any_unstaged,
                     cudaStream_t s, const int *glist = nullptr, int n_list = 0);
long long amd_slab_ints(int n, int nnz);
int amd_stage_capacity_ints();                  // ints of shared memory one CTA can stage a graph in
long long amd_stage_need_ints(int n, int nnz);  // ints a (symmetric) graph needs to be staged

}  // namespace pb
