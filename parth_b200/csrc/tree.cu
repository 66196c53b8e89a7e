// parth_b200/csrc/tree.cu -- kernels over the flat separator tree: permutation assembly (e),
// dim expansion (e), DOF->node map, node-pair classification for the dirty sets (b).
//
// Flat layout instead of the reference's per-node std::vectors (HMD.h:18-104): dofs[] / labels[]
// are the concatenation of HMDNode::DOFs / permuted_new_label in heap order, node_ptr[] delimits
// the nodes, meta[node*7 + {0..6}] = left, right, node_id, parent, offset, level, org_size.
#include "tree.cuh"

namespace pb {

// last node i with node_ptr[i] <= k (empty nodes have zero-length slices and are skipped)
__device__ __forceinline__ int node_of_entry(const int *__restrict__ node_ptr, int nodes, int k) {
  int lo = 0, hi = nodes;  // first i with node_ptr[i] > k
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(node_ptr + mid) <= k) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

// HMD::assignCoarseLocalPermToGlobal (HMD.cpp:348-363) as one flat scatter:
// mesh_perm[labels[k] + offset(node)] = dofs[k] for every entry of every node the recursion from
// the root reaches (visit[node] != 0).
__global__ void assemble_all_kernel(const int *__restrict__ node_ptr, const int *__restrict__ meta,
                                    const int *__restrict__ visit, int nodes,
                                    const int *__restrict__ dofs, const int *__restrict__ labels,
                                    int total, int *__restrict__ mesh_perm) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  int nd = node_of_entry(node_ptr, nodes, k);
  if (!__ldg(visit + nd)) return;
  mesh_perm[__ldg(labels + k) + __ldg(meta + nd * 7 + 4)] = __ldg(dofs + k);
}

// same for an explicit node list (Integrator::permuteFineHMDNodes tail, Integrator.cpp:886-891);
// one CTA per listed node
__global__ void assemble_nodes_kernel(const int *__restrict__ list, const int *__restrict__ node_ptr,
                                      const int *__restrict__ meta, const int *__restrict__ dofs,
                                      const int *__restrict__ labels, int *__restrict__ mesh_perm) {
  const int nd = list[blockIdx.x];
  const int a = node_ptr[nd], b = node_ptr[nd + 1], off = meta[nd * 7 + 4];
  for (int k = a + threadIdx.x; k < b; k += blockDim.x) mesh_perm[labels[k] + off] = dofs[k];
}

__global__ void dof2node_kernel(const int *__restrict__ node_ptr, int nodes,
                                const int *__restrict__ dofs, int total, int *__restrict__ dof2node) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  dof2node[__ldg(dofs + k)] = node_of_entry(node_ptr, nodes, k);
}

// ParthAPI::mapMeshPermToMatrixPerm (ParthAPI.cpp:313-318): out[i*dim + d] = mesh_perm[i]*dim + d.
// One thread per OUTPUT element so stores are fully coalesced; the mesh_perm re-reads hit L1.
template <int DIM>
__global__ void expand_kernel(const int *__restrict__ mesh_perm, long long n_out, int dim,
                              int *__restrict__ out) {
  long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  if (DIM == 1) { out[o] = __ldg(mesh_perm + o); return; }
  const int d = DIM > 0 ? DIM : dim;
  long long i = o / d;
  int r = (int)(o - i * d);
  out[o] = __ldg(mesh_perm + i) * d + r;
}

// vectorised form: one thread expands 4 consecutive mesh indices into DIM 128-bit stores
// (4*DIM contiguous ints, 16-byte aligned), so every store instruction moves full vectors
template <int DIM>
__global__ void expand_vec_kernel(const int4 *__restrict__ mesh_perm4, int n4, int4 *__restrict__ out4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int4 v = ld_stream4(mesh_perm4 + i);
  const int src[4] = {v.x, v.y, v.z, v.w};
  int o[4 * DIM];
#pragma unroll
  for (int k = 0; k < 4 * DIM; k++) o[k] = src[k / DIM] * DIM + (k % DIM);
#pragma unroll
  for (int q = 0; q < DIM; q++)
    st_stream4(out4 + (size_t)i * DIM + q, make_int4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
}

// Integrator::findRegionConnections + first loop of filterChanges (Integrator.cpp:685-747):
// every changed DOF edge marks either its node (same node) or the LCA of its two nodes.
__global__ void mark_dirty_nodes_kernel(const int *__restrict__ edges, int n_edges,
                                        const int *__restrict__ dof2node,
                                        const int *__restrict__ meta, int *__restrict__ within_flag,
                                        int *__restrict__ lca_flag) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int na = __ldg(dof2node + edges[2 * e]), nb = __ldg(dof2node + edges[2 * e + 1]);
  int big = na, small = nb;
  if (small > big) { int t = big; big = small; small = t; }
  if (small == big) { within_flag[small] = 1; return; }
  if (is_separator_dev(small, big)) return;
  lca_flag[common_separator_dev(meta, small, big)] = 1;
}

__global__ void mark_dirty_nodes_dev_kernel(const int *__restrict__ edges, const int *__restrict__ n_edges_dev, int cap,
                                            const int *__restrict__ dof2node, const int *__restrict__ meta,
                                            int *__restrict__ within_flag, int *__restrict__ lca_flag,
                                            int *__restrict__ count_out) {
  const int n_all = *n_edges_dev;
  if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = n_all;
  const int n_edges = n_all < cap ? n_all : cap;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    int na = __ldg(dof2node + edges[2 * e]), nb = __ldg(dof2node + edges[2 * e + 1]);
    int big = na, small = nb;
    if (small > big) { int t = big; big = small; small = t; }
    if (small == big) { within_flag[small] = 1; continue; }
    if (is_separator_dev(small, big)) continue;
    lca_flag[common_separator_dev(meta, small, big)] = 1;
  }
}

// ------------------------------------------------------------------ host launchers
int launch_assemble_all(const int *node_ptr, const int *meta, const int *visit, int nodes,
                        const int *dofs, const int *labels, int total, int *mesh_perm,
                        cudaStream_t s) {
  if (total <= 0) return 0;
  assemble_all_kernel<<<(total + 255) / 256, 256, 0, s>>>(node_ptr, meta, visit, nodes, dofs, labels,
                                                          total, mesh_perm);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_assemble_nodes(const int *list, int n_list, const int *node_ptr, const int *meta,
                          const int *dofs, const int *labels, int *mesh_perm, cudaStream_t s) {
  if (n_list <= 0) return 0;
  assemble_nodes_kernel<<<n_list, 256, 0, s>>>(list, node_ptr, meta, dofs, labels, mesh_perm);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_dof2node(const int *node_ptr, int nodes, const int *dofs, int total, int *dof2node,
                    cudaStream_t s) {
  if (total <= 0) return 0;
  dof2node_kernel<<<(total + 255) / 256, 256, 0, s>>>(node_ptr, nodes, dofs, total, dof2node);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_expand(const int *mesh_perm, int M_n, int dim, int *out, cudaStream_t s) {
  long long n_out = (long long)M_n * dim;
  if (n_out <= 0) return 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(mesh_perm) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (aligned && dim >= 1 && dim <= 4 && M_n >= 4) {
    const int n4 = M_n / 4;
    const unsigned vb = (unsigned)((n4 + 255) / 256);
    const int4 *in4 = reinterpret_cast<const int4 *>(mesh_perm);
    int4 *out4 = reinterpret_cast<int4 *>(out);
    if (dim == 1) expand_vec_kernel<1><<<vb, 256, 0, s>>>(in4, n4, out4);
    else if (dim == 2) expand_vec_kernel<2><<<vb, 256, 0, s>>>(in4, n4, out4);
    else if (dim == 3) expand_vec_kernel<3><<<vb, 256, 0, s>>>(in4, n4, out4);
    else expand_vec_kernel<4><<<vb, 256, 0, s>>>(in4, n4, out4);
    PB_CUDA(cudaGetLastError());
    const int rest = M_n - 4 * n4;  // at most 3 mesh indices left: the scalar kernel on the tail
    if (rest > 0) {
      expand_kernel<0><<<1, 32, 0, s>>>(mesh_perm + 4 * n4, (long long)rest * dim, dim, out + (long long)4 * n4 * dim);
      PB_CUDA(cudaGetLastError());
      return 2;
    }
    return 1;
  }
  unsigned blocks = (unsigned)((n_out + 255) / 256);
  if (dim == 1) expand_kernel<1><<<blocks, 256, 0, s>>>(mesh_perm, n_out, dim, out);
  else if (dim == 2) expand_kernel<2><<<blocks, 256, 0, s>>>(mesh_perm, n_out, dim, out);
  else if (dim == 3) expand_kernel<3><<<blocks, 256, 0, s>>>(mesh_perm, n_out, dim, out);
  else expand_kernel<0><<<blocks, 256, 0, s>>>(mesh_perm, n_out, dim, out);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_mark_dirty_nodes_dev(const int *edges, const int *n_edges_dev, int cap, const int *dof2node,
                                const int *meta, int *within_flag, int *lca_flag, int *count_out, cudaStream_t s) {
  mark_dirty_nodes_dev_kernel<<<kNumSMsB200, 256, 0, s>>>(edges, n_edges_dev, cap, dof2node, meta, within_flag, lca_flag,
                                                         count_out);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_mark_dirty_nodes(const int *edges, int n_edges, const int *dof2node, const int *meta,
                            int *within_flag, int *lca_flag, cudaStream_t s) {
  if (n_edges <= 0) return 0;
  mark_dirty_nodes_kernel<<<(n_edges + 255) / 256, 256, 0, s>>>(edges, n_edges, dof2node, meta,
                                                                within_flag, lca_flag);
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
