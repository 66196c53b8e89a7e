// parth_b200/csrc/extract.cu -- stage (c), extraction: sub-meshes of dirty separator-tree nodes.
//
// Replaces HMD::createMultiSubMesh (reference src/Parth/HMD.cpp:538-612) and, for whole
// sub-trees, HMD::getCoarseMesh -> createSubMesh (HMD.cpp:465-523).  The reference scans all M_n
// columns per call; here the fine-node path is driven by the node list, so its traffic is
// 4*(3 + 3*deg) bytes per DIRTY vertex only: DOF id, two row pointers,
// neighbour ids, neighbour node ids, output.
//
// Local id of a DOF = its rank by ascending global id inside the node = its position in the
// node's dofs[] slice (slices are kept ascending by every stage that edits them).
#include "reorder.cuh"

namespace pb {

// last g with vtx_ptr[g] <= v
__device__ __forceinline__ int graph_of_vertex(const int *__restrict__ vtx_ptr, int count, int v) {
  int lo = 0, hi = count;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(vtx_ptr + mid) <= v) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

__global__ void extract_count_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr,
                                     int count, int total, const int *__restrict__ node_ptr,
                                     const int *__restrict__ dofs, const int *__restrict__ Mp,
                                     const int *__restrict__ Mi, const int *__restrict__ dof2node,
                                     int *__restrict__ cnt, int row0, int row1) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int g = graph_of_vertex(vtx_ptr, count, v);
  const int nd = __ldg(list + g);
  const int d = __ldg(dofs + __ldg(node_ptr + nd) + (v - __ldg(vtx_ptr + g)));
  // row-block mode: rows outside [row0, row1) live on another rank (their counts arrive by all-reduce)
  if (d < row0 || d >= row1) { cnt[v] = 0; return; }
  int c = 0;
  for (int p = __ldg(Mp + d); p < __ldg(Mp + d + 1); p++) c += __ldg(dof2node + __ldg(Mi + p)) == nd;
  cnt[v] = c;
}

__global__ void extract_fill_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr,
                                    int count, int total, const int *__restrict__ node_ptr,
                                    const int *__restrict__ dofs, const int *__restrict__ Mp,
                                    const int *__restrict__ Mi, const int *__restrict__ dof2node,
                                    const int *__restrict__ G, int *__restrict__ sub_Mi, int row0, int row1) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int g = graph_of_vertex(vtx_ptr, count, v);
  const int nd = __ldg(list + g);
  const int a = __ldg(node_ptr + nd), n = __ldg(node_ptr + nd + 1) - a;
  const int *slice = dofs + a;
  const int d = __ldg(slice + (v - __ldg(vtx_ptr + g)));
  if (d < row0 || d >= row1) return;
  int o = G[v];
  for (int p = __ldg(Mp + d); p < __ldg(Mp + d + 1); p++) {
    const int nb = __ldg(Mi + p);
    if (__ldg(dof2node + nb) != nd) continue;
    int lo = 0, hi = n;  // position of nb in the ascending slice
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(slice + mid) < nb) lo = mid + 1; else hi = mid;
    }
    sub_Mi[o++] = lo;
  }
}

// one CTA per listed node
__global__ void mark_nodes_kernel(const int *__restrict__ list, const int *__restrict__ node_ptr,
                                  const int *__restrict__ dofs, int *__restrict__ chosen) {
  const int nd = list[blockIdx.x];
  for (int k = node_ptr[nd] + threadIdx.x; k < node_ptr[nd + 1]; k += blockDim.x) chosen[dofs[k]] = 1;
}

__global__ void compact_chosen_kernel(const int *__restrict__ chosen, const int *__restrict__ g2l,
                                      int M, int *__restrict__ l2g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M && chosen[i]) l2g[g2l[i]] = i;
}

__global__ void marker_count_kernel(const int *__restrict__ l2g, int n, const int *__restrict__ Mp,
                                    const int *__restrict__ Mi, const int *__restrict__ chosen,
                                    int *__restrict__ cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int d = l2g[k];
  int c = 0;
  for (int p = Mp[d]; p < Mp[d + 1]; p++) c += __ldg(chosen + Mi[p]);
  cnt[k] = c;
}

__global__ void marker_fill_kernel(const int *__restrict__ l2g, int n, const int *__restrict__ Mp,
                                   const int *__restrict__ Mi, const int *__restrict__ chosen,
                                   const int *__restrict__ g2l, const int *__restrict__ G,
                                   int *__restrict__ sub_Mi) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int d = l2g[k];
  int o = G[k];
  for (int p = Mp[d]; p < Mp[d + 1]; p++) {
    const int nb = Mi[p];
    if (__ldg(chosen + nb)) sub_Mi[o++] = __ldg(g2l + nb);
  }
}

__global__ void apply_order_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr,
                                   int count, int total, const int *__restrict__ node_ptr,
                                   const int *__restrict__ meta, const int *__restrict__ dofs,
                                   const int *__restrict__ perm, int *__restrict__ labels,
                                   int *__restrict__ mesh_perm) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int g = graph_of_vertex(vtx_ptr, count, v);
  const int nd = __ldg(list + g);
  const int a = __ldg(node_ptr + nd);
  const int j = v - __ldg(vtx_ptr + g);
  const int k = perm[v];  // local id of the j-th pivot
  labels[a + k] = j;
  mesh_perm[j + __ldg(meta + nd * 7 + 4)] = __ldg(dofs + a + k);
}

__global__ void pack_labels_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr, int count, int total,
                                   const int *__restrict__ node_ptr, const int *__restrict__ labels,
                                   int *__restrict__ buf) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int g = graph_of_vertex(vtx_ptr, count, v);
  buf[v] = labels[__ldg(node_ptr + __ldg(list + g)) + (v - __ldg(vtx_ptr + g))];
}

__global__ void unpack_labels_kernel(const int *__restrict__ list, const int *__restrict__ vtx_ptr, int count, int total,
                                     const int *__restrict__ node_ptr, const int *__restrict__ meta,
                                     const int *__restrict__ dofs, const int *__restrict__ buf,
                                     int *__restrict__ labels, int *__restrict__ mesh_perm) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const int g = graph_of_vertex(vtx_ptr, count, v);
  const int nd = __ldg(list + g);
  const int k = __ldg(node_ptr + nd) + (v - __ldg(vtx_ptr + g));
  const int lab = buf[v];
  labels[k] = lab;
  mesh_perm[lab + __ldg(meta + nd * 7 + 4)] = __ldg(dofs + k);
}

// ------------------------------------------------------------------ host launchers
static inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

int launch_extract_count(const int *list, const int *vtx_ptr, int count, int total,
                         const int *node_ptr, const int *dofs, const int *Mp, const int *Mi,
                         const int *dof2node, int *cnt, cudaStream_t s, int row0, int row1) {
  if (total <= 0) return 0;
  extract_count_kernel<<<blocks_for(total), 256, 0, s>>>(list, vtx_ptr, count, total, node_ptr, dofs, Mp,
                                                        Mi, dof2node, cnt, row0, row1);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_extract_fill(const int *list, const int *vtx_ptr, int count, int total,
                        const int *node_ptr, const int *dofs, const int *Mp, const int *Mi,
                        const int *dof2node, const int *G, int *sub_Mi, cudaStream_t s, int row0, int row1) {
  if (total <= 0) return 0;
  extract_fill_kernel<<<blocks_for(total), 256, 0, s>>>(list, vtx_ptr, count, total, node_ptr, dofs, Mp,
                                                       Mi, dof2node, G, sub_Mi, row0, row1);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_mark_nodes(const int *list, int n_list, const int *node_ptr, const int *dofs, int *chosen,
                      cudaStream_t s) {
  if (n_list <= 0) return 0;
  mark_nodes_kernel<<<n_list, 256, 0, s>>>(list, node_ptr, dofs, chosen);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_compact_chosen(const int *chosen, const int *g2l, int M, int *l2g, cudaStream_t s) {
  if (M <= 0) return 0;
  compact_chosen_kernel<<<blocks_for(M), 256, 0, s>>>(chosen, g2l, M, l2g);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_marker_count(const int *l2g, int n, const int *Mp, const int *Mi, const int *chosen,
                        int *cnt, cudaStream_t s) {
  if (n <= 0) return 0;
  marker_count_kernel<<<blocks_for(n), 256, 0, s>>>(l2g, n, Mp, Mi, chosen, cnt);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_marker_fill(const int *l2g, int n, const int *Mp, const int *Mi, const int *chosen,
                       const int *g2l, const int *G, int *sub_Mi, cudaStream_t s) {
  if (n <= 0) return 0;
  marker_fill_kernel<<<blocks_for(n), 256, 0, s>>>(l2g, n, Mp, Mi, chosen, g2l, G, sub_Mi);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_pack_labels(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                       const int *labels, int *buf, cudaStream_t s) {
  if (total <= 0) return 0;
  pack_labels_kernel<<<blocks_for(total), 256, 0, s>>>(list, vtx_ptr, count, total, node_ptr, labels, buf);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_unpack_labels(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                         const int *meta, const int *dofs, const int *buf, int *labels, int *mesh_perm,
                         cudaStream_t s) {
  if (total <= 0) return 0;
  unpack_labels_kernel<<<blocks_for(total), 256, 0, s>>>(list, vtx_ptr, count, total, node_ptr, meta, dofs, buf, labels,
                                                        mesh_perm);
  PB_CUDA(cudaGetLastError());
  return 1;
}

int launch_apply_order(const int *list, const int *vtx_ptr, int count, int total, const int *node_ptr,
                       const int *meta, const int *dofs, const int *perm, int *labels, int *mesh_perm,
                       cudaStream_t s) {
  if (total <= 0) return 0;
  apply_order_kernel<<<blocks_for(total), 256, 0, s>>>(list, vtx_ptr, count, total, node_ptr, meta, dofs,
                                                      perm, labels, mesh_perm);
  PB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pb
