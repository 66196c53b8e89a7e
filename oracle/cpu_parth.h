/*
 * oracle/cpu_parth.h -- TEST INFRASTRUCTURE.
 *
 * One C ABI, two CPU implementations of Parth's dynamic-reuse ordering path:
 *   oracle/_ref/libparth_ref.so     kind "reference": the UNMODIFIED reference sources compiled
 *                                   where they lie under /root/reference (oracle/Makefile), driven
 *                                   through ref_driver.cpp.
 *   oracle/liboracle_parth.so       kind "port": the plain-C restatement (parth_oracle.c,
 *                                   amd_restated.c, symbolic_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load either library.
 *
 * Tree fixture layout (flat int32, shared with include/parth_b200.h):
 *   meta[node*7 + {0..6}] = left, right, node_id, parent, offset, level, org_size  (heap order,
 *   children of i are 2i+1 / 2i+2, empty node <=> node_id == -1; reference: HMD.h:18-27)
 *   node_ptr[nodes+1], dofs[M_n] (per node ascending global ids), labels[M_n]
 *   (labels[k] = position of dofs[k] inside its node's slice of the permutation).
 */
#ifndef CPU_PARTH_H
#define CPU_PARTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpu_parth cpu_parth;

#define CPU_PARTH_OK 0
#define CPU_PARTH_ERR (-1)
#define CPU_PARTH_NEEDS_REDISSECT (-10) /* port only: dirty coarse nodes need the METIS separator call */
#define CPU_PARTH_NO_TREE (-11)         /* port only: cold start needs METIS */
#define CPU_PARTH_UNSUPPORTED (-12)

const char *cpu_parth_kind(void);
cpu_parth *cpu_parth_create(void);
void cpu_parth_destroy(cpu_parth *h);
void cpu_parth_set_options(cpu_parth *h, int reorder_type, int nd_levels, int lagging,
                           int aggressive, int relocate_lvl, int lag_lvl, int verbose);

/* map == NULL or map_len == 0: no new-to-old DOF map */
int cpu_parth_set_matrix(cpu_parth *h, int N, const int *Ap, const int *Ai, int dim,
                         const int *map, int map_len);
int cpu_parth_set_mesh(cpu_parth *h, int n, const int *Mp, const int *Mi, const int *map,
                       int map_len);
int cpu_parth_mesh_dims(cpu_parth *h, int *M_n, int *nnz);
int cpu_parth_mesh_copy(cpu_parth *h, int *Mp, int *Mi);

/* returns number of ints written (M_n*dim) or a negative status */
int cpu_parth_compute_permutation(cpu_parth *h, int dim, int *perm_out, int cap);
double cpu_parth_get_reuse(cpu_parth *h);
int cpu_parth_get_num_changes(cpu_parth *h);
int cpu_parth_get_changed_edges(cpu_parth *h, int *pairs, int cap_pairs);
int cpu_parth_get_dirty(cpu_parth *h, int *fine, int *nfine, int *coarse, int *ncoarse);
int cpu_parth_get_timers(cpu_parth *h, double t[5]);
int cpu_parth_get_mesh_perm(cpu_parth *h, int *mesh_perm, int cap);

int cpu_parth_tree_dims(cpu_parth *h, int *M_n, int *levels, int *nodes);
int cpu_parth_tree_export(cpu_parth *h, int *meta, int *node_ptr, int *dofs, int *labels,
                          int *dof2node, int *mesh_perm);
int cpu_parth_tree_import(cpu_parth *h, int M_n, int levels, int nodes, const int *meta,
                          const int *node_ptr, const int *dofs, const int *labels,
                          const int *Mp_prev, const int *Mi_prev);

/* stage-wise entry points on the current mesh + tree */
int cpu_parth_integrate(cpu_parth *h);  /* Integrator::integrateDOFChanges */
int cpu_parth_detect(cpu_parth *h);     /* Integrator::computeChangedEdges -> count */
int cpu_parth_permute_fine(cpu_parth *h, const int *nodes, int n); /* permuteFineHMDNodes */
/* coarse == 0: createMultiSubMesh for one node; coarse == 1: getCoarseMesh of the sub-tree.
 * Two-call protocol: pass NULL arrays to get sizes. */
int cpu_parth_submesh(cpu_parth *h, int node, int coarse, int *n, int *nnz, int *Mp, int *Mi,
                      int *local_to_global);
int cpu_parth_map_mesh_perm(cpu_parth *h, const int *mesh_perm, int M_n, int dim, int *out);

/* free functions */
int cpu_parth_amd(int n, const int *Ap, const int *Ai, int *P);
/* etree of P*A*P' where (Ap,Ai) is the full symmetric pattern (diagonal optional) and
 * perm[new] = old.  parent[j] == -1 for roots. */
int cpu_parth_etree(int n, const int *Ap, const int *Ai, const int *perm, int *parent);
/* column counts of L (diagonal included); returns nnz(L) or -1 if unsupported */
int64_t cpu_parth_colcount(int n, const int *Ap, const int *Ai, const int *perm,
                           const int *parent, int *colcount);
/* fundamental supernodes from (parent, colcount): port only (see symbolic_oracle.c); returns nsuper */
int cpu_parth_supernodes(int n, const int *parent, const int *colcount, int *super, int *supermap,
                         int *sparent, int64_t *sizes);

#ifdef __cplusplus
}
#endif
#endif
