// parth_b200/csrc/capi_todo.cu -- entry points not implemented yet (shrinks as stages land).
#include "handle.cuh"
#define TODO(h) do { if (h) (h)->err = std::string(__func__) + ": not implemented yet"; return PARTH_B200_ERR_UNSUPPORTED; } while (0)
extern "C" {
int parth_b200_set_new_to_old_map(parth_b200 *h, const int *map, int len) { if (h && len == 0) { h->map_len = 0; return 0; } TODO(h); }
int parth_b200_import_tree(parth_b200 *h, int, int, int, const int *, const int *, const int *, const int *, const int *, const int *) { TODO(h); }
int parth_b200_tree_dims(parth_b200 *h, int *a, int *b, int *c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return PARTH_B200_ERR_NO_TREE; }
int parth_b200_export_tree(parth_b200 *h, int *, int *, int *, int *, int *, int *) { TODO(h); }
int parth_b200_compute_permutation(parth_b200 *h, int *, int) { TODO(h); }
int parth_b200_compute_permutation_dev(parth_b200 *h, int *, int) { TODO(h); }
int parth_b200_map_mesh_perm_to_matrix_perm(parth_b200 *h, const int *, int, int *, int) { TODO(h); }
int parth_b200_map_mesh_perm_to_matrix_perm_dev(parth_b200 *h, const int *, int, int *, int) { TODO(h); }
double parth_b200_get_reuse(parth_b200 *h) { return h ? h->reuse : 0; }
int parth_b200_get_num_changes(parth_b200 *h) { return h ? h->n_changed : 0; }
int parth_b200_get_changed_edges(parth_b200 *h, int *, int) { TODO(h); }
int parth_b200_get_dirty_nodes(parth_b200 *h, int *, int *, int *, int *) { TODO(h); }
int parth_b200_get_timers(parth_b200 *h, double t[5]) { if (!h) return -1; for (int i = 0; i < 5; i++) t[i] = h->timers[i]; return 0; }
int parth_b200_get_mesh_perm(parth_b200 *h, int *, int) { TODO(h); }
int parth_b200_mesh_perm_dev(parth_b200 *h, const int **) { TODO(h); }
int parth_b200_stage_integrate(parth_b200 *h) { TODO(h); }
int parth_b200_stage_detect(parth_b200 *h, int *) { TODO(h); }
int parth_b200_stage_permute_fine(parth_b200 *h, const int *, int) { TODO(h); }
int parth_b200_stage_submesh(parth_b200 *h, int, int, int *, int *, int *, int *, int *) { TODO(h); }
int parth_b200_amd_batch(parth_b200 *h, int, const int *, const int *, const int *, int *) { TODO(h); }
int parth_b200_symbolic(parth_b200 *h, const int *, int *, int *, int64_t *) { TODO(h); }
}
