// parth_b200/csrc/capi_todo.cu -- stage drivers not implemented yet (shrinks as stages land).
#include "stages.cuh"
namespace pb {
#define TODO_STAGE(name) throw StatusError(PARTH_B200_ERR_UNSUPPORTED, name ": not implemented yet")
void stage_set_map(parth_b200 &h, const int *, int len) { if (len == 0) { h.map_len = 0; return; } TODO_STAGE("set_new_to_old_map"); }
void stage_integrate(parth_b200 &) { TODO_STAGE("integrate"); }
void stage_relocate(parth_b200 &, int, std::vector<int> &) { TODO_STAGE("relocate"); }
int stage_coarse(parth_b200 &) { TODO_STAGE("coarse"); }
void stage_symbolic(parth_b200 &, const int *, int *, int *, long long *) { TODO_STAGE("symbolic"); }
}
