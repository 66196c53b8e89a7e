/* oracle/inc/metis.h -- TEST INFRASTRUCTURE. Minimal 32-bit METIS prototypes so the
 * reference (src/Parth/HMD.cpp:9,52,54,259) compiles against the 64-bit METIS that
 * ships inside the CUDA toolkit; see metis_shim.cpp. */
#ifndef ORACLE_METIS_H
#define ORACLE_METIS_H
#include <stdint.h>
typedef int32_t idx_t;
typedef float real_t;
#define METIS_OK 1
#ifdef __cplusplus
extern "C" {
#endif
int METIS_NodeND(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *options,
                 idx_t *perm, idx_t *iperm);
int METIS_ComputeVertexSeparator(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                                 idx_t *options, idx_t *sepsize, idx_t *part);
#ifdef __cplusplus
}
#endif
#endif
