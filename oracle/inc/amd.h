/* oracle/inc/amd.h -- TEST INFRASTRUCTURE. Prototype of the one SuiteSparse-AMD entry
 * point the reference calls (src/Parth/HMD.cpp:6,74); implemented by amd_restated.c. */
#ifndef ORACLE_AMD_H
#define ORACLE_AMD_H
#ifdef __cplusplus
extern "C" {
#endif
int amd_order(int n, const int *Ap, const int *Ai, int *P, double *Control, double *Info);
#ifdef __cplusplus
}
#endif
#endif
