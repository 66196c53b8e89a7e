// oracle/metis_shim.cpp -- TEST INFRASTRUCTURE (glue for the oracle/_ref build only).
// The only METIS in this image is the 64-bit-idx_t static library bundled with the CUDA
// toolkit; the reference calls METIS with 32-bit indices (HMD.cpp:52,54,259).  The
// Makefile renames the library's two entry points to M64_* with objcopy and this file
// widens/narrows the arrays.  It contains no Parth logic.
#include <cstdint>
#include <vector>
extern "C" {
int M64_NodeND(int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *);
int M64_ComputeVertexSeparator(int64_t *, int64_t *, int64_t *, int64_t *, int64_t *, int64_t *,
                               int64_t *);

int METIS_NodeND(int32_t *n, int32_t *xadj, int32_t *adj, int32_t *, int32_t *, int32_t *perm,
                 int32_t *iperm) {
  int64_t N = *n;
  std::vector<int64_t> X(xadj, xadj + N + 1), A(adj, adj + xadj[N]), P(N), IP(N);
  int r = M64_NodeND(&N, X.data(), A.data(), nullptr, nullptr, P.data(), IP.data());
  for (int64_t i = 0; i < N; i++) {
    perm[i] = (int32_t)P[i];
    iperm[i] = (int32_t)IP[i];
  }
  return r;
}

int METIS_ComputeVertexSeparator(int32_t *n, int32_t *xadj, int32_t *adj, int32_t *vw, int32_t *,
                                 int32_t *sep, int32_t *part) {
  int64_t N = *n, S = 0;
  std::vector<int64_t> X(xadj, xadj + N + 1), A(adj, adj + xadj[N]), W(N), P(N);
  for (int64_t i = 0; i < N; i++) W[i] = vw ? vw[i] : 1;
  int r = M64_ComputeVertexSeparator(&N, X.data(), A.data(), W.data(), nullptr, &S, P.data());
  *sep = (int32_t)S;
  for (int64_t i = 0; i < N; i++) part[i] = (int32_t)P[i];
  return r;
}
}
