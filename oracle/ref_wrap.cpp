// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/libref.so).
// extern "C" doors onto the reference's own CPU translation units, compiled UNCHANGED from
// /root/reference/src: SpMV_compute.cpp (computeSpMV, LUSolve_ignoreZero), formatConvert.cpp
// (coo2csr_in, coo2csrDouble_in, LDcsc2csrMySpMatrix), partition.cpp (partition).  Used to pin oracle.cpp and the
// netlist front end's hand-over to the solver.
#include <cstdlib>
#include <cstring>
#include "SpMV.h"
#include "cs.h"
#include "gpuData.h"
void coo2csr_in(int n, int nz, float* a, int* i_idx, int* j_idx);
void coo2csrDouble_in(int n, int nz, double* a, int* i_idx, int* j_idx);
void partition(cs_dl* A, int npart, int nNodes, UF_long* part_size, UF_long* node_part,
               UF_long* pinv, UF_long* q);
void writeCSR(int m, int n, int nnz, int* rowPtr, int* colIdx, float* val, const char* filename);
extern "C" {
void ref_writeCSR(int m, int n, int nnz, int* rowPtr, int* colIdx, float* val, const char* filename) {
  writeCSR(m, n, nnz, rowPtr, colIdx, val, filename);  // src/formatConvert.cpp:425-445
}
void ref_computeSpMV(float* x, const float* val, const int* rp, const int* ci, const float* y, int n) {
  computeSpMV(x, val, rp, ci, y, n);
}
void ref_LUSolve_ignoreZero(float* x, const float* lv, const int* lp, const int* li, const float* uv,
                            const int* up, const int* ui, const float* y, int n) {
  LUSolve_ignoreZero(x, lv, lp, li, uv, up, ui, y, n);
}
void ref_coo2csr_in(int n, int nz, float* a, int* i_idx, int* j_idx) { coo2csr_in(n, nz, a, i_idx, j_idx); }
void ref_coo2csrDouble_in(int n, int nz, double* a, int* i_idx, int* j_idx) {
  coo2csrDouble_in(n, nz, a, i_idx, j_idx);
}
void ref_partition(long m, long* Ap, long* Ai, int npart, int nNodes, long* part_size, long* node_part,
                   long* pinv, long* q) {
  cs_dl A;
  A.nzmax = Ap[m]; A.m = m; A.n = m; A.p = Ap; A.i = Ai; A.x = 0; A.nz = -1;
  partition(&A, npart, nNodes, part_size, node_part, pinv, q);
}
// A.1: the power-grid driver's hand-over of a cs_dl matrix to the GPU solver (src/formatConvert.cpp:300-334):
// column-compressed long/double -> float CSR by cycle-chasing COO->CSR and a per-row bubble sort
void ref_LDcsc2csr(long m, long n, long* p, long* i, double* x, int* rowptr_out, int* col_out, float* val_out) {
  ucr_cs_dl M;
  M.nzmax = p[n]; M.m = m; M.n = n; M.p = p; M.i = i; M.x = x; M.nz = -1;
  MySpMatrix S;
  std::memset(&S, 0, sizeof S);
  LDcsc2csrMySpMatrix(&S, &M);
  const long nnz = p[n];
  for (long k = 0; k <= m; ++k) rowptr_out[k] = S.rowIndices[k];
  for (long k = 0; k < nnz; ++k) { col_out[k] = S.indices[k]; val_out[k] = S.val[k]; }
  free(S.rowIndices); free(S.indices); free(S.val);
}
}
