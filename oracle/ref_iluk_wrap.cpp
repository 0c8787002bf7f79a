// oracle/ref_iluk_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
// C interface over the ILU(k) the reference vendors from ITSOL-2 (src/iluk.cpp: ilukC / lofC, src/itsol.cpp),
// compiled UNCHANGED from /root/reference into oracle/_ref/libref_iluk.so.  tests/test_host_logic.py pins the fill
// pattern of the product's bg_host_iluk against lofC's for every level, and its values against ilukC's float arithmetic.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gpuData.h"

int ilukC(int lofM, csptr csmat, iluptr lu, FILE* fp);

namespace {
ILUSpar g_lu;
bool g_have = false;
int g_n = 0;
}  // namespace

// CSR (float values) in; returns ilukC's code (0 ok, -2 zero diagonal)
extern "C" int refk_factor(int n, const int* rp, const int* ci, const float* v, int level) {
  SparMat A;
  setupCS(&A, n, 1);
  for (int i = 0; i < n; ++i) {
    const int len = rp[i + 1] - rp[i];
    A.nzcount[i] = len;
    A.ja[i] = (int*)malloc(sizeof(int) * (len > 0 ? len : 1));
    A.ma[i] = (float*)malloc(sizeof(float) * (len > 0 ? len : 1));
    for (int k = 0; k < len; ++k) { A.ja[i][k] = ci[rp[i] + k]; A.ma[i][k] = v[rp[i] + k]; }
  }
  std::memset(&g_lu, 0, sizeof g_lu);
  const int rc = ilukC(level, &A, &g_lu, stderr);
  g_have = rc == 0;
  g_n = n;
  return rc;
}
extern "C" int refk_row_counts(int* nl, int* nu) {  // strictly lower / strictly upper entries per row
  if (!g_have) return -1;
  for (int i = 0; i < g_n; ++i) { nl[i] = g_lu.L->nzcount[i]; nu[i] = g_lu.U->nzcount[i]; }
  return 0;
}
// rows in ITSOL's own order (L ascending by construction, U in discovery order); dinv = the stored reciprocal diagonal
extern "C" int refk_get(int* lcol, float* lval, int* ucol, float* uval, float* dinv) {
  if (!g_have) return -1;
  size_t a = 0, b = 0;
  for (int i = 0; i < g_n; ++i) {
    for (int k = 0; k < g_lu.L->nzcount[i]; ++k) { lcol[a] = g_lu.L->ja[i][k]; lval[a++] = g_lu.L->ma[i][k]; }
    for (int k = 0; k < g_lu.U->nzcount[i]; ++k) { ucol[b] = g_lu.U->ja[i][k]; uval[b++] = g_lu.U->ma[i][k]; }
    dinv[i] = g_lu.D[i];
  }
  return 0;
}
