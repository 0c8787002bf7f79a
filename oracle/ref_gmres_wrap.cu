// oracle/ref_gmres_wrap.cu -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/libref_gmres.so).
//
// extern "C" doors onto the reference's OWN solver translation units, compiled UNCHANGED from
// /root/reference/src by nvcc: gmres.cu (host twins GMRES / GMRESilu, getLevel), preconditioner.cu
// (MyILU0::HostPrecond, MyILUPPfloat::HostPrecond_*), leftILU.cu (generateLevel, leftILU), SpMV_kernel.cu,
// plus SpMV_compute.cpp / iluk.cpp / itsol.cpp.  Headers the reference needs and this image lacks
// (helper_cuda.h, CUSP, the legacy cuSPARSE declarations) are stand-ins under oracle/shim/; legacy
// cuBLAS/cuSPARSE entry points are link-time stubs that abort (ref_stubs.cpp) -- only HOST paths are
// called from here, except ref_leftILU, which needs a GPU exactly like the reference does.
//
// The preconditioner objects are filled through their data members instead of Initilize(), because the
// reference's Initilize() uploads to a GPU and runs cuSPARSE analysis (src/preconditioner.cu:288-362,
// 1205-1334); the host twins only read the host arrays set here.
#define private public   // MyILUPPfloat keeps its host arrays private; layout is unaffected
#include "gmres.h"
#include "leftILU.h"
#undef private
#include <cstring>
#include <vector>

// In src/ the split-preconditioner hooks are pure virtual in Preconditioner and MyILU0 / MyNONE do not override
// them (only the thermal tree instantiates those classes), so give them empty overriders: GMRES() calls
// HostPrecond() only.
#define BG_EMPTY_SPLIT_HOOKS                                                              \
  void HostPrecond_rhs(const ValueType*, ValueType*) {}                                    \
  void HostPrecond_right(const ValueType*, ValueType*) {}                                  \
  void HostPrecond_left(const ValueType*, ValueType*) {}                                   \
  void HostPrecond_starting_value(const ValueType*, ValueType*) {}                         \
  void DevPrecond_rhs(float*, float*) {}                                                   \
  void DevPrecond_right(float*, float*) {}                                                 \
  void DevPrecond_left(float*, float*) {}                                                  \
  void DevPrecond_starting_value(float*, float*) {}
struct RefILU0 : MyILU0 { BG_EMPTY_SPLIT_HOOKS };
struct RefNONE : MyNONE { BG_EMPTY_SPLIT_HOOKS };

extern "C" {

// GMRES(val,rowIndices,indices,x,b,n,m,max_iter,tol,MyILU0)  -- src/gmres.cu:1904-2066 with
// MyILU0::HostPrecond (src/preconditioner.cu:205-251)
int ref_gmres_ilu0(int n, const float* val, const int* rp, const int* ci, float* lv, int* lp, int* li, float* uv,
                   int* up, int* ui, float* x, const float* b, int m, int* max_iter, float* tol) {
  RefILU0* M = (RefILU0*)operator new(sizeof(RefILU0));  // members are borrowed: never run the destructor
  new (M) RefILU0;
  M->numRows = n;
  M->l_val = lv; M->l_rowIndices = lp; M->l_indices = li;
  M->u_val = uv; M->u_rowIndices = up; M->u_indices = ui;
  const int flag = GMRES(val, rp, ci, x, b, n, m, max_iter, tol, *M);
  operator delete(M);  // members are borrowed: skip the destructor
  return flag;
}
// same loop without preconditioning (MyNONE)
int ref_gmres_none(int n, const float* val, const int* rp, const int* ci, float* x, const float* b, int m, int* max_iter,
                   float* tol) {
  RefNONE M;
  M.numRows = n;
  return GMRES(val, rp, ci, x, b, n, m, max_iter, tol, M);
}
// M^-1 application alone
void ref_ilu0_hostprecond(int n, float* lv, int* lp, int* li, float* uv, int* up, int* ui, const float* in, float* out) {
  RefILU0* M = (RefILU0*)operator new(sizeof(RefILU0));
  new (M) RefILU0;
  M->numRows = n;
  M->l_val = lv; M->l_rowIndices = lp; M->l_indices = li;
  M->u_val = uv; M->u_rowIndices = up; M->u_indices = ui;
  M->HostPrecond(in, out);
  operator delete(M);
}

// GMRESilu (src/gmres.cu:2069-2252) with MyILUPPfloat host twins (src/preconditioner.cu:1350-1416):
// split preconditioning on ILU++ factors (double) with float vectors
static MyILUPPfloat* make_split(int n, double* lv, int* lp, int* li, double* uv, int* up, int* ui, float* middle,
                                int* permRow, int* permCol, float* lscale, float* rscale, float* tmp) {
  MyILUPPfloat* M = (MyILUPPfloat*)operator new(sizeof(MyILUPPfloat));
  std::memset((void*)M, 0, sizeof(MyILUPPfloat));
  new (M) MyILUPPfloat;
  M->numRows = n;
  M->l_val = lv; M->l_rowIndices = lp; M->l_indices = li;
  M->u_val = uv; M->u_rowIndices = up; M->u_indices = ui;
  M->middle_val = middle;
  M->permRow_indices = permRow; M->permCol_indices = permCol;
  M->lscale_val = lscale; M->rscale_val = rscale;
  M->tmpvector = tmp;
  return M;
}
int ref_gmresilu_split(int n, const float* val, const int* rp, const int* ci, double* lv, int* lp, int* li, double* uv,
                       int* up, int* ui, float* middle, int* permRow, int* permCol, float* lscale, float* rscale,
                       float* x, const float* b, int m, int* max_iter, float* tol) {
  std::vector<float> tmp(n);
  MyILUPPfloat* M = make_split(n, lv, lp, li, uv, up, ui, middle, permRow, permCol, lscale, rscale, tmp.data());
  const int flag = GMRESilu(val, rp, ci, x, b, n, m, max_iter, tol, *M);
  operator delete(M);
  return flag;
}
// which: 1 left, 2 right, 3 starting_value, 4 rhs
void ref_split_hostprecond(int which, int n, double* lv, int* lp, int* li, double* uv, int* up, int* ui, float* middle,
                           int* permRow, int* permCol, float* lscale, float* rscale, const float* in, float* out) {
  std::vector<float> tmp(n);
  MyILUPPfloat* M = make_split(n, lv, lp, li, uv, up, ui, middle, permRow, permCol, lscale, rscale, tmp.data());
  if (which == 1) M->HostPrecond_left(in, out);
  else if (which == 2) M->HostPrecond_right(in, out);
  else if (which == 3) M->HostPrecond_starting_value(in, out);
  else M->HostPrecond_rhs(in, out);
  operator delete(M);
}

// getLevel (src/gmres.cu:1033-1077), generateLevel (src/leftILU.cu:339-368)
void ref_getLevel(int is_lower, const float* val, const int* rp, const int* ci, int* level, int n) {
  getLevel(is_lower != 0, val, rp, ci, level, n);
}
void ref_generateLevel(int n, const float* val, const int* rp, const int* ci, int* node, int* level) {
  std::vector<int> nd, lv;
  generateLevel(n, val, rp, ci, nd, lv);
  for (int i = 0; i < n; ++i) { node[i] = nd[i]; level[i] = lv[i]; }
}
// leftILU (src/leftILU.cu:27-336): needs a CUDA device, exactly like the reference.  Returns nnzL, nnzU through
// the last two arguments; the six arrays are malloc'ed by the reference and copied out here.
int ref_leftILU(int n, float* val, int* rp, int* ci, float* lv, int* lp, int* li, float* uv, int* up, int* ui) {
  float *l_val = 0, *u_val = 0;
  int *l_rp = 0, *l_ci = 0, *u_rp = 0, *u_ci = 0;
  leftILU(n, val, rp, ci, l_val, l_rp, l_ci, u_val, u_rp, u_ci);
  std::memcpy(lp, l_rp, sizeof(int) * (n + 1)); std::memcpy(up, u_rp, sizeof(int) * (n + 1));
  std::memcpy(li, l_ci, sizeof(int) * l_rp[n]); std::memcpy(lv, l_val, sizeof(float) * l_rp[n]);
  std::memcpy(ui, u_ci, sizeof(int) * u_rp[n]); std::memcpy(uv, u_val, sizeof(float) * u_rp[n]);
  return 0;
}
}
