// oracle/ilupp_ref.cpp -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/).
//
// Thin C entry points around the reference's vendored, UNMODIFIED ILU++ 1.1.1 headers
// (/root/reference/include/ILU++_1.1.1/lib, compiled from where they lie).  It reproduces the
// factor-producing calls of the power-grid driver
//   src/mna_solve_gpu_gmres.cpp:316-344  (preprocessing + param.init(L,11) + threshold/MEM_FACTOR/MAX_LEVELS)
//   src/mna_solve_gpu_gmres.cpp:396-405  (extract_left/right/middle/perm/scaling)
//   src/mna_solve_gpu_gmres.cpp:96-142   (ILUPPmat2csrMySpMatrixDouble: column-oriented L -> CSR)
// and the CPU baseline solve `iluplusplus::gmres(..., SPLIT, ...)`
//   src/mna_solve_gpu_gmres.cpp:2069.
// No reference source is copied: this file only calls the library's public API.
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#include "iluplusplus.h"

typedef iluplusplus::Real Real;
typedef iluplusplus::matrix_sparse<Real> Matrix;
typedef iluplusplus::vector_dense<Real> Vector;
typedef iluplusplus::multilevelILUCDPPreconditioner<Real, Matrix, Vector> Prec;

struct IluppObj {
  int n = 0;
  Prec P;
  Matrix* A = nullptr;  // column-oriented copy (owned)
  // extracted, converted to row-major CSR (rows sorted by column, like coo2csrDouble_in)
  std::vector<int> lp, li, up, ui, permRow, permCol;
  std::vector<double> lv, uv, middle, lscale, rscale;
  double fill = 0;
};

static void to_csr(const Matrix& M, std::vector<int>& rp, std::vector<int>& ci, std::vector<double>& v) {
  int rows = M.rows(), cols = M.columns();
  int nnz = M.get_pointer(M.get_pointer_size() - 1);
  rp.assign(rows + 1, 0);
  ci.resize(nnz);
  v.resize(nnz);
  if (M.orient() == iluplusplus::ROW) {
    for (int r = 0; r <= rows; ++r) rp[r] = M.get_pointer(r);
    for (int p = 0; p < nnz; ++p) { ci[p] = M.get_index(p); v[p] = M.get_data(p); }
    return;
  }
  // COLUMN oriented: counting transpose.  Columns are visited ascending, so each CSR row ends
  // up sorted by column -- the same final layout the reference reaches with its in-place
  // COO->CSR plus per-row bubble sort (formatConvert.cpp:165-215).
  for (int p = 0; p < nnz; ++p) rp[M.get_index(p) + 1]++;
  for (int r = 0; r < rows; ++r) rp[r + 1] += rp[r];
  std::vector<int> cur(rp.begin(), rp.end() - 1);
  for (int c = 0; c < cols; ++c)
    for (int p = M.get_pointer(c); p < M.get_pointer(c + 1); ++p) {
      int q = cur[M.get_index(p)]++;
      ci[q] = c;
      v[q] = M.get_data(p);
    }
}

extern "C" {

// A given as CSC (double values, int indices), exactly what the driver hands ILU++.
void* ilupp_create(int n, const int* Ap, const int* Ai, const double* Ax, double threshold,
                   double mem_factor) {
  IluppObj* o = new IluppObj;
  o->n = n;
  int nnz = Ap[n];
  Real* d = new Real[nnz];
  int* idx = new int[nnz];
  int* ptr = new int[n + 1];
  std::memcpy(d, Ax, sizeof(double) * nnz);
  std::memcpy(idx, Ai, sizeof(int) * nnz);
  std::memcpy(ptr, Ap, sizeof(int) * (n + 1));
  o->A = new Matrix(d, idx, ptr, n, n, iluplusplus::COLUMN);

  iluplusplus::preprocessing_sequence L;
  L.set_MAX_WEIGHTED_MATCHING_ORDERING_DD_MOV_COR_IM();
  iluplusplus::iluplusplus_precond_parameter param;
  param.init(L, 11, "test");
  param.set_threshold(threshold);
  param.set_MEM_FACTOR(mem_factor);
  param.set_MAX_LEVELS(1);
  o->P.make_preprocessed_multilevelILUCDP(*o->A, param);
  o->fill = ((Real)o->P.total_nnz()) / (Real)o->A->non_zeroes();

  Matrix Lm = o->P.extract_left_matrix(0), Um = o->P.extract_right_matrix(0);
  Vector mid = o->P.extract_middle_matrix(0), ls = o->P.extract_left_scaling(0),
         rs = o->P.extract_right_scaling(0);
  iluplusplus::index_list pr = o->P.extract_permutation_rows(0),
                          pc = o->P.extract_inverse_permutation_columns(0);
  to_csr(Lm, o->lp, o->li, o->lv);
  to_csr(Um, o->up, o->ui, o->uv);
  o->middle.resize(n); o->lscale.resize(n); o->rscale.resize(n);
  o->permRow.resize(n); o->permCol.resize(n);
  for (int i = 0; i < n; ++i) {
    o->middle[i] = mid.get(i); o->lscale[i] = ls.get(i); o->rscale[i] = rs.get(i);
    o->permRow[i] = pr.get(i); o->permCol[i] = pc.get(i);
  }
  return o;
}

void ilupp_sizes(void* h, int* nnzL, int* nnzU, double* fill) {
  IluppObj* o = (IluppObj*)h;
  *nnzL = (int)o->lv.size(); *nnzU = (int)o->uv.size(); *fill = o->fill;
}

void ilupp_get(void* h, int* lp, int* li, double* lv, int* up, int* ui, double* uv, double* middle,
               int* permRow, int* permCol, double* lscale, double* rscale) {
  IluppObj* o = (IluppObj*)h;
  std::copy(o->lp.begin(), o->lp.end(), lp); std::copy(o->li.begin(), o->li.end(), li);
  std::copy(o->lv.begin(), o->lv.end(), lv); std::copy(o->up.begin(), o->up.end(), up);
  std::copy(o->ui.begin(), o->ui.end(), ui); std::copy(o->uv.begin(), o->uv.end(), uv);
  std::copy(o->middle.begin(), o->middle.end(), middle);
  std::copy(o->permRow.begin(), o->permRow.end(), permRow);
  std::copy(o->permCol.begin(), o->permCol.end(), permCol);
  std::copy(o->lscale.begin(), o->lscale.end(), lscale);
  std::copy(o->rscale.begin(), o->rscale.end(), rscale);
}

// iluplusplus::gmres SPLIT, as mna_solve_cpu_ilu_gmres (mna_solve_gpu_gmres.cpp:2069).
// rel_tol/abs_tol are -log10 values in ILU++; returns iterations, writes achieved exponents.
int ilupp_gmres(void* h, const double* b, double* x, int restart, int max_iter, double* rel_tol,
                double* abs_tol) {
  IluppObj* o = (IluppObj*)h;
  Vector bb(o->n), xx(o->n);
  for (int i = 0; i < o->n; ++i) bb[i] = b[i];
  xx.resize(o->n, 0.0);
  int min_iter = 0, it = max_iter;
  iluplusplus::gmres<Real, Matrix, Vector>(o->P, iluplusplus::SPLIT, *o->A, bb, xx, restart, min_iter,
                                           it, *rel_tol, *abs_tol, true);
  for (int i = 0; i < o->n; ++i) x[i] = xx[i];
  return it;
}

void ilupp_destroy(void* h) {
  IluppObj* o = (IluppObj*)h;
  delete o->A;
  delete o;
}

}  // extern "C"
