// gmres.h (B200 shim) -- the entry points of the reference's src/gmres.h (and src_thermal/gmres.h) with unchanged
// signatures, implemented over the C ABI (include/b200gmres.h).
//  * device entry points (gmres.h:236-282, 360-398): GMRES_GPU (two overloads), GMRES_GPU_tran, GMRESilu_GPU,
//    GMRES_GPU_leftDiag, GMRES_GPU_leftILU0, Update_GPU, sgemv_GPU  -> bg_gmres_solve / bg_spmv_axpby / bg_basis_update
//  * host twins (gmres.h:178-226, 320-346): GMRES (two overloads), GMRES_tran, GMRESilu, GMRES_leftDiag,
//    GMRES_leftILU0, GMRES_right_diag take HOST arrays like the reference's CPU path; this build has no CPU compute
//    path, so they upload, run the same algorithm on the device (bg_gmres_solve_host) and copy the answer back
//  * getLevel, ApplyPlaneRotation, GeneratePlaneRotation, difftime: host scalar/integer helpers, exact semantics
#ifndef _GMRES_H_
#define _GMRES_H_
#include <assert.h>
#include <math.h>
#include <sys/time.h>
#include <iostream>
#include <vector>
#include <cuda_runtime_api.h>
#include <vector_types.h>  // dim3 (kept in the signatures; ignored: kernels size their own grids)
#include "SpMV.h"
#include "config.h"
#include "SpMV_kernel.h"
#include "defs.h"
#include "preconditioner.h"
#define REAL float
using namespace std;
#define timeSentence(a, b) a

// workspace handle for the transient loop (gmres.h:82-112); owns a bg_solver_t
class GMRES_GPU_Data {
 public:
  int numRows;
  float *s, *cs, *sn, *H;
  float *d_r, *d_rr, *d_bb, *d_temp, *d_v, *d_w, *d_ww;
  void* bg_handle;
  GMRES_GPU_Data();
  void Initilize(const int m, const int n);
  ~GMRES_GPU_Data();
};

float difftime(timeval& st, timeval& et);
void ApplyPlaneRotation(float* dx, float* dy, float cs, float sn);
void GeneratePlaneRotation(float dx, float dy, float* cs, float* sn);
void getLevel(bool isLmatrix, const float* val, const int* rowIndices, const int* indices, int* level, int n);

// ---- device entry points
int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
              const int m, int* max_iter, float* tol, Preconditioner& preconditioner);
int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
              const int m, int* max_iter, float* tol);
int GMRES_GPU_tran(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
                   const int m, const int max_iter, const float tol, Preconditioner& preconditioner,
                   GMRES_GPU_Data& gmres_gpu_data);
int GMRESilu_GPU(float* d_val, int* d_rowIndices, int* d_indices, int nnz, float* d_x, float* d_b, const int n, const int m,
                 int* max_iter, float* tol, Preconditioner& preconditioner);
// m_* : HOST CSR of an explicit (diagonal) inverse applied by one SpMV (src/gmres.cu:1186-1231)
int GMRES_GPU_leftDiag(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
                       const int m, int* max_iter, float* tol, const float* m_val, const int* m_rowIndices,
                       const int* m_indices);
// l_* / u_* : HOST CSR factors in leftILU's layout (src/gmres.cu:1440-1700)
int GMRES_GPU_leftILU0(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
                       const int m, int* max_iter, float* tol, const float* l_val, const int* l_rowIndices,
                       const int* l_indices, const float* u_val, const int* u_rowIndices, const int* u_indices);
// x += V[:,0..k] y after the host back-substitution of the (k+1) x (k+1) triangle of H (src/gmres.cu:164-189)
void Update_GPU(float* d_x, const int k, const float* H, const int m, const float* s, const float* d_v, const int n);
// v = alpha*A*x + beta*y (gmres.h:229-233)
void sgemv_GPU(float* v, const SpMatrixGPU* Sparse, const SpMatrix* spm, dim3* grid, dim3* block, const float alpha,
               const float* x, const float beta, const float* y, const int numRows, const int numCols);

// ---- host twins (host arrays in and out; computed on the device)
int GMRES(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n, const int m,
          int* max_iter, float* tol, Preconditioner& preconditioner);
int GMRES(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n, const int m,
          int* max_iter, float* tol);
int GMRES_tran(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n, const int m,
               const int max_iter, const float tol, Preconditioner& preconditioner);
int GMRESilu(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n, const int m,
             int* max_iter, float* tol, Preconditioner& preconditioner);
int GMRES_leftDiag(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n,
                   const int m, int* max_iter, float* tol, const float* m_val, const int* m_rowIndices, const int* m_indices);
int GMRES_leftILU0(const float* val, const int* rowIndices, const int* indices, float* x, const float* b, const int n,
                   const int m, int* max_iter, float* tol, const float* l_val, const int* l_rowIndices, const int* l_indices,
                   const float* u_val, const int* u_rowIndices, const int* u_indices);
#endif
