// gmres.h (B200 shim) -- device-side entry points of the reference's src/gmres.h with unchanged
// signatures (gmres.h:360-398, 82-112, 164-167, 229-233), implemented over bg_gmres_solve.
#ifndef _GMRES_H_
#define _GMRES_H_
#include <vector_types.h>  // dim3 (kept in the signatures; ignored: kernels size their own grids)
#include "SpMV.h"
#include "defs.h"
#include "preconditioner.h"
#define REAL float

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

int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
              const int m, int* max_iter, float* tol, Preconditioner& preconditioner);
int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
              const int m, int* max_iter, float* tol);
int GMRES_GPU_tran(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n,
                   const int m, const int max_iter, const float tol, Preconditioner& preconditioner,
                   GMRES_GPU_Data& gmres_gpu_data);
int GMRESilu_GPU(float* d_val, int* d_rowIndices, int* d_indices, int nnz, float* d_x, float* d_b, const int n, const int m,
                 int* max_iter, float* tol, Preconditioner& preconditioner);
// v = alpha*A*x + beta*y (gmres.h:229-233)
void sgemv_GPU(float* v, const SpMatrixGPU* Sparse, const SpMatrix* spm, dim3* grid, dim3* block, const float alpha,
               const float* x, const float beta, const float* y, const int numRows, const int numCols);
void getLevel(bool isLmatrix, const float* val, const int* rowIndices, const int* indices, int* level, int n);
#endif
