// gmres_interface_pg.h (B200 shim) -- the power-grid facade (src/gmres_interface_pg.h:5-73) with the
// same public data members the driver reads and writes (mna_solve_gpu_gmres.cpp:507-544, 608-621, 1180, 1252).
#ifndef _GMRES_INTERFACE_PG_H_
#define _GMRES_INTERFACE_PG_H_
#include "SpMV.h"

// FP64-scaling variant used by the `-gmres` CPU driver (mna_solve_cpu_gmres): same solve, run on the device
class gmresInterfacePG {
 public:
  gmresInterfacePG();
  ~gmresInterfacePG();
  int matrixSize;
  float* h_val; int* h_rowPtr; int* h_colIdx;
  float *x_h, *x_d;
  float *xgmres_h, *rhs_h;
  void* Precond;
  int max_it;  // both input and output
  float tol;
  void setPrecondPG(MySpMatrix* A, MySpMatrixDouble* PrLeft, MySpMatrixDouble* PrRight, MySpMatrix* PrMiddle_mySpM,
                    MySpMatrix* PrPermRow, MySpMatrix* PrPermCol, MySpMatrixDouble* PrLscale, MySpMatrixDouble* PrRscale);
  int GMRES_host_PG();  // host vectors in, host vectors out; the solve itself runs on the device
 private:
  void *bg_matrix, *bg_solver;
};

class gmresInterfacePGfloat {
 public:
  gmresInterfacePGfloat();
  ~gmresInterfacePGfloat();
  int matrixSize;
  int nnz;
  float* h_val; int* h_rowPtr; int* h_colIdx;
  float* d_val; int* d_rowPtr; int* d_colIdx;
  float *x_h, *x_d;
  float *xgmres_h, *rhs_h;
  float *xgmres_d, *rhs_d;
  void* Precond;
  int max_it;  // both input and output
  float tol;
  void setPrecondPG(MySpMatrix* A, MySpMatrixDouble* PrLeft, MySpMatrixDouble* PrRight, MySpMatrix* PrMiddle_mySpM,
                    MySpMatrix* PrPermRow, MySpMatrix* PrPermCol, MySpMatrix* PrLscale, MySpMatrix* PrRscale);
  int GMRES_host_PG();  // same call as GMRES_dev_PG here: host vectors, device solve (no CPU compute path)
  int GMRES_dev_PG();
 private:
  void *bg_matrix, *bg_solver;
};
#endif
