/* examples/solve_csr.c -- the C ABI from plain C: GMRES(32) + ILU(0) on a small 2-D Poisson matrix, host vectors in
   and out (the call gmresInterfacePGfloat::GMRES_dev_PG makes, src/gmres_interface_pg.cu:110-139).

     gcc -std=c99 -Iinclude examples/solve_csr.c -Lgpu-gmres_b200 -lb200gmres -Wl,-rpath,$PWD/gpu-gmres_b200 -lm -o solve_csr
     ./solve_csr [nx]

   Without a GPU every compute entry point fails with BG_ERR_NO_GPU (there is no CPU fallback): the program prints
   the library's message and exits with status 3. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200gmres.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ < 0) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, bg_last_error()); \
      return rc_ == BG_ERR_NO_GPU ? 3 : 1;                                 \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 200, n = nx * nx;
  int* rp = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  int* ci = (int*)malloc(sizeof(int) * (size_t)n * 5);
  double* v = (double*)malloc(sizeof(double) * (size_t)n * 5);
  double* b = (double*)malloc(sizeof(double) * (size_t)n);
  double* x = (double*)calloc((size_t)n, sizeof(double));
  int nnz = 0;
  rp[0] = 0;
  for (int y = 0; y < nx; ++y)
    for (int xx = 0; xx < nx; ++xx) {
      const int i = y * nx + xx;
      if (y > 0) { ci[nnz] = i - nx; v[nnz++] = -1.0; }
      if (xx > 0) { ci[nnz] = i - 1; v[nnz++] = -1.0; }
      ci[nnz] = i; v[nnz++] = 4.05;
      if (xx < nx - 1) { ci[nnz] = i + 1; v[nnz++] = -1.0; }
      if (y < nx - 1) { ci[nnz] = i + nx; v[nnz++] = -1.0; }
      rp[i + 1] = nnz;
      b[i] = 1.0 + 0.001 * (double)(i % 17);
    }

  bg_context_t ctx;
  bg_matrix_t A;
  bg_precond_t M;
  bg_solver_t S;
  CHECK(bg_context_create(&ctx, 0));
  CHECK(bg_matrix_create_csr(ctx, n, nnz, rp, ci, v, BG_F64, BG_HOST, BG_F64, BG_SPMV_AUTO, &A));
  CHECK(bg_precond_create_ilu0(ctx, n, rp, ci, v, BG_F64, BG_F64, NULL, &M)); /* reference-order ILU(0), all-SM triangular solves */
  CHECK(bg_solver_create(ctx, n, 32, BG_F64, &S));
  bg_gmres_opts o;
  bg_gmres_opts_default(&o); /* restart 32, strict MGS: the reference's choices */
  o.tol = 1e-8;
  o.max_iter = 5000;
  bg_gmres_result r;
  CHECK(bg_gmres_solve_host(ctx, S, A, M, b, x, &o, &r));

  double res = 0.0, nb = 0.0; /* true residual on the host */
  for (int i = 0; i < n; ++i) {
    double a = b[i];
    for (int p = rp[i]; p < rp[i + 1]; ++p) a -= v[p] * x[ci[p]];
    res += a * a;
    nb += b[i] * b[i];
  }
  printf("n = %d: flag %d, %d iterations, relative residual %.3e (true %.3e), %.3f ms on the device\n", n, r.flag, r.iters, r.resid,
         sqrt(res / nb), r.solve_ms);
  bg_solver_destroy(S);
  bg_precond_destroy(M);
  bg_matrix_destroy(A);
  bg_context_destroy(ctx);
  free(rp); free(ci); free(v); free(b); free(x);
  return r.flag == 0 ? 0 : 2;
}
