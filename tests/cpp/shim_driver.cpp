// shim_driver.cpp -- a miniature of the reference's own drivers written against the reference's
// headers as re-created by the shim (gpu-gmres_b200/host/shim):
//   thermal  : src_thermal/main1.cu:126-405  (SpMatrix triplets -> genCSRFormat -> allocateSparseMatrixGPU
//              -> MyILU0/MyDIAG/MyNONE::Initilize -> GMRES_GPU)
//   tran     : src_thermal/main2.cu:437-515  (GMRES_GPU_Data workspace, GMRES_GPU_tran per time step, warm start)
//   pg       : src/mna_solve_gpu_gmres.cpp:507-531, 608-621 (gmresInterfacePGfloat::setPrecondPG with ILU++
//              factors, rhs_h / xgmres_h public members, GMRES_dev_PG for DC and for a warm-started step)
// Test infrastructure: built and run by tests/test_shim_gpu.py; results are compared with the oracle there.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200gmres.h"
#include "gmres.h"
#include "gmres_interface_pg.h"
#include "leftILU.h"
#include "preconditioner.h"

static std::vector<char> slurp(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> b(sz);
  if (std::fread(b.data(), 1, sz, f) != (size_t)sz) std::exit(2);
  std::fclose(f);
  return b;
}
struct Reader {
  const char* p;
  template <class T> const T* take(size_t n) { const T* r = (const T*)p; p += sizeof(T) * n; return r; }
};
static bg_context_t g_ctx;
static float* to_dev(const float* h, int n) {
  float* d = nullptr;
  if (bg_device_malloc(g_ctx, sizeof(float) * n, (void**)&d) < 0 || bg_memcpy_h2d(g_ctx, d, h, sizeof(float) * n) < 0) std::exit(3);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: shim_driver <in> <out> thermal-ilu0|thermal-diag|thermal-none|tran|pg\n"); return 2; }
  const std::string mode = argv[3];
  std::vector<char> buf = slurp(argv[1]);
  Reader R{buf.data()};
  const int n = *R.take<int>(1), nnz = *R.take<int>(1);
  const int* rp = R.take<int>(n + 1);
  const int* ci = R.take<int>(nnz);
  const float* val = R.take<float>(nnz);
  const float* b = R.take<float>(n);
  if (bg_context_create(&g_ctx, 0) < 0) { std::fprintf(stderr, "%s\n", bg_last_error()); return 3; }
  FILE* out = std::fopen(argv[2], "wb");
  auto emit = [&](int flag, int iters, float tol, const float* x) {
    std::fwrite(&flag, 4, 1, out); std::fwrite(&iters, 4, 1, out); std::fwrite(&tol, 4, 1, out); std::fwrite(x, 4, n, out);
  };
  std::vector<float> x(n, 0.f);

  if (mode == "pg") {
    const int nnzL = *R.take<int>(1), nnzU = *R.take<int>(1);
    MySpMatrix A; MySpMatrixDouble L, U; MySpMatrix D, Pr, Pc, ls, rs;
    std::memset(&A, 0, sizeof(A)); std::memset(&L, 0, sizeof(L)); std::memset(&U, 0, sizeof(U));
    std::memset(&D, 0, sizeof(D)); std::memset(&Pr, 0, sizeof(Pr)); std::memset(&Pc, 0, sizeof(Pc)); std::memset(&ls, 0, sizeof(ls)); std::memset(&rs, 0, sizeof(rs));
    A.numRows = A.numCols = n; A.numNZEntries = nnz; A.val = (float*)val; A.rowIndices = (int*)rp; A.indices = (int*)ci;
    L.numRows = U.numRows = n;
    L.rowIndices = (int*)R.take<int>(n + 1); L.indices = (int*)R.take<int>(nnzL); L.val = (double*)R.take<double>(nnzL);
    U.rowIndices = (int*)R.take<int>(n + 1); U.indices = (int*)R.take<int>(nnzU); U.val = (double*)R.take<double>(nnzU);
    D.val = (float*)R.take<float>(n); ls.val = (float*)R.take<float>(n); rs.val = (float*)R.take<float>(n);
    Pr.indices = (int*)R.take<int>(n); Pc.indices = (int*)R.take<int>(n);
    gmresInterfacePGfloat G;
    G.setPrecondPG(&A, &L, &U, &D, &Pr, &Pc, &ls, &rs);
    // DC solve from x = 0 (mna_solve_gpu_gmres.cpp:525-531)
    for (int i = 0; i < n; ++i) { G.xgmres_h[i] = 0.f; G.rhs_h[i] = b[i]; }
    int flag = G.GMRES_dev_PG();
    emit(flag, G.max_it, G.tol, G.xgmres_h);
    // one "time step": new right-hand side, xgmres_h keeps the previous solution (warm start, :615-619)
    for (int i = 0; i < n; ++i) G.rhs_h[i] = b[i] * 1.01f;
    flag = G.GMRES_dev_PG();
    emit(flag, G.max_it, G.tol, G.xgmres_h);
    std::fclose(out);
    return 0;
  }

  // thermal drivers: triplets (row-major sorted) + rowPtrs, as readSparseMatrix leaves them
  SpMatrix spm;
  spm.numRows = spm.numCols = n; spm.numNZEntries = nnz; spm.colPtrs = nullptr;
  std::vector<NZEntry> nz(nnz);
  std::vector<int> rowPtrs(rp, rp + n);
  for (int i = 0; i < n; ++i)
    for (int p = rp[i]; p < rp[i + 1]; ++p) nz[p] = NZEntry{i, ci[p], val[p]};
  spm.nzentries = nz.data(); spm.rowPtrs = rowPtrs.data();
  std::vector<float> h_val(nnz);
  std::vector<int> h_rowIndices(n + 1), h_indices(nnz);
  genCSRFormat(&spm, h_val.data(), h_rowIndices.data(), h_indices.data());
  SpMatrixGPU Sparse;
  allocateSparseMatrixGPU(&Sparse, &spm, h_val.data(), h_rowIndices.data(), h_indices.data(), n, n);
  dim3 grid(1, (n + 31) / 32), block(1, 512);  // the reference's launch shape (main1.cu:207-214); ignored by the shim
  MySpMatrix mySpM;
  mySpM.Initilize(spm);
  Preconditioner* precond = nullptr;
  if (mode == "thermal-none") precond = new MyNONE();
  else if (mode == "thermal-diag") precond = new MyDIAG();
  else precond = new MyILU0();
  precond->Initilize(mySpM);
  float* d_x = to_dev(x.data(), n);
  float* d_b = to_dev(b, n);

  if (mode == "tran") {
    GMRES_GPU_Data data;
    data.Initilize(restart, n);
    std::vector<float> bk(n);
    for (int k = 0; k < 3; ++k) {  // main2.cu:489-506: new rhs every step, d_x carried over
      for (int i = 0; i < n; ++i) bk[i] = b[i] * (1.f + 0.05f * k);
      bg_memcpy_h2d(g_ctx, d_b, bk.data(), sizeof(float) * n);
      const int flag = GMRES_GPU_tran(&Sparse, &spm, &grid, &block, d_x, d_b, n, restart, 2000, 1e-6f, *precond, data);
      bg_memcpy_d2h(g_ctx, x.data(), d_x, sizeof(float) * n);
      emit(flag, k, 0.f, x.data());
    }
  } else {
    int max_it = 2000;
    float tol = 1e-6f;  // src/defs.h:12
    const int flag = GMRES_GPU(&Sparse, &spm, &grid, &block, d_x, d_b, n, restart, &max_it, &tol, *precond);
    bg_memcpy_d2h(g_ctx, x.data(), d_x, sizeof(float) * n);
    emit(flag, max_it, tol, x.data());
    // the preconditioner on its own, as foreign GMRES loops call it
    float* d_z = to_dev(x.data(), n);
    precond->DevPrecond(d_b, d_z);
    bg_memcpy_d2h(g_ctx, x.data(), d_z, sizeof(float) * n);
    emit(0, 0, 0.f, x.data());
    bg_device_free(g_ctx, d_z);
  }
  std::fclose(out);
  bg_device_free(g_ctx, d_x);
  bg_device_free(g_ctx, d_b);
  delete precond;
  freeSparseMatrixGPU(&Sparse);
  mySpMatrixFree(&mySpM);
  return 0;
}
