// shim.cpp -- the reference's solver API (gmres.h, preconditioner.h, gmres_interface_pg.h, SpMV.h,
// leftILU.h of sheldonucr/GPU-GMRES) implemented on top of the C ABI in include/b200gmres.h.
//
// Drop-in layer for the unchanged drivers: src/mna_solve_gpu_gmres.cpp (power grid) and the
// thermal mains src_thermal/main1.cu / main2.cu.  Same class names, public data members, argument
// meaning and return conventions (0 = converged, 1 = not converged; *max_iter / *tol are in-out,
// SURVEY Q2/Q3; hard failures print and abort like the reference's exit(-1)).
// Pure host C++: device memory is reached through bg_device_malloc / bg_memcpy_* so this file
// needs no CUDA runtime.  There is no CPU compute path here: the Host* twins abort.
//
// The vectors are float (the reference's REAL, gmres.h:55); environment switches select the
// device schedule without touching driver source:
//   B200GMRES_DEVICE=<id>      device ordinal (default 0; the reference hard-codes 1, setGPUdevice.cu:24)
//   B200GMRES_TRSV=block       block-Jacobi ILU(0) solves (performance mode) instead of the
//                              reference-exact global level schedule, for MyILU0 / MyILUK
//   B200GMRES_ORTH=mgs|cgs     strict modified Gram-Schmidt (reference order) or fused CGS (default)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "../../../include/b200gmres.h"
#include "gmres.h"
#include "gmres_interface_pg.h"
#include "leftILU.h"
#include "preconditioner.h"

namespace {

[[noreturn]] void die(const char* what) {
  std::fprintf(stderr, "b200gmres shim: %s failed: %s\n", what, bg_last_error());
  std::exit(-1);
}
#define CHK(call)            \
  do {                       \
    if ((call) < 0) die(#call); \
  } while (0)

bg_context_t g_ctx = nullptr;
bg_context_t ctx() {
  if (!g_ctx) {
    const char* e = std::getenv("B200GMRES_DEVICE");
    CHK(bg_context_create(&g_ctx, e ? std::atoi(e) : 0));
  }
  return g_ctx;
}

// device CSR handed over as raw pointers (SpMatrixGPU / MySpMatrix / GMRESilu_GPU arguments):
// one bg_matrix_t per distinct value array, built on first use
struct MatKey {
  const void* val; int n; long long nnz;
  bool operator<(const MatKey& o) const { return val != o.val ? val < o.val : n != o.n ? n < o.n : nnz < o.nnz; }
};
std::map<MatKey, bg_matrix_t> g_mats;
bg_matrix_t matrix_for(const float* d_val, const int* d_rowIndices, const int* d_indices, int n, long long nnz) {
  MatKey k{d_val, n, nnz};
  auto it = g_mats.find(k);
  if (it != g_mats.end()) return it->second;
  bg_matrix_t A = nullptr;
  CHK(bg_matrix_create_csr(ctx(), n, nnz, d_rowIndices, d_indices, d_val, BG_F32, BG_DEVICE, BG_F32, BG_SPMV_AUTO, &A));
  g_mats[k] = A;
  return A;
}
void forget_matrix(const void* d_val) {
  for (auto it = g_mats.begin(); it != g_mats.end();)
    if (it->first.val == d_val) { bg_matrix_destroy(it->second); it = g_mats.erase(it); } else ++it;
}

std::map<std::pair<int, int>, bg_solver_t> g_solvers;  // (n, m) -> workspace, FP32
bg_solver_t solver_for(int n, int m) {
  auto key = std::make_pair(n, m);
  auto it = g_solvers.find(key);
  if (it != g_solvers.end()) return it->second;
  bg_solver_t S = nullptr;
  CHK(bg_solver_create(ctx(), n, m, BG_F32, &S));
  g_solvers[key] = S;
  return S;
}

int orth_mode() {
  const char* e = std::getenv("B200GMRES_ORTH");
  return e && std::strcmp(e, "mgs") == 0 ? BG_ORTH_MGS : BG_ORTH_CGS_ADAPTIVE;
}

// common tail of every GMRES*_GPU entry point: in-out max_iter / tol, 0/1 result
int run_gmres(bg_solver_t S, bg_matrix_t A, Preconditioner& M, const float* d_b, float* d_x, int m, int* max_iter,
              float* tol) {
  bg_gmres_opts o;
  bg_gmres_opts_default(&o);
  o.m = m; o.max_iter = *max_iter; o.tol = (double)*tol; o.orth = orth_mode();
  bg_gmres_result r;
  const int rc = bg_gmres_solve(ctx(), S, A, (bg_precond_t)M.bg_handle, d_b, d_x, &o, &r);
  if (rc < 0) die("bg_gmres_solve");
  *max_iter = r.iters;
  *tol = (float)r.resid;
  return rc;  // BG_OK (0) or BG_NOT_CONVERGED (1)
}

MyNONE* none_precond(int n) {
  static std::map<int, MyNONE*> cache;
  auto it = cache.find(n);
  if (it != cache.end()) return it->second;
  MyNONE* p = new MyNONE;
  MySpMatrix dummy;
  std::memset(&dummy, 0, sizeof(dummy));
  dummy.numRows = dummy.numCols = n;
  p->Initilize(dummy);
  cache[n] = p;
  return p;
}

}  // namespace

// ------------------------------------------------------------------------------------------ SpMV.h
void setGPUdevice() { (void)ctx(); }

void MySpMatrix::Initilize(SpMatrix& M) {  // src/mySpMatrix.cu:11-39
  isCSR = 1;
  numRows = M.numRows; numCols = M.numCols; numNZEntries = M.numNZEntries;
  val = new float[M.numNZEntries];
  indices = new int[M.numNZEntries];
  rowIndices = new int[M.numRows + 1];
  for (int i = 0; i < numNZEntries; ++i) { val[i] = M.nzentries[i].val; indices[i] = M.nzentries[i].colNum; }
  for (int i = 0; i < numRows; ++i) rowIndices[i] = M.rowPtrs[i];
  rowIndices[numRows] = M.numNZEntries;
  CHK(bg_device_malloc(ctx(), sizeof(float) * (size_t)numNZEntries, (void**)&d_val));
  CHK(bg_device_malloc(ctx(), sizeof(int) * (size_t)numNZEntries, (void**)&d_indices));
  CHK(bg_device_malloc(ctx(), sizeof(int) * (size_t)(numRows + 1), (void**)&d_rowIndices));
  CHK(bg_memcpy_h2d(ctx(), d_val, val, sizeof(float) * (size_t)numNZEntries));
  CHK(bg_memcpy_h2d(ctx(), d_indices, indices, sizeof(int) * (size_t)numNZEntries));
  CHK(bg_memcpy_h2d(ctx(), d_rowIndices, rowIndices, sizeof(int) * (size_t)(numRows + 1)));
}
void mySpMatrixFree(MySpMatrix* m) {  // src/SpMV_alloc.cu:181-195
  forget_matrix(m->d_val);
  delete[] m->val; delete[] m->indices; delete[] m->rowIndices;
  if (m->d_val) { bg_device_free(ctx(), m->d_val); bg_device_free(ctx(), m->d_indices); bg_device_free(ctx(), m->d_rowIndices); }
  m->val = nullptr; m->indices = m->rowIndices = nullptr; m->d_val = nullptr; m->d_indices = m->d_rowIndices = nullptr;
}
void mySpMatrixDoubleFree(MySpMatrixDouble* m) {  // src/SpMV_alloc.cu:197-211
  delete[] m->val; delete[] m->indices; delete[] m->rowIndices;
  m->val = nullptr; m->indices = m->rowIndices = nullptr;
}

void genCSRFormat(SpMatrix* m, float* val, int* rowIndices, int* indices) {  // src_thermal/SpMV_gen.cpp:189-201
  for (int i = 0; i < m->numNZEntries; ++i) { val[i] = m->nzentries[i].val; indices[i] = m->nzentries[i].colNum; }
  for (int i = 0; i < m->numRows; ++i) rowIndices[i] = m->rowPtrs[i];
  rowIndices[m->numRows] = m->numNZEntries;
}
void gpuMallocCpyCSR(SpMatrixGPU* spm, float* h_val, int* h_rowIndices, int* h_indices, const int numRows, const int) {
  const size_t nnz = (size_t)h_rowIndices[numRows];  // src/SpMV_alloc.cu:121-143
  std::memset(spm, 0, sizeof(*spm));
  CHK(bg_device_malloc(ctx(), sizeof(float) * nnz, (void**)&spm->d_val));
  CHK(bg_device_malloc(ctx(), sizeof(int) * nnz, (void**)&spm->d_indices));
  CHK(bg_device_malloc(ctx(), sizeof(int) * (size_t)(numRows + 1), (void**)&spm->d_rowIndices));
  CHK(bg_memcpy_h2d(ctx(), spm->d_val, h_val, sizeof(float) * nnz));
  CHK(bg_memcpy_h2d(ctx(), spm->d_indices, h_indices, sizeof(int) * nnz));
  CHK(bg_memcpy_h2d(ctx(), spm->d_rowIndices, h_rowIndices, sizeof(int) * (size_t)(numRows + 1)));
}
void allocateSparseMatrixGPU(SpMatrixGPU* spm, SpMatrix*, float* h_val, int* h_rowIndices, int* h_indices, const int numRows,
                             const int numCols) {  // src/SpMV_alloc.cu:16-118 (inspector arrays are not needed here)
  gpuMallocCpyCSR(spm, h_val, h_rowIndices, h_indices, numRows, numCols);
}
void gpuMallocCpyCSRmySpM(SpMatrixGPU* spm, MySpMatrix* m) {  // src/SpMV_alloc.cu:145-162
  gpuMallocCpyCSR(spm, m->val, m->rowIndices, m->indices, m->numRows, m->numCols);
}
void freeSparseMatrixGPU(SpMatrixGPU* spm) {  // src/SpMV_alloc.cu:165-179
  forget_matrix(spm->d_val);
  if (spm->d_val) { bg_device_free(ctx(), spm->d_val); bg_device_free(ctx(), spm->d_indices); bg_device_free(ctx(), spm->d_rowIndices); }
  std::memset(spm, 0, sizeof(*spm));
}
void SpMV_cuda(float* d_x, SpMatrixGPU* spm, const float* d_y, const int numRows, const int, const int nnz) {
  CHK(bg_spmv(ctx(), matrix_for(spm->d_val, spm->d_rowIndices, spm->d_indices, numRows, nnz), d_y, d_x));
  CHK(bg_context_synchronize(ctx()));
}

// ------------------------------------------------------------------------------------------ preconditioner.h
Preconditioner::Preconditioner()
    : numRows(0), d_r(nullptr), d_rr(nullptr), d_bb(nullptr), d_y(nullptr), s(nullptr), cs(nullptr), sn(nullptr), H(nullptr),
      d_v(nullptr), d_w(nullptr), d_ww(nullptr), bg_handle(nullptr) {}
Preconditioner::~Preconditioner() {
  if (bg_handle) bg_precond_destroy((bg_precond_t)bg_handle);
}
static void no_host_path(const char* what) {
  std::fprintf(stderr, "b200gmres shim: %s is the reference's CPU twin; this build has no CPU compute path\n", what);
  std::exit(-1);
}
void Preconditioner::HostPrecond(const ValueType*, ValueType*) { no_host_path("HostPrecond"); }
void Preconditioner::HostPrecond_rhs(const ValueType*, ValueType*) { no_host_path("HostPrecond_rhs"); }
void Preconditioner::HostPrecond_right(const ValueType*, ValueType*) { no_host_path("HostPrecond_right"); }
void Preconditioner::HostPrecond_left(const ValueType*, ValueType*) { no_host_path("HostPrecond_left"); }
void Preconditioner::HostPrecond_starting_value(const ValueType*, ValueType*) { no_host_path("HostPrecond_starting_value"); }
static void dev_apply(Preconditioner* p, int which, const float* in, float* out) {
  CHK(bg_precond_apply(ctx(), (bg_precond_t)p->bg_handle, which, in, out));
  CHK(bg_context_synchronize(ctx()));  // the reference's DevPrecond ends with cudaDeviceSynchronize (preconditioner.cu:283)
}
void Preconditioner::DevPrecond(const ValueType* i, ValueType* o) { dev_apply(this, BG_PC_FULL, i, o); }
void Preconditioner::DevPrecond_rhs(float* i, float* o) { dev_apply(this, BG_PC_RHS, i, o); }
void Preconditioner::DevPrecond_right(float* i, float* o) { dev_apply(this, BG_PC_RIGHT, i, o); }
void Preconditioner::DevPrecond_left(float* i, float* o) { dev_apply(this, BG_PC_LEFT, i, o); }
void Preconditioner::DevPrecond_starting_value(float* i, float* o) { dev_apply(this, BG_PC_START, i, o); }

void MyNONE::Initilize(const MySpMatrix& A) {
  numRows = A.numRows;
  CHK(bg_precond_create_none(ctx(), numRows, BG_F32, (bg_precond_t*)&bg_handle));
}
void MyDIAG::Initilize(const MySpMatrix& A) {  // src/preconditioner.cu:500-540
  numRows = A.numRows;
  CHK(bg_precond_create_diag(ctx(), numRows, A.rowIndices, A.indices, A.val, BG_F32, BG_F32, (bg_precond_t*)&bg_handle));
}
MyILU0::MyILU0() : l_val(nullptr), u_val(nullptr), l_rowIndices(nullptr), l_indices(nullptr), u_rowIndices(nullptr), u_indices(nullptr) {}
MyILU0::~MyILU0() {
  std::free(l_val); std::free(u_val); std::free(l_rowIndices); std::free(l_indices); std::free(u_rowIndices); std::free(u_indices);
}
void MyILU0::Initilize(const MySpMatrix& A) {  // src/preconditioner.cu:288-362
  numRows = A.numRows;
  // host factors as public members, like the reference (leftILU output, caller-visible)
  leftILU(numRows, A.val, A.rowIndices, A.indices, l_val, l_rowIndices, l_indices, u_val, u_rowIndices, u_indices);
  bg_precond_opts o;
  bg_precond_opts_default(&o);
  const char* e = std::getenv("B200GMRES_TRSV");
  if (e && std::strcmp(e, "block") == 0) {
    o.trsv_schedule = BG_TRSV_BLOCK;
    o.block_method = BG_PART_METIS;
    CHK(bg_precond_create_ilu0(ctx(), numRows, A.rowIndices, A.indices, A.val, BG_F32, BG_F32, &o, (bg_precond_t*)&bg_handle));
  } else {
    CHK(bg_precond_create_lu(ctx(), numRows, l_rowIndices, l_indices, l_val, u_rowIndices, u_indices, u_val, BG_F32, BG_F32, &o,
                             (bg_precond_t*)&bg_handle));
  }
}
void MyILUPPfloat::Initilize(const MySpMatrix&) {
  std::fprintf(stderr, "b200gmres shim: MyILUPPfloat needs the ILU++ factors (use the 8-argument Initilize)\n");
  std::exit(-1);
}
void MyILUPPfloat::Initilize(const MySpMatrixDouble& L, const MySpMatrixDouble& U, const MySpMatrix& D, const MySpMatrix& Pr,
                             const MySpMatrix& Pc, const MySpMatrix& ls, const MySpMatrix& rs, int) {
  // src/preconditioner.cu:1205-1334: L lower non-unit (diagonal last), U upper (diagonal first), FP64
  // values; middle / lscale / rscale are float vectors; permutations live in the `indices` arrays
  numRows = L.numRows;
  CHK(bg_precond_create_split(ctx(), numRows, L.rowIndices, L.indices, L.val, U.rowIndices, U.indices, U.val, BG_F64, D.val, ls.val,
                              rs.val, BG_F32, Pr.indices, Pc.indices, BG_F32, nullptr, (bg_precond_t*)&bg_handle));
}

// ------------------------------------------------------------------------------------------ leftILU.h
static int* dup_i(const std::vector<int>& v) {
  int* p = (int*)std::malloc(sizeof(int) * (v.size() ? v.size() : 1));
  std::memcpy(p, v.data(), sizeof(int) * v.size());
  return p;
}
void leftILU(const int numRows, float* h_val, int* h_rowIndices, int* h_indices, float*& l_val, int*& l_rowIndices, int*& l_indices,
             float*& u_val, int*& u_rowIndices, int*& u_indices) {  // src/leftILU.cu:27-336; six malloc'ed arrays, caller owns
  bg_factors_t F = nullptr;
  CHK(bg_host_ilu0(numRows, h_rowIndices, h_indices, h_val, BG_F32, nullptr, &F));
  bg_factors_info fi;
  CHK(bg_factors_get_info(F, &fi));
  std::vector<int> lp(numRows + 1), li(fi.nnzL), up(numRows + 1), ui(fi.nnzU);
  l_val = (float*)std::malloc(sizeof(float) * (fi.nnzL ? fi.nnzL : 1));
  u_val = (float*)std::malloc(sizeof(float) * (fi.nnzU ? fi.nnzU : 1));
  CHK(bg_factors_get(F, lp.data(), li.data(), l_val, up.data(), ui.data(), u_val));
  l_rowIndices = dup_i(lp); l_indices = dup_i(li); u_rowIndices = dup_i(up); u_indices = dup_i(ui);
  bg_factors_destroy(F);
}
void generateLevel(const int numRows, const float* h_val, const int* h_rowIndices, const int* h_indices, std::vector<int>& node,
                   std::vector<int>& level) {  // src/leftILU.cu:339-368; bucketing 41-49
  level.assign(numRows, 0);
  CHK(bg_host_factor_levels(numRows, h_rowIndices, h_indices, h_val, level.data()));
  int nl = 0;
  for (int v : level) nl = v + 1 > nl ? v + 1 : nl;
  std::vector<int> lp(nl + 1);
  node.assign(numRows, 0);
  CHK(bg_host_bucket_levels(numRows, level.data(), &nl, lp.data(), node.data()));
}

// ------------------------------------------------------------------------------------------ gmres.h
GMRES_GPU_Data::GMRES_GPU_Data()
    : numRows(0), s(nullptr), cs(nullptr), sn(nullptr), H(nullptr), d_r(nullptr), d_rr(nullptr), d_bb(nullptr), d_temp(nullptr),
      d_v(nullptr), d_w(nullptr), d_ww(nullptr), bg_handle(nullptr) {}
void GMRES_GPU_Data::Initilize(const int m, const int n) {  // gmres.h:89-105: workspace allocated once per transient run
  numRows = n;
  CHK(bg_solver_create(ctx(), n, m, BG_F32, (bg_solver_t*)&bg_handle));
}
GMRES_GPU_Data::~GMRES_GPU_Data() {
  if (bg_handle) bg_solver_destroy((bg_solver_t)bg_handle);
}

void getLevel(bool isLmatrix, const float* val, const int* rowIndices, const int* indices, int* level, int n) {
  CHK(bg_host_trsv_levels(isLmatrix ? 1 : 0, n, rowIndices, indices, val, BG_F32, level));  // src/gmres.cu:1033-1077
}

void sgemv_GPU(float* v, const SpMatrixGPU* Sparse, const SpMatrix* spm, dim3*, dim3*, const float alpha, const float* x,
               const float beta, const float* y, const int numRows, const int) {  // src/gmres.cu:1010-1030
  bg_matrix_t A = matrix_for(Sparse->d_val, Sparse->d_rowIndices, Sparse->d_indices, numRows, spm->numNZEntries);
  CHK(bg_spmv_axpby(ctx(), A, alpha, x, beta, y, v));
  CHK(bg_context_synchronize(ctx()));
}

int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3*, dim3*, float* d_x, const float* d_b, const int n, const int m, int* max_iter,
              float* tol, Preconditioner& M) {  // src/gmres.cu:2567-2732
  bg_matrix_t A = matrix_for(Sparse->d_val, Sparse->d_rowIndices, Sparse->d_indices, n, spm->numNZEntries);
  return run_gmres(solver_for(n, m), A, M, d_b, d_x, m, max_iter, tol);
}
int GMRES_GPU(SpMatrixGPU* Sparse, SpMatrix* spm, dim3* grid, dim3* block, float* d_x, const float* d_b, const int n, const int m,
              int* max_iter, float* tol) {  // gmres.h:236-240: no preconditioner
  return GMRES_GPU(Sparse, spm, grid, block, d_x, d_b, n, m, max_iter, tol, *none_precond(n));
}
int GMRES_GPU_tran(SpMatrixGPU* Sparse, SpMatrix* spm, dim3*, dim3*, float* d_x, const float* d_b, const int n, const int m,
                   const int max_iter, const float tol, Preconditioner& M, GMRES_GPU_Data& data) {  // src/gmres.cu:2736-2827
  bg_matrix_t A = matrix_for(Sparse->d_val, Sparse->d_rowIndices, Sparse->d_indices, n, spm->numNZEntries);
  if (!data.bg_handle) data.Initilize(m, n);
  int it = max_iter;
  float t = tol;
  return run_gmres((bg_solver_t)data.bg_handle, A, M, d_b, d_x, m, &it, &t);
}
int GMRESilu_GPU(float* d_val, int* d_rowIndices, int* d_indices, int nnz, float* d_x, float* d_b, const int n, const int m,
                 int* max_iter, float* tol, Preconditioner& M) {  // src/gmres.cu:2254-2446
  return run_gmres(solver_for(n, m), matrix_for(d_val, d_rowIndices, d_indices, n, nnz), M, d_b, d_x, m, max_iter, tol);
}

// ------------------------------------------------------------------------------------------ gmres_interface_pg.h
gmresInterfacePGfloat::gmresInterfacePGfloat()
    : matrixSize(0), nnz(0), h_val(nullptr), h_rowPtr(nullptr), h_colIdx(nullptr), d_val(nullptr), d_rowPtr(nullptr), d_colIdx(nullptr),
      x_h(nullptr), x_d(nullptr), xgmres_h(nullptr), rhs_h(nullptr), xgmres_d(nullptr), rhs_d(nullptr), Precond(nullptr), max_it(0),
      tol(0.f), bg_matrix(nullptr), bg_solver(nullptr) {}
void gmresInterfacePGfloat::setPrecondPG(MySpMatrix* A, MySpMatrixDouble* L, MySpMatrixDouble* U, MySpMatrix* D, MySpMatrix* Pr,
                                         MySpMatrix* Pc, MySpMatrix* ls, MySpMatrix* rs) {  // src/gmres_interface_pg.cu:30-60
  matrixSize = A->numRows;
  h_val = A->val; h_rowPtr = A->rowIndices; h_colIdx = A->indices;  // aliased, not copied (caller keeps ownership)
  nnz = h_rowPtr[matrixSize];
  CHK(bg_matrix_create_csr(ctx(), matrixSize, nnz, h_rowPtr, h_colIdx, h_val, BG_F32, BG_HOST, BG_F32, BG_SPMV_AUTO, (bg_matrix_t*)&bg_matrix));
  xgmres_h = (float*)std::malloc(sizeof(float) * (size_t)matrixSize);
  rhs_h = (float*)std::malloc(sizeof(float) * (size_t)matrixSize);
  MyILUPPfloat* p = new MyILUPPfloat();
  p->Initilize(*L, *U, *D, *Pr, *Pc, *ls, *rs, restart);
  Precond = (Preconditioner*)p;
  CHK(bg_solver_create(ctx(), matrixSize, restart, BG_F32, (bg_solver_t*)&bg_solver));
}
int gmresInterfacePGfloat::GMRES_dev_PG() {  // src/gmres_interface_pg.cu:110-139
  max_it = 10000;
  tol = 1e-7f;  // gmres_tol_global, gmres_interface_pg.cu:7
  bg_gmres_opts o;
  bg_gmres_opts_default(&o);
  o.m = restart; o.max_iter = max_it; o.tol = (double)tol; o.orth = orth_mode();
  bg_gmres_result r;
  // H2D of xgmres_h (warm start) and rhs_h, solve, D2H of the solution: one call
  const int rc = bg_gmres_solve_host(ctx(), (bg_solver_t)bg_solver, (bg_matrix_t)bg_matrix, (bg_precond_t)((Preconditioner*)Precond)->bg_handle,
                                     rhs_h, xgmres_h, &o, &r);
  if (rc < 0) die("bg_gmres_solve_host");
  max_it = r.iters;
  tol = (float)r.resid;
  if (rc != 0) std::printf("Failed to converge.\n");
  return rc;
}
int gmresInterfacePGfloat::GMRES_host_PG() {
  no_host_path("gmresInterfacePGfloat::GMRES_host_PG");
  return 1;
}
gmresInterfacePGfloat::~gmresInterfacePGfloat() {
  std::free(xgmres_h);
  std::free(rhs_h);
  if (bg_solver) bg_solver_destroy((bg_solver_t)bg_solver);
  if (bg_matrix) bg_matrix_destroy((bg_matrix_t)bg_matrix);
  delete (Preconditioner*)Precond;
}
