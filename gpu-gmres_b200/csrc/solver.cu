// solver.cu -- contexts, matrices, preconditioners and the restarted GMRES(m) driver behind the
// C ABI of include/b200gmres.h.
//
// The driver replaces GMRESilu_GPU / GMRES_GPU / GMRES_GPU_tran (src/gmres.cu:2254-2446,
// 2567-2732, 2736-2827).  Differences in mechanism, not in algebra:
//  * H, s, cs, sn, the residual estimate and the convergence test live on the device (orth.cu);
//    the host only enqueues kernels and polls a pinned status word every `poll_chunk` inner
//    iterations, one chunk behind the GPU, so there is no host round trip per iteration
//    (the reference blocks (i+2) times per iteration on cublasSdot/Snrm2 results).
//  * every kernel tests the device-side `converged` flag first, so work enqueued past the
//    stopping iteration is a no-op and the iteration count equals the reference's rule exactly.
//  * all workspace is owned by bg_solver_t: no cudaMalloc/cudaFree per call.
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <stdexcept>

#include <type_traits>

#include "internal.h"

namespace bg {

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("BG_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? BG_ERR_NO_GPU : BG_ERR_CUDA;
}
int DevBuf::alloc(size_t b) {
  release();
  if (b == 0) b = 16;
  cudaError_t e = cudaMalloc(&p, b);
  if (e != cudaSuccess) { p = nullptr; set_error("cudaMalloc(%zu) failed: %s", b, cudaGetErrorString(e)); return BG_ERR_ALLOC; }
  bytes = b;
  return BG_OK;
}
int DevBuf::upload(const void* h, size_t b, cudaStream_t s) {
  BG_TRY(alloc(b));
  if (b) BG_CUDA(cudaMemcpyAsync(p, h, b, cudaMemcpyHostToDevice, s));
  return BG_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

template <class T> static int upload_vec(DevBuf& d, const std::vector<T>& v, cudaStream_t s) {
  return d.upload(v.data(), v.size() * sizeof(T), s);
}

// values of any supported dtype -> std::vector<T>
template <class T> static void convert_into(T* o, const void* p, int dt, size_t n) {
  parallel_ranges(n, [=](size_t b, size_t e) {
    if (dt == BG_F64) { const double* s = (const double*)p; for (size_t i = b; i < e; ++i) o[i] = (T)s[i]; }
    else { const float* s = (const float*)p; for (size_t i = b; i < e; ++i) o[i] = (T)s[i]; }
  });
}
template <class T> static std::vector<T> to_vec(const void* p, int dt, size_t n) {
  std::vector<T> o(n);
  convert_into(o.data(), p, dt, n);
  return o;
}
template <class A, class B> static Csr<B> convert_csr(const Csr<A>& M) {
  Csr<B> o;
  o.n = M.n; o.rp = M.rp; o.ci = M.ci;
  o.v.resize(M.v.size());
  for (size_t i = 0; i < M.v.size(); ++i) o.v[i] = (B)M.v[i];
  return o;
}

}  // namespace bg

using namespace bg;

// =========================================================================== misc / context
extern "C" const char* bg_last_error(void) { return g_err; }
extern "C" int bg_version(void) { return 100; }
extern "C" int bg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int bg_context_create(bg_context_t* out, int device) {
  if (!out) return BG_ERR_ARG;
  *out = nullptr;
  int cnt = bg_device_count();
  if (cnt <= 0) { set_error("no CUDA device: libb200gmres has no CPU fallback"); return BG_ERR_NO_GPU; }
  if (device < 0 || device >= cnt) { set_error("device %d out of range (%d devices)", device, cnt); return BG_ERR_ARG; }
  BG_CUDA(cudaSetDevice(device));
  bg_context_t c = new bg_context_s;
  c->device = device;
  cudaDeviceProp prop;
  BG_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  BG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = c;
  return BG_OK;
}
extern "C" int bg_context_destroy(bg_context_t c) {
  if (!c) return BG_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return BG_OK;
}
extern "C" int bg_context_synchronize(bg_context_t c) {
  if (!c) return BG_ERR_ARG;
  BG_CUDA(cudaStreamSynchronize(c->stream));
  return BG_OK;
}
extern "C" int bg_context_stream(bg_context_t c, void** s) {
  if (!c || !s) return BG_ERR_ARG;
  *s = (void*)c->stream;
  return BG_OK;
}
extern "C" int64_t bg_context_launch_count(bg_context_t c) { return c ? c->launches : 0; }

extern "C" int bg_device_malloc(bg_context_t c, size_t bytes, void** p) {
  if (!c || !p) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return BG_ERR_ALLOC; }
  return BG_OK;
}
extern "C" int bg_device_free(bg_context_t c, void* p) {
  if (!c) return BG_ERR_ARG;
  BG_CUDA(cudaFree(p));
  return BG_OK;
}
extern "C" int bg_memcpy_h2d(bg_context_t c, void* d, const void* h, size_t b) {
  if (!c) return BG_ERR_ARG;
  BG_CUDA(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, c->stream));
  BG_CUDA(cudaStreamSynchronize(c->stream));
  return BG_OK;
}
extern "C" int bg_memcpy_d2h(bg_context_t c, void* h, const void* d, size_t b) {
  if (!c) return BG_ERR_ARG;
  BG_CUDA(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, c->stream));
  BG_CUDA(cudaStreamSynchronize(c->stream));
  return BG_OK;
}
extern "C" int bg_host_alloc_pinned(size_t b, void** p) {
  if (!p) return BG_ERR_ARG;
  cudaError_t e = cudaMallocHost(p, b ? b : 16);
  if (e != cudaSuccess) { set_error("cudaMallocHost(%zu): %s", b, cudaGetErrorString(e)); cudaGetLastError(); return BG_ERR_ALLOC; }
  return BG_OK;
}
extern "C" int bg_host_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
  return BG_OK;
}

// =========================================================================== matrix
// B200GMRES_HOST_SETUP=1 keeps the sliced-ELL fill and the block-major reordering on the host (the round-1 path)
static bool host_setup() {
  const char* e = getenv("B200GMRES_HOST_SETUP");
  return e && e[0] == '1';
}
// SpMV plan and storage of a matrix whose CSR arrays are already on the device (A->rp, A->ci, A->val); rp = host row
// pointers.  The sliced-ELL copy is filled on the device (sell_fill_kernel) unless host arrays are handed in.
template <class T>
static int matrix_finish(bg_matrix_t A, const int* rp, int format, const int* h_ci, const T* h_v) {
  cudaStream_t s = A->ctx->stream;
  A->plan = inspect_rows(A->n, rp);
  if (format == BG_SPMV_AUTO) format = A->plan.format;
  A->format = format;
  const double mean = A->plan.mean_row;
  A->lanes = format == BG_SPMV_CSR_WARP ? 32 : mean <= 3 ? 2 : mean <= 6 ? 4 : mean <= 12 ? 8 : mean <= 24 ? 16 : 32;
  if (format != BG_SPMV_SELL) return BG_OK;
  A->sell_rpt = sizeof(T) == 4 ? 4 : 2;
  A->sell_height = 32 * A->sell_rpt;
  std::vector<int64_t> sp;
  sell_slice_ptr(A->n, rp, A->sell_height, 0, sp);
  A->sell_slices = (int)sp.size() - 1;
  // a few short slices (grid boundaries) must not cost the fixed-width kernels: pad every slice to the
  // longest row when that adds at most 5 % of storage
  int min_width = 0;
  if ((int64_t)A->plan.max_row * A->sell_height * A->sell_slices <= sp.back() + sp.back() / 20 && (int64_t)A->plan.max_row * A->sell_height * A->sell_slices != sp.back()) {
    min_width = A->plan.max_row;
    sell_slice_ptr(A->n, rp, A->sell_height, min_width, sp);
  }
  A->sell_width = A->sell_slices > 0 ? (int)((sp[1] - sp[0]) / A->sell_height) : 0;
  for (int q = 0; q < A->sell_slices; ++q)
    if (sp[q + 1] - sp[q] != (int64_t)A->sell_width * A->sell_height) { A->sell_width = 0; break; }
  A->plan.sell_padded = sp.back();
  A->plan.slice = A->sell_height;
  BG_TRY(upload_vec(A->s_ptr, sp, s));
  if (h_ci && h_v) {
    std::vector<int> sc;
    std::vector<T> sv;
    build_sell(A->n, rp, h_ci, h_v, A->sell_height, sp, sc, sv, min_width);
    BG_TRY(upload_vec(A->s_col, sc, s));
    BG_TRY(upload_vec(A->s_val, sv, s));
    BG_CUDA(cudaStreamSynchronize(s));  // sc, sv leave scope
    return BG_OK;
  }
  BG_TRY(A->s_col.alloc(sizeof(int) * (size_t)std::max<int64_t>(sp.back(), 1)));
  BG_TRY(A->s_val.alloc(sizeof(T) * (size_t)std::max<int64_t>(sp.back(), 1)));
  BG_TRY(launch_sell_fill(A->ctx, A->dtype, A->n, A->sell_slices, A->sell_height, A->rp.as<int>(), A->ci.as<int>(), A->val.p,
                          A->s_ptr.as<int64_t>(), A->s_col.as<int>(), A->s_val.p));
  BG_CUDA(cudaStreamSynchronize(s));  // sp leaves scope (its upload is asynchronous)
  return BG_OK;
}

template <class T>
static int matrix_build(bg_matrix_t A, std::vector<int>&& rp_in, std::vector<int>&& ci_in, const void* val, int val_dtype,
                        int format) {
  cudaStream_t s = A->ctx->stream;
  A->requested_format = format;
  // the host copy (kept for the block-Jacobi reordering, build_permuted) takes the index arrays over and receives
  // the values already converted to the compute type: one pass over the caller's data instead of three
  A->h_rp = std::move(rp_in);
  A->h_ci = std::move(ci_in);
  const std::vector<int>&rp = A->h_rp, &ci = A->h_ci;
  const size_t nnz = ci.size();
  A->h_val.resize(nnz * sizeof(T));
  T* const v = reinterpret_cast<T*>(A->h_val.data());
  convert_into(v, val, val_dtype, nnz);
  if (A->ncols < A->n) A->ncols = A->n;
  BG_TRY(upload_vec(A->rp, rp, s));
  BG_TRY(upload_vec(A->ci, ci, s));
  BG_TRY(A->val.upload(v, nnz * sizeof(T), s));
  BG_TRY(matrix_finish<T>(A, rp.data(), format, host_setup() ? ci.data() : nullptr, host_setup() ? v : nullptr));
  if (A->format == BG_SPMV_SELL && A->ncols > A->n) {  // local rows of a row-partitioned matrix: which slices read halo slots
    std::vector<unsigned char> fl(std::max(1, A->sell_slices), 0);
    for (int i = 0; i < A->n; ++i)
      for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (ci[p] >= A->n) { fl[i / A->sell_height] = 1; break; }
    BG_TRY(upload_vec(A->s_halo, fl, s));
  }
  BG_CUDA(cudaStreamSynchronize(s));
  return BG_OK;
}

extern "C" int bg_matrix_create_csr(bg_context_t ctx, int n, int64_t nnz, const int* rowptr, const int* col,
                                    const void* val, int val_dtype, int mem, int compute_dtype, int spmv_format,
                                    bg_matrix_t* out) {
  if (!ctx || !out || n < 0 || nnz < 0 || (n > 0 && (!rowptr || (nnz > 0 && (!col || !val))))) { set_error("bg_matrix_create_csr: bad argument"); return BG_ERR_ARG; }
  if (nnz > 0x7fffffffLL) { set_error("nnz > 2^31-1 needs the row-partitioned path"); return BG_ERR_UNSUPPORTED; }
  *out = nullptr;
  BG_CUDA(cudaSetDevice(ctx->device));
  std::vector<int> rp(n + 1, 0), ci(nnz);
  std::vector<unsigned char> hv;  // staging for device-resident values only: host values are read in place
  if (mem == BG_DEVICE) {
    hv.resize(nnz * dtype_size(val_dtype));
    if (n) BG_CUDA(cudaMemcpy(rp.data(), rowptr, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost));
    if (nnz) { BG_CUDA(cudaMemcpy(ci.data(), col, sizeof(int) * nnz, cudaMemcpyDeviceToHost)); BG_CUDA(cudaMemcpy(hv.data(), val, hv.size(), cudaMemcpyDeviceToHost)); }
    val = hv.data();
  } else {
    if (n) std::memcpy(rp.data(), rowptr, sizeof(int) * (n + 1));
    int* d = ci.data();
    parallel_ranges((size_t)nnz, [=](size_t b, size_t e) { std::copy(col + b, col + e, d + b); });
  }
  if (n && (rp[0] != 0 || rp[n] != nnz)) { set_error("rowptr[0] must be 0 and rowptr[n] == nnz"); return BG_ERR_ARG; }
  return matrix_create_internal(ctx, n, n, std::move(rp), std::move(ci), val, val_dtype, compute_dtype, spmv_format, out);
}
namespace bg {
int matrix_create_internal(bg_context_t ctx, int n, int ncols, std::vector<int> rp, std::vector<int> ci,
                           const void* val, int val_dtype, int compute_dtype, int spmv_format, bg_matrix_t* out) {
  const int64_t nnz = (int64_t)ci.size();
  for (int64_t p = 0; p < nnz; ++p)
    if (ci[p] < 0 || ci[p] >= ncols) { set_error("column index out of range at entry %lld", (long long)p); return BG_ERR_ARG; }
  bg_matrix_t A = new bg_matrix_s;
  A->ctx = ctx; A->n = n; A->ncols = ncols; A->nnz = nnz; A->dtype = compute_dtype;
  int rc = compute_dtype == BG_F32 ? matrix_build<float>(A, std::move(rp), std::move(ci), val, val_dtype, spmv_format)
                                   : matrix_build<double>(A, std::move(rp), std::move(ci), val, val_dtype, spmv_format);
  if (rc < 0) { delete A; return rc; }
  *out = A;
  return BG_OK;
}
}  // namespace bg
extern "C" int bg_matrix_destroy(bg_matrix_t A) { delete A; return BG_OK; }
extern "C" int bg_matrix_get_info(bg_matrix_t A, bg_matrix_info* o) {
  if (!A || !o) return BG_ERR_ARG;
  std::memset(o, 0, sizeof(*o));
  o->n = A->n; o->nnz = A->nnz; o->dtype = A->dtype; o->format = A->format;
  o->max_row_nnz = A->plan.max_row; o->mean_row_nnz = A->plan.mean_row;
  for (int i = 0; i < 8; ++i) o->hist[i] = A->plan.hist[i];
  o->sell_padded_nnz = A->format == BG_SPMV_SELL ? A->plan.sell_padded : 0;
  o->sell_slice = A->format == BG_SPMV_SELL ? A->sell_height : 0;
  const int64_t s = dtype_size(A->dtype);
  o->spmv_bytes = (s + 4) * A->nnz + 4 * ((int64_t)A->n + 1) + 2 * s * A->n;
  return BG_OK;
}
extern "C" int bg_spmv(bg_context_t ctx, bg_matrix_t A, const void* x, void* y) {
  if (!ctx || !A || !x || !y) return BG_ERR_ARG;
  return launch_spmv(ctx, A, 1.0, x, 0.0, nullptr, y, false, nullptr);
}
extern "C" int bg_spmv_axpby(bg_context_t ctx, bg_matrix_t A, double alpha, const void* x, double beta, const void* b, void* y) {
  if (!ctx || !A || !x || !y || !b) return BG_ERR_ARG;
  return launch_spmv(ctx, A, alpha, x, beta, b, y, true, nullptr);
}

// =========================================================================== host artefacts
extern "C" int bg_host_factor_levels(int n, const int* rp, const int* ci, const float* v, int* level) {
  if (n < 0 || (n && (!rp || !ci || !v || !level))) return BG_ERR_ARG;
  factor_levels(n, rp, ci, v, level);
  return BG_OK;
}
extern "C" int bg_host_trsv_levels(int is_lower, int n, const int* rp, const int* ci, const void* v, int dt, int* level) {
  if (n < 0 || (n && (!rp || !ci || !v || !level))) return BG_ERR_ARG;
  if (dt == BG_F64) trsv_levels<double>(is_lower != 0, n, rp, ci, (const double*)v, level);
  else trsv_levels<float>(is_lower != 0, n, rp, ci, (const float*)v, level);
  return BG_OK;
}
extern "C" int bg_host_bucket_levels(int n, const int* level, int* n_levels, int* levelPtr, int* nodes) {
  if (n < 0 || (n && !level)) return BG_ERR_ARG;
  std::vector<int> lp, nd;
  const int nl = bucket_levels(n, level, lp, nd);
  if (n_levels) *n_levels = nl;
  if (levelPtr) std::copy(lp.begin(), lp.end(), levelPtr);
  if (nodes) std::copy(nd.begin(), nd.end(), nodes);
  return BG_OK;
}
enum { FACTOR_ILU0 = 0, FACTOR_ILUT = 1, FACTOR_ILUK = 2 };  // `lfil`: ILUT row budget / ILU(k) level
template <class T>
static void factor_dispatch(int method, int n, const int* rp, const int* ci, const void* v, int dt, double tau, int lfil,
                            const int* blk, Factors<T>& F) {
  std::vector<T> vv = to_vec<T>(v, dt, (size_t)rp[n]);
  if (method == FACTOR_ILUT) ilut<T>(n, rp, ci, vv.data(), tau, lfil, blk, F);
  else if (method == FACTOR_ILUK) iluk<T>(n, rp, ci, vv.data(), lfil, blk, F);
  else ilu0<T>(n, rp, ci, vv.data(), blk, F);
}
static int host_factor(int method, int n, const int* rp, const int* ci, const void* v, int dt, double tau, int lfil,
                       const int* blk, bg_factors_t* out) {
  if (!out || n < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  bg_factors_t F = new bg_factors_s;
  F->n = n; F->dtype = dt;
  try {
    if (dt == BG_F64) factor_dispatch<double>(method, n, rp, ci, v, dt, tau, lfil, blk, F->f64);
    else factor_dispatch<float>(method, n, rp, ci, v, dt, tau, lfil, blk, F->f32);
  } catch (const std::exception& e) { set_error("factorisation failed: %s", e.what()); delete F; return BG_ERR_ARG; }
  *out = F;
  return BG_OK;
}
extern "C" int bg_host_ilu0(int n, const int* rp, const int* ci, const void* v, int dt, const int* blk, bg_factors_t* F) {
  return host_factor(FACTOR_ILU0, n, rp, ci, v, dt, 0, 0, blk, F);
}
extern "C" int bg_host_ilut(int n, const int* rp, const int* ci, const void* v, int dt, double tau, int lfil, const int* blk, bg_factors_t* F) {
  if (lfil < 0 || tau < 0) return BG_ERR_ARG;
  return host_factor(FACTOR_ILUT, n, rp, ci, v, dt, tau, lfil, blk, F);
}
extern "C" int bg_host_iluk(int n, const int* rp, const int* ci, const void* v, int dt, int level, const int* blk, bg_factors_t* F) {
  if (level < 0) { set_error("bg_host_iluk: level of fill must be >= 0"); return BG_ERR_ARG; }
  return host_factor(FACTOR_ILUK, n, rp, ci, v, dt, 0, level, blk, F);
}
extern "C" int bg_host_ainv(int n, const int* rp, const int* ci, const void* v, int dt, double droptol, bg_factors_t* out) {
  if (!out || n < 0 || droptol < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  bg_factors_t F = new bg_factors_s;
  F->n = n; F->dtype = dt;
  try {
    if (dt == BG_F64) ainv_biconjugation<double>(n, rp, ci, (const double*)v, droptol, F->f64);
    else ainv_biconjugation<float>(n, rp, ci, (const float*)v, droptol, F->f32);
  } catch (const std::exception& e) { set_error("AINV construction failed: %s", e.what()); delete F; return BG_ERR_ARG; }
  *out = F;
  return BG_OK;
}
extern "C" int bg_factors_get_info(bg_factors_t F, bg_factors_info* o) {
  if (!F || !o) return BG_ERR_ARG;
  o->n = F->n; o->dtype = F->dtype;
  if (F->dtype == BG_F64) { o->nnzL = F->f64.L.v.size(); o->nnzU = F->f64.U.v.size(); o->zero_pivots = F->f64.zero_pivots; }
  else { o->nnzL = F->f32.L.v.size(); o->nnzU = F->f32.U.v.size(); o->zero_pivots = F->f32.zero_pivots; }
  return BG_OK;
}
template <class T> static void factors_copy(const Factors<T>& f, int* lp, int* li, void* lv, int* up, int* ui, void* uv) {
  if (lp) std::copy(f.L.rp.begin(), f.L.rp.end(), lp);
  if (li) std::copy(f.L.ci.begin(), f.L.ci.end(), li);
  if (lv) std::memcpy(lv, f.L.v.data(), f.L.v.size() * sizeof(T));
  if (up) std::copy(f.U.rp.begin(), f.U.rp.end(), up);
  if (ui) std::copy(f.U.ci.begin(), f.U.ci.end(), ui);
  if (uv) std::memcpy(uv, f.U.v.data(), f.U.v.size() * sizeof(T));
}
extern "C" int bg_factors_get(bg_factors_t F, int* lp, int* li, void* lv, int* up, int* ui, void* uv) {
  if (!F) return BG_ERR_ARG;
  if (F->dtype == BG_F64) factors_copy(F->f64, lp, li, lv, up, ui, uv);
  else factors_copy(F->f32, lp, li, lv, up, ui, uv);
  return BG_OK;
}
extern "C" int bg_factors_destroy(bg_factors_t F) { delete F; return BG_OK; }
extern "C" int bg_host_partition(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, int64_t nn, int64_t* ps,
                                 int64_t* np, int64_t* pinv, int64_t* q) {
  if (m < 0 || npart < 1 || !Ap || !ps || !np || !pinv || !q) return BG_ERR_ARG;
  partition_ref(m, Ap, Ai, npart, nn, ps, np, pinv, q);
  return BG_OK;
}
extern "C" int bg_host_partition3_post(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, const int64_t* part,
                                       int64_t* ps, int64_t* np, int64_t* pinv, int64_t* q) {
  if (m < 0 || npart < 1 || !Ap || !part || !ps || !np || !pinv || !q) return BG_ERR_ARG;
  partition3_post(m, Ap, Ai, npart, part, ps, np, pinv, q);
  return BG_OK;
}
extern "C" int bg_host_block_partition(int method, int n, const int* rp, const int* ci, int nblocks, const int* dims, int* blk) {
  if (n < 0 || !blk || (n && (!rp || !ci))) return BG_ERR_ARG;
  if ((method == BG_PART_GRID2D || method == BG_PART_GRID3D) && !dims) return BG_ERR_ARG;
  try { return block_partition(method, n, rp, ci, nblocks, dims, blk); }
  catch (const std::exception& e) { set_error("block partition failed: %s", e.what()); return BG_ERR_ARG; }
}

// =========================================================================== preconditioners
extern "C" void bg_precond_opts_default(bg_precond_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->trsv_schedule = BG_TRSV_LEVEL;
  o->block_method = BG_PART_METIS;
  o->block_rows = 1024;
  o->factor_dtype = -1;
  o->use_tma = 1;
  o->ilut_lfil = 10;
  o->ilut_tau = 1e-3;
  o->chunk_bytes = 4096;
  o->stages = 2;
  o->warps_per_cta = 3;
  o->lattice = 1;
  o->device_factor = 1;
}

static uint64_t g_order_serial = 0;
template <class T> static int upload_syncfree_stream(bg_context_t ctx, const SyncFreeStream<T>& h, SyncFreeDev& d) {
  cudaStream_t s = ctx->stream;
  d.n = h.n; d.n_levels = h.n_levels; d.n_slices = h.n_slices; d.max_K = h.max_K; d.max_level_rows = h.max_level_rows;
  BG_TRY(upload_vec(d.meta, h.meta, s));
  BG_TRY(upload_vec(d.rows, h.rows, s)); BG_TRY(upload_vec(d.cols, h.cols, s));
  BG_TRY(upload_vec(d.vals, h.vals, s)); BG_TRY(upload_vec(d.diag, h.diag, s));
  d.h_rows = h.rows;
  d.bytes = (int64_t)h.cols.size() * (4 + sizeof(T)) + (int64_t)h.rows.size() * (4 + sizeof(T)) + (int64_t)h.meta.size() * 4;
  BG_CUDA(cudaStreamSynchronize(s));
  return BG_OK;
}
template <class T> static int upload_lattice_global(bg_context_t ctx, const LatticeGlobal<T>& h, LatticeGlobalDev& d) {
  cudaStream_t s = ctx->stream;
  d.nx = h.nx; d.ny = h.ny; d.strips = h.strips;
  BG_TRY(upload_vec(d.lower, h.lower, s)); BG_TRY(upload_vec(d.upper, h.upper, s)); BG_TRY(upload_vec(d.udinv, h.udinv, s));
  const size_t bnd = (size_t)2 * h.strips * h.ny * 16;
  BG_TRY(d.bnd.alloc(bnd));
  BG_CUDA(cudaMemsetAsync(d.bnd.p, 0, bnd, s));
  d.bytes = (int64_t)(h.lower.size() + h.upper.size() + h.udinv.size()) * sizeof(T);
  d.on = true;
  BG_CUDA(cudaStreamSynchronize(s));
  return BG_OK;
}
template <class T> static int upload_fused_stream(bg_context_t ctx, const FusedBlockStream<T>& h, FusedBlockDev& d) {
  cudaStream_t s = ctx->stream;
  d.n = h.n; d.nblocks = h.nblocks; d.max_rows = h.max_rows; d.max_levels_L = h.max_levels_L; d.max_levels_U = h.max_levels_U;
  d.max_chunk_bytes = h.max_chunk_bytes; d.max_chunks = h.max_chunks; d.total_steps = h.total_steps;
  d.h_perm = h.perm;
  d.lattice = h.lattice;
  d.serial = ++g_order_serial;
  std::vector<int> inv(h.n);
  for (int i = 0; i < h.n; ++i) inv[h.perm[i]] = i;
  BG_TRY(upload_vec(d.perm, h.perm, s)); BG_TRY(upload_vec(d.iperm, inv, s));
  BG_TRY(d.blk.upload(h.blk.data(), h.blk.size() * sizeof(BlockDesc), s));
  // fixed-stride chunk table so that a block's descriptor and chunk offsets are independent loads
  d.tab_stride = (h.max_chunks + 1 + 3) / 4 * 4;
  std::vector<uint32_t> tab((size_t)h.nblocks * d.tab_stride, 0u);
  for (int b = 0; b < h.nblocks; ++b)
    for (int c = 0; c <= h.blk[b].nchunks; ++c) tab[(size_t)b * d.tab_stride + c] = h.chunk_off[h.blk[b].chunk0 + c];
  BG_TRY(upload_vec(d.chunk_tab, tab, s)); BG_TRY(upload_vec(d.stream, h.stream, s));
  BG_TRY(upload_vec(d.dinv, h.dinv, s));
  d.h_dinv.assign(h.dinv.begin(), h.dinv.end());
  d.bytes = (int64_t)h.stream.size() + (int64_t)h.blk.size() * sizeof(BlockDesc) + (int64_t)tab.size() * 4;
  BG_CUDA(cudaStreamSynchronize(s));
  return BG_OK;
}

// factors (host, type FH) -> device schedules with storage type FT
template <class FH, class FT>
static int precond_install(bg_precond_t M, const Csr<FH>& L, const Csr<FH>& U, bool unit_lower, const int* block_of_row,
                           int nblocks) {
  // same storage type: use the caller's factors in place (a copy of 10 M-row factors costs ~0.3 s of setup)
  Csr<FT> Lc, Uc;
  if constexpr (!std::is_same<FH, FT>::value) { Lc = convert_csr<FH, FT>(L); Uc = convert_csr<FH, FT>(U); }
  const Csr<FT>& Lf = [&]() -> const Csr<FT>& { if constexpr (std::is_same<FH, FT>::value) return L; else return Lc; }();
  const Csr<FT>& Uf = [&]() -> const Csr<FT>& { if constexpr (std::is_same<FH, FT>::value) return U; else return Uc; }();
  M->nnzL = (int64_t)L.v.size(); M->nnzU = (int64_t)U.v.size();
  M->unit_lower = unit_lower;
  try {
    if (M->schedule == BG_TRSV_BLOCK) {
      FusedBlockStream<FT> fs;
      pack_fused_block_stream<FT>(unit_lower, Lf, Uf, nblocks, block_of_row, M->fb_chunk_bytes, fs, M->allow_lattice);
      BG_TRY(upload_fused_stream(M->ctx, fs, M->fb));
    } else {
      // the parity artefact: getLevel of both factors (src/gmres.cu:1033-1077)
      M->levelL.resize(M->n); M->levelU.resize(M->n);
      trsv_levels<FT>(true, M->n, Lf.rp.data(), Lf.ci.data(), Lf.v.data(), M->levelL.data());
      trsv_levels<FT>(false, M->n, Uf.rp.data(), Uf.ci.data(), Uf.v.data(), M->levelU.data());
      int nx = 0;
      if (M->allow_lattice && unit_lower && M->kind == BG_PRECOND_LU && detect_lattice2d<FT>(Lf, Uf, &nx) &&
          (nx + 31) / 32 <= lattice_global_capacity(M->ctx, M->dtype, M->factor_dtype)) {
        LatticeGlobal<FT> lg;
        pack_lattice_global<FT>(Lf, Uf, nx, lg);
        BG_TRY(upload_lattice_global(M->ctx, lg, M->lg));
        M->lg_levels = nx + M->n / nx - 1;
      } else {
        SyncFreeStream<FT> ll, lu;
        pack_syncfree_stream<FT>(true, unit_lower, Lf, ll);
        pack_syncfree_stream<FT>(false, false, Uf, lu);
        BG_TRY(upload_syncfree_stream(M->ctx, ll, M->lsL));
        BG_TRY(upload_syncfree_stream(M->ctx, lu, M->lsU));
      }
    }
  } catch (const std::exception& e) { set_error("schedule construction failed: %s", e.what()); return BG_ERR_ARG; }
  if (M->schedule == BG_TRSV_LEVEL && !M->lg.on) {  // flag-in-data packets of the sync-free solves (sequence 0 = never written)
    BG_TRY(M->xsL.alloc((size_t)M->n * 16));
    BG_TRY(M->xsU.alloc((size_t)M->n * 16));
    BG_CUDA(cudaMemset(M->xsL.p, 0, (size_t)M->n * 16));
    BG_CUDA(cudaMemset(M->xsU.p, 0, (size_t)M->n * 16));
  }
  BG_TRY(M->tmp.alloc((size_t)M->n * 8));
  if (M->kind == BG_PRECOND_SPLIT) BG_TRY(M->tmp2.alloc((size_t)M->n * 8));
  return BG_OK;
}
template <class FH>
static int precond_install_dt(bg_precond_t M, const Csr<FH>& L, const Csr<FH>& U, bool unit_lower, const int* blk, int nb) {
  return M->factor_dtype == BG_F64 ? precond_install<FH, double>(M, L, U, unit_lower, blk, nb)
                                   : precond_install<FH, float>(M, L, U, unit_lower, blk, nb);
}

static bg_precond_t precond_new(bg_context_t ctx, int kind, int n, int dtype) {
  bg_precond_t M = new bg_precond_s;
  { const char* e = getenv("BG_TRSV_LATTICE"); M->allow_lattice = !(e && e[0] == '0'); }
  M->ctx = ctx; M->kind = kind; M->n = n; M->dtype = dtype; M->factor_dtype = dtype;
  return M;
}

extern "C" int bg_precond_create_none(bg_context_t ctx, int n, int dtype, bg_precond_t* out) {
  if (!ctx || !out || n < 0) return BG_ERR_ARG;
  *out = precond_new(ctx, BG_PRECOND_NONE, n, dtype);
  return BG_OK;
}

extern "C" int bg_precond_create_diag(bg_context_t ctx, int n, const int* rp, const int* ci, const void* v, int vdt,
                                      int dtype, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  // MyDIAG::Initilize (src/preconditioner.cu:500-540): the stored entry closest to the diagonal
  // (first one on ties), inverted in the vector precision.
  std::vector<double> d(n);
  for (int i = 0; i < n; ++i) {
    int dist = 2 * n;
    double dv = 0.0;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int dd = std::abs(ci[p] - i);
      if (dd < dist) { dist = dd; dv = vdt == BG_F64 ? ((const double*)v)[p] : (double)((const float*)v)[p]; }
    }
    if (is_zero(dv)) { set_error("zero diagonal in row %d", i); return BG_ERR_ZERO_PIVOT; }
    d[i] = dtype == BG_F32 ? (double)(1.0f / (float)dv) : 1.0 / dv;
  }
  bg_precond_t M = precond_new(ctx, BG_PRECOND_DIAG, n, dtype);
  int rc = dtype == BG_F32 ? upload_vec(M->dinv, to_vec<float>(d.data(), BG_F64, n), ctx->stream) : upload_vec(M->dinv, d, ctx->stream);
  if (rc < 0) { delete M; return rc; }
  cudaStreamSynchronize(ctx->stream);
  *out = M;
  return BG_OK;
}

extern "C" int bg_precond_create_ainv(bg_context_t ctx, int n, const int* p1, const int* c1, const void* v1, const void* diag,
                                      const int* p2, const int* c2, const void* v2, int vdt, int dtype, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!p1 || !c1 || !v1)) || (p2 && (!c2 || !v2))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_t M = precond_new(ctx, BG_PRECOND_AINV, n, dtype);
  int rc = bg_matrix_create_csr(ctx, n, n ? p1[n] : 0, p1, c1, v1, vdt, BG_HOST, dtype, BG_SPMV_AUTO, &M->m1);
  if (rc >= 0 && p2) rc = bg_matrix_create_csr(ctx, n, n ? p2[n] : 0, p2, c2, v2, vdt, BG_HOST, dtype, BG_SPMV_AUTO, &M->m2);
  if (rc >= 0 && diag)
    rc = dtype == BG_F32 ? upload_vec(M->dinv, to_vec<float>(diag, vdt, n), ctx->stream) : upload_vec(M->dinv, to_vec<double>(diag, vdt, n), ctx->stream);
  if (rc >= 0) rc = M->tmp.alloc((size_t)n * 8);
  if (rc >= 0) rc = M->tmp2.alloc((size_t)n * 8);
  cudaStreamSynchronize(ctx->stream);
  if (rc < 0) { delete M; return rc; }
  M->nnzL = M->m1->nnz; M->nnzU = M->m2 ? M->m2->nnz : 0;
  *out = M;
  return BG_OK;
}

static int make_blocks(const bg_precond_opts& o, int n, const int* rp, const int* ci, std::vector<int>& blk, int* nblocks) {
  blk.assign(n, 0);
  const int rows = std::max(32, o.block_rows);
  int nb = std::max(1, (n + rows - 1) / rows);
  int dims[6] = {o.dims[0], o.dims[1], o.dims[2], o.dims[3], o.dims[4], o.dims[5]};
  if (o.block_method == BG_PART_GRID2D && dims[3] <= 0) { dims[3] = dims[4] = std::max(1, (int)std::floor(std::sqrt((double)rows))); }
  if (o.block_method == BG_PART_GRID3D && dims[3] <= 0) { dims[3] = dims[4] = dims[5] = std::max(1, (int)std::floor(std::cbrt((double)rows))); }
  try { nb = block_partition(o.block_method, n, rp, ci, nb, dims, blk.data()); }
  catch (const std::exception& e) { set_error("block partition failed: %s", e.what()); return BG_ERR_ARG; }
  *nblocks = nb;
  return BG_OK;
}

template <class T>
static int create_factored(bg_context_t ctx, int method, int n, const int* rp, const int* ci, const void* v, int vdt,
                           int dtype, const bg_precond_opts& o, bg_precond_t* out, int iluk_level = 0) {
  const bool use_ilut = method != FACTOR_ILU0;  // anything but ILU(0) of grid tiles is factored on the host
  std::vector<int> blk;
  int nb = 1;
  const bool block = o.trsv_schedule == BG_TRSV_BLOCK;
  if (block) BG_TRY(make_blocks(o, n, rp, ci, blk, &nb));
  bg_precond_t M = precond_new(ctx, BG_PRECOND_LU, n, dtype);
  M->schedule = o.trsv_schedule;
  M->factor_dtype = o.factor_dtype < 0 ? dtype : o.factor_dtype;
  M->use_tma = o.use_tma != 0;
  M->allow_lattice = M->allow_lattice && o.lattice != 0;
  M->fb_chunk_bytes = o.chunk_bytes > 0 ? o.chunk_bytes : 4096;
  M->fb.stages = std::min(8, std::max(2, o.stages > 0 ? o.stages : 2));
  M->fb.warps_per_cta = std::min(16, std::max(1, o.warps_per_cta > 0 ? o.warps_per_cta : 3));
  // ILU(0) of grid tiles: factor AND pack on the device (factor.cu); anything else: host factorisation + host packers
  static const bool dev_factor = [] { const char* e = getenv("BG_DEVICE_FACTOR"); return !(e && e[0] == '0'); }();
  if (dev_factor && !use_ilut && block && o.block_method == BG_PART_GRID2D && M->allow_lattice && M->factor_dtype == dtype && o.device_factor != 0) {
    int dims[6] = {o.dims[0], o.dims[1], o.dims[2], o.dims[3], o.dims[4], o.dims[5]};
    if (dims[3] <= 0) dims[3] = dims[4] = std::max(1, (int)std::floor(std::sqrt((double)std::max(32, o.block_rows))));
    const int stages = M->fb.stages, warps = M->fb.warps_per_cta;
    int rc = lattice_tiles_factor_device<T>(ctx, M, n, rp, ci, v, vdt, dims, ++g_order_serial);
    if (rc < 0) { delete M; return rc; }
    if (rc == BG_OK) {
      M->fb.stages = stages; M->fb.warps_per_cta = warps;
      M->device_factored = true;
      rc = M->tmp.alloc((size_t)n * 8);
      if (rc < 0) { delete M; return rc; }
      *out = M;
      return BG_OK;
    }
    M->fb.stages = stages; M->fb.warps_per_cta = warps;  // not a 5-point grid in this numbering: the general path
  }
  Factors<T> F;
  try { factor_dispatch<T>(method, n, rp, ci, v, vdt, o.ilut_tau, method == FACTOR_ILUK ? iluk_level : o.ilut_lfil, block ? blk.data() : nullptr, F); }
  catch (const std::exception& e) { set_error("factorisation failed: %s", e.what()); delete M; return BG_ERR_ARG; }
  int rc = precond_install_dt<T>(M, F.L, F.U, true, block ? blk.data() : nullptr, nb);
  if (rc < 0) { delete M; return rc; }
  *out = M;
  return BG_OK;
}

extern "C" int bg_precond_create_ilu0(bg_context_t ctx, int n, const int* rp, const int* ci, const void* v, int vdt,
                                      int dtype, const bg_precond_opts* o, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_opts d;
  bg_precond_opts_default(&d);
  if (o) d = *o;
  return dtype == BG_F32 ? create_factored<float>(ctx, FACTOR_ILU0, n, rp, ci, v, vdt, dtype, d, out)
                         : create_factored<double>(ctx, FACTOR_ILU0, n, rp, ci, v, vdt, dtype, d, out);
}
extern "C" int bg_precond_create_ilut(bg_context_t ctx, int n, const int* rp, const int* ci, const void* v, int vdt,
                                      int dtype, const bg_precond_opts* o, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_opts d;
  bg_precond_opts_default(&d);
  if (o) d = *o;
  return dtype == BG_F32 ? create_factored<float>(ctx, FACTOR_ILUT, n, rp, ci, v, vdt, dtype, d, out)
                         : create_factored<double>(ctx, FACTOR_ILUT, n, rp, ci, v, vdt, dtype, d, out);
}
extern "C" int bg_precond_create_iluk(bg_context_t ctx, int n, const int* rp, const int* ci, const void* v, int vdt,
                                      int dtype, int level, const bg_precond_opts* o, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || level < 0 || (n && (!rp || !ci || !v))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_opts d;
  bg_precond_opts_default(&d);
  if (o) d = *o;
  return dtype == BG_F32 ? create_factored<float>(ctx, FACTOR_ILUK, n, rp, ci, v, vdt, dtype, d, out, level)
                         : create_factored<double>(ctx, FACTOR_ILUK, n, rp, ci, v, vdt, dtype, d, out, level);
}

template <class T> static Csr<T> csr_from(int n, const int* rp, const int* ci, const void* v, int vdt) {
  Csr<T> M;
  M.n = n;
  M.rp.assign(rp, rp + n + 1);
  M.ci.assign(ci, ci + rp[n]);
  M.v = to_vec<T>(v, vdt, (size_t)rp[n]);
  return M;
}
// structural check of caller-supplied factors: L diagonal last, U diagonal first
template <class T> static int check_layout(const Csr<T>& L, const Csr<T>& U) {
  for (int i = 0; i < L.n; ++i) {
    if (L.rp[i + 1] <= L.rp[i] || L.ci[L.rp[i + 1] - 1] != i) { set_error("L row %d: diagonal must be the last entry", i); return BG_ERR_ARG; }
    if (U.rp[i + 1] <= U.rp[i] || U.ci[U.rp[i]] != i) { set_error("U row %d: diagonal must be the first entry", i); return BG_ERR_ARG; }
    for (int p = L.rp[i]; p < L.rp[i + 1] - 1; ++p) if (L.ci[p] >= i) { set_error("L row %d has an entry right of the diagonal", i); return BG_ERR_ARG; }
    for (int p = U.rp[i] + 1; p < U.rp[i + 1]; ++p) if (U.ci[p] <= i) { set_error("U row %d has an entry left of the diagonal", i); return BG_ERR_ARG; }
  }
  return BG_OK;
}

extern "C" int bg_precond_create_lu(bg_context_t ctx, int n, const int* lp, const int* li, const void* lv, const int* up,
                                    const int* ui, const void* uv, int fdt, int dtype, const bg_precond_opts* o,
                                    bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!lp || !li || !lv || !up || !ui || !uv))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_opts d;
  bg_precond_opts_default(&d);
  if (o) d = *o;
  if (d.trsv_schedule == BG_TRSV_BLOCK) { set_error("caller-supplied global factors need BG_TRSV_LEVEL"); return BG_ERR_UNSUPPORTED; }
  Csr<double> L = csr_from<double>(n, lp, li, lv, fdt), U = csr_from<double>(n, up, ui, uv, fdt);
  BG_TRY(check_layout(L, U));
  bg_precond_t M = precond_new(ctx, BG_PRECOND_LU, n, dtype);
  M->schedule = BG_TRSV_LEVEL;
  M->factor_dtype = d.factor_dtype < 0 ? fdt : d.factor_dtype;
  M->allow_lattice = M->allow_lattice && d.lattice != 0;
  int rc = precond_install_dt<double>(M, L, U, true, nullptr, 1);
  if (rc < 0) { delete M; return rc; }
  *out = M;
  return BG_OK;
}

extern "C" int bg_precond_create_split(bg_context_t ctx, int n, const int* lp, const int* li, const void* lv,
                                       const int* up, const int* ui, const void* uv, int fdt, const void* middle,
                                       const void* lscale, const void* rscale, int sdt, const int* permRow,
                                       const int* permCol, int dtype, const bg_precond_opts* o, bg_precond_t* out) {
  if (!ctx || !out || n < 0 || (n && (!lp || !li || !lv || !up || !ui || !uv || !middle || !lscale || !rscale || !permRow || !permCol))) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_precond_opts d;
  bg_precond_opts_default(&d);
  if (o) d = *o;
  if (d.trsv_schedule == BG_TRSV_BLOCK) { set_error("the split (ILU++) preconditioner needs BG_TRSV_LEVEL"); return BG_ERR_UNSUPPORTED; }
  Csr<double> L = csr_from<double>(n, lp, li, lv, fdt), U = csr_from<double>(n, up, ui, uv, fdt);
  BG_TRY(check_layout(L, U));
  std::vector<int> inv(n, -1);
  for (int i = 0; i < n; ++i) {
    if (permRow[i] < 0 || permRow[i] >= n || permCol[i] < 0 || permCol[i] >= n || inv[permCol[i]] != -1) { set_error("permRow/permCol is not a permutation"); return BG_ERR_ARG; }
    inv[permCol[i]] = i;
  }
  bg_precond_t M = precond_new(ctx, BG_PRECOND_SPLIT, n, dtype);
  M->schedule = BG_TRSV_LEVEL;
  M->factor_dtype = d.factor_dtype < 0 ? fdt : d.factor_dtype;
  cudaStream_t s = ctx->stream;
  int rc = precond_install_dt<double>(M, L, U, false, nullptr, 1);
  auto up_scale = [&](DevBuf& b, const void* p) {
    return dtype == BG_F32 ? upload_vec(b, to_vec<float>(p, sdt, n), s) : upload_vec(b, to_vec<double>(p, sdt, n), s);
  };
  if (rc >= 0) rc = up_scale(M->middle, middle);
  if (rc >= 0) rc = up_scale(M->lscale, lscale);
  if (rc >= 0) rc = up_scale(M->rscale, rscale);
  if (rc >= 0) rc = M->permRow.upload(permRow, sizeof(int) * n, s);
  if (rc >= 0) rc = M->permCol.upload(permCol, sizeof(int) * n, s);
  if (rc >= 0) rc = upload_vec(M->invPermCol, inv, s);
  if (rc >= 0) {  // DevPrecond_left reads in[permRow[r]], DevPrecond_right writes out[invPermCol[r]]: composed with the slot order
    std::vector<int> src(n), dst(n);
    for (int k = 0; k < n; ++k) { src[k] = permRow[M->lsL.h_rows[k]]; dst[k] = inv[M->lsU.h_rows[k]]; }
    rc = upload_vec(M->lsL.src_idx, src, s);
    if (rc >= 0) rc = upload_vec(M->lsU.dst_idx, dst, s);
  }
  if (rc >= 0) rc = upload_vec(M->u_rp, U.rp, s);
  if (rc >= 0) rc = upload_vec(M->u_ci, U.ci, s);
  if (rc >= 0) rc = M->factor_dtype == BG_F64 ? upload_vec(M->u_val, U.v, s) : upload_vec(M->u_val, to_vec<float>(U.v.data(), BG_F64, U.v.size()), s);
  cudaStreamSynchronize(s);
  if (rc < 0) { delete M; return rc; }
  *out = M;
  return BG_OK;
}

namespace bg {
// internal apply with the solver's skip flag
int precond_apply(bg_context_t ctx, bg_precond_t M, int which, const void* in, void* out, const int* skip, bool permuted) {
  const int n = M->n;
  switch (M->kind) {
    case BG_PRECOND_NONE: return launch_copy(ctx, M->dtype, n, in, out, skip);
    case BG_PRECOND_DIAG:
      if (which == BG_PC_RIGHT || which == BG_PC_START) return launch_copy(ctx, M->dtype, n, in, out, skip);
      return launch_diag_scale(ctx, M->dtype, n, in, M->dinv.p, out, skip);
    case BG_PRECOND_AINV: {
      if (which == BG_PC_RIGHT || which == BG_PC_START) return launch_copy(ctx, M->dtype, n, in, out, skip);
      const bool two = M->m2 != nullptr, scaled = M->dinv.p != nullptr;
      void* t1 = two || scaled ? M->tmp.p : out;
      BG_TRY(launch_spmv(ctx, M->m1, 1.0, in, 0.0, nullptr, t1, false, skip));
      if (scaled) BG_TRY(launch_diag_scale(ctx, M->dtype, n, t1, M->dinv.p, two ? M->tmp2.p : out, skip));
      if (two) BG_TRY(launch_spmv(ctx, M->m2, 1.0, scaled ? M->tmp2.p : t1, 0.0, nullptr, out, false, skip));
      return BG_OK;
    }
    case BG_PRECOND_LU: {
      if (which == BG_PC_RIGHT || which == BG_PC_START) return launch_copy(ctx, M->dtype, n, in, out, skip);
      if (M->schedule == BG_TRSV_BLOCK)  // `permuted`: vectors already in the block-major order of the stream
        return launch_trsv_block(ctx, M->fb, M->dtype, M->factor_dtype, !permuted, in, out, skip);
      if (M->lg.on) return launch_trsv_lattice_global(ctx, M->lg, M->dtype, M->factor_dtype, in, M->tmp.p, out, skip);
      TrsvXform xf;
      BG_TRY(launch_trsv_level(ctx, M->lsL, M->dtype, M->factor_dtype, M->unit_lower, in, M->xsL.p, M->tmp.p, xf, skip));
      return launch_trsv_level(ctx, M->lsU, M->dtype, M->factor_dtype, false, M->tmp.p, M->xsU.p, out, xf, skip);
    }
    case BG_PRECOND_SPLIT: {
      TrsvXform xf;
      if (which == BG_PC_LEFT || which == BG_PC_RHS) {
        xf.pro = 1; xf.perm = M->permRow.as<int>(); xf.scale_pro = M->lscale.p;
        return launch_trsv_level(ctx, M->lsL, M->dtype, M->factor_dtype, false, in, M->xsL.p, out, xf, skip);
      }
      if (which == BG_PC_RIGHT) {
        xf.pro = 2; xf.scale_pro = M->middle.p; xf.epi = 1; xf.perm = M->invPermCol.as<int>(); xf.scale_epi = M->rscale.p;
        return launch_trsv_level(ctx, M->lsU, M->dtype, M->factor_dtype, false, in, M->xsU.p, out, xf, skip);
      }
      if (which == BG_PC_START) {
        // y = diag(middle)^-1 U scatter_{permCol}(rscale .* x)   (preconditioner.cu:1561-1589)
        BG_TRY(launch_scatter_scale(ctx, M->dtype, n, in, M->rscale.p, M->permCol.as<int>(), M->tmp.p));
        BG_TRY(launch_csr_spmv_mixed(ctx, n, M->u_rp.as<int>(), M->u_ci.as<int>(), M->u_val.p, M->factor_dtype, M->tmp.p, M->tmp2.p, BG_F64, nullptr));
        return launch_divide_cast(ctx, M->dtype, n, M->tmp2.p, M->middle.p, out);
      }
      set_error("a split preconditioner has no single-sided application (use LEFT/RIGHT)");
      return BG_ERR_ARG;
    }
  }
  return BG_ERR_ARG;
}
}  // namespace bg

extern "C" int bg_precond_apply(bg_context_t ctx, bg_precond_t M, int which, const void* in, void* out) {
  if (!ctx || !M || !in || !out) return BG_ERR_ARG;
  return precond_apply(ctx, M, which, in, out, nullptr, false);
}
extern "C" int bg_precond_get_info(bg_precond_t M, bg_precond_info* o) {
  if (!M || !o) return BG_ERR_ARG;
  std::memset(o, 0, sizeof(*o));
  o->kind = M->kind; o->n = M->n; o->dtype = M->dtype; o->factor_dtype = M->factor_dtype; o->trsv_schedule = M->schedule;
  o->nnzL = M->nnzL; o->nnzU = M->nnzU;
  const int64_t s = dtype_size(M->dtype), sf = dtype_size(M->factor_dtype), n = M->n;
  if (M->kind == BG_PRECOND_LU || M->kind == BG_PRECOND_SPLIT) {
    if (M->schedule == BG_TRSV_BLOCK) {
      o->levelsL = M->fb.max_levels_L; o->levelsU = M->fb.max_levels_U; o->nblocks = M->fb.nblocks; o->max_block_rows = M->fb.max_rows;
      o->stream_bytes = M->fb.bytes;
      o->lattice = M->fb.lattice ? 1 : 0;
      o->device_factored = M->device_factored ? 1 : 0;
    } else {
      o->levelsL = M->lg.on ? M->lg_levels : M->lsL.n_levels; o->levelsU = M->lg.on ? M->lg_levels : M->lsU.n_levels;
      o->nblocks = 1; o->max_block_rows = M->n;
      o->stream_bytes = M->lg.on ? M->lg.bytes : M->lsL.bytes + M->lsU.bytes;
      o->lattice = M->lg.on ? 1 : 0;
    }
    // B_trsv = (s_f+4)(nnzL+nnzU) + 8(n+1) + 4 s n  (+ 3 s n + 8 n for the split vectors)
    o->trsv_bytes = (sf + 4) * (M->nnzL + M->nnzU) + 8 * (n + 1) + 4 * s * n + (M->kind == BG_PRECOND_SPLIT ? 3 * s * n + 8 * n : 0);
  } else if (M->kind == BG_PRECOND_DIAG) {
    o->trsv_bytes = 3 * s * n;
  } else if (M->kind == BG_PRECOND_AINV) {
    o->trsv_bytes = (s + 4) * (M->nnzL + M->nnzU) + (M->m2 ? 2 : 1) * (4 * (n + 1) + 2 * s * n) + (M->dinv.p ? 3 * s * n : 0);
  }
  return BG_OK;
}
extern "C" int bg_precond_get_levels(bg_precond_t M, int is_lower, int* level) {
  if (!M || !level) return BG_ERR_ARG;
  if (M->schedule != BG_TRSV_LEVEL || (M->kind != BG_PRECOND_LU && M->kind != BG_PRECOND_SPLIT)) return BG_ERR_UNSUPPORTED;
  const std::vector<int>& l = is_lower ? M->levelL : M->levelU;
  std::copy(l.begin(), l.end(), level);
  return BG_OK;
}
extern "C" int bg_precond_destroy(bg_precond_t M) { delete M; return BG_OK; }

// =========================================================================== solver
extern "C" void bg_gmres_opts_default(bg_gmres_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->m = 32; o->max_iter = 10000; o->tol = 1e-8; o->orth = BG_ORTH_MGS;  // the reference's order; CGS modes are opt-in
  o->reorth_eta = 0.70710678118654752; o->poll_chunk = 8; o->use_graph = 0; o->history_cap = 0;
}

static const int kHistCap = 1 << 16;

extern "C" int bg_solver_create(bg_context_t ctx, int n, int m, int dtype, bg_solver_t* out) {
  return solver_create_internal(ctx, n, m, dtype, 0, out);
}
extern "C" int bg_solver_create_ex(bg_context_t ctx, int n, int m, int dtype, int basis_dtype, bg_solver_t* out) {
  if (basis_dtype != dtype && !(dtype == BG_F64 && basis_dtype == BG_F32)) {
    set_error("bg_solver_create_ex: the basis is stored in the vector type, or in FP32 under FP64 vectors");
    return BG_ERR_UNSUPPORTED;
  }
  if (basis_dtype != dtype && m > 32) { set_error("bg_solver_create_ex: the mixed-precision basis runs the fused CGS step (m <= 32)"); return BG_ERR_UNSUPPORTED; }
  return solver_create_internal(ctx, n, m, dtype, 0, out, basis_dtype);
}
int bg::solver_create_internal(bg_context_t ctx, int n, int m, int dtype, int extra_rows, bg_solver_t* out, int basis_dtype) {
  if (!ctx || !out || n < 0 || m < 1 || m > 1024 || extra_rows < 0) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  bg_solver_t S = new bg_solver_s;
  S->ctx = ctx; S->n = n; S->m = m; S->dtype = dtype;
  S->basis_dtype = basis_dtype < 0 ? dtype : basis_dtype;
  // every vector has room for `extra_rows` halo values behind its n own entries (row-partitioned runs)
  S->ld = ((int64_t)n + extra_rows + 31) / 32 * 32;
  const size_t s = dtype_size(dtype);
  int rc = S->V.alloc((size_t)S->ld * (m + 1) * dtype_size(S->basis_dtype));
  if (rc >= 0 && S->basis_dtype != dtype) rc = S->vcur.alloc((size_t)S->ld * s);
  for (DevBuf* b : {&S->w, &S->ww, &S->r, &S->rr, &S->bb, &S->yv})
    if (rc >= 0) rc = b->alloc((size_t)S->ld * s);
  // state block: flags | scal | h | h2 | y | H | s | cs | sn | ticket | hist | partial
  const int pl = m + 2, nblk = ctx->sm_count * 8;  // dots grids never exceed 8 resident CTAs per SM
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
  const size_t o_flags = take(F_COUNT * sizeof(int)), o_scal = take(S_COUNT * 8), o_h = take(pl * 8), o_h2 = take(pl * 8),
               o_y = take((m + 1) * 8), o_H = take((size_t)(m + 1) * m * 8), o_s = take((m + 1) * 8), o_cs = take((m + 1) * 8),
               o_sn = take((m + 1) * 8), o_t = take(16), o_hist = take((size_t)kHistCap * 8), o_p = take((size_t)nblk * pl * 8);
  if (rc >= 0) rc = S->state.alloc(off);
  if (rc >= 0) rc = S->red.alloc(sizeof(double) * (size_t)(m + 8));
  if (rc >= 0) rc = S->mgs_slots.alloc(mgs_slot_bytes(ctx, m));
  if (rc >= 0 && cudaMemset(S->mgs_slots.p, 0, S->mgs_slots.bytes) != cudaSuccess) rc = BG_ERR_CUDA;
  if (rc < 0) { delete S; return rc; }
  cudaMemset(S->state.p, 0, off);
  unsigned char* b = S->state.as<unsigned char>();
  S->dev.flags = (int*)(b + o_flags); S->dev.scal = (double*)(b + o_scal); S->dev.h = (double*)(b + o_h);
  S->dev.h2 = (double*)(b + o_h2); S->dev.y = (double*)(b + o_y); S->dev.H = (double*)(b + o_H);
  S->dev.s = (double*)(b + o_s); S->dev.cs = (double*)(b + o_cs); S->dev.sn = (double*)(b + o_sn);
  S->dev.ticket = (unsigned*)(b + o_t); S->dev.hist = (double*)(b + o_hist); S->dev.partial = (double*)(b + o_p);
  S->dev.m = m; S->dev.hist_cap = 0; S->dev.partial_ld = pl;
  if (cudaMallocHost(&S->pinned, 4 * 128) != cudaSuccess) { cudaGetLastError(); delete S; set_error("pinned alloc failed"); return BG_ERR_ALLOC; }
  cudaEventCreateWithFlags(&S->ev[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&S->ev[1], cudaEventDisableTiming);
  cudaEventCreate(&S->t0);
  cudaEventCreate(&S->t1);
  *out = S;
  return BG_OK;
}
extern "C" int bg_solver_destroy(bg_solver_t S) {
  if (!S) return BG_OK;
  if (S->pinned) cudaFreeHost(S->pinned);
  for (cudaEvent_t e : {S->ev[0], S->ev[1], S->t0, S->t1}) if (e) cudaEventDestroy(e);
  delete S;
  return BG_OK;
}

namespace {
struct Status { int flags[F_COUNT]; double scal[S_COUNT]; };
static_assert(sizeof(Status) <= 128, "status slot");

// optional per-phase timing (bg_solver_profile): event pairs around kernel groups
struct Prof {
  std::vector<cudaEvent_t>* pool = nullptr;
  std::vector<std::pair<int, int>> spans[7];  // event index pairs per phase
  int64_t calls[7] = {0, 0, 0, 0, 0, 0, 0};
  size_t used = 0;
  cudaStream_t stream = nullptr;
  int mark() {
    if (used == pool->size()) { cudaEvent_t e; cudaEventCreate(&e); pool->push_back(e); }
    cudaEventRecord((*pool)[used], stream);
    return (int)used++;
  }
};
enum { PH_SPMV = 0, PH_TRSV = 1, PH_ORTH = 2, PH_OTHER = 3, PH_DOTS = 4, PH_UPDATE = 5, PH_HALO = 6 };

struct Driver {
  bg_context_t ctx; bg_solver_t S; bg_matrix_t A; bg_precond_t M; const bg_gmres_opts& o;
  const void* b; void* x;
  int n, m; int64_t ld; size_t es; bool split; const int* skip; GmresDev st;
  Prof* prof = nullptr;
  bool permuted = false;  // A, b, x and all work vectors are in the preconditioner's block-major row order
  DistHooks* dh = nullptr;  // row-partitioned run: halo exchange before every SpMV, allreduce inside every dots step
  bool fused_orth = true;   // BG_FUSED_ORTH=0 selects the separate dots / update kernels
  bool mixed = false;       // FP32 basis columns under FP64 work vectors: the newest basis vector also lives in S->vcur
  // y = alpha A x (+ beta b): x must have room for the halo behind its n entries when dh is set
  int spmv(double alpha, void* xv, double beta, const void* bv, void* y, bool axpby) {
    if (dh) {
      HaloArgs ha;
      if (spmv_supports_halo(A) && dh->halo_args(&ha)) return launch_spmv(ctx, A, alpha, xv, beta, bv, y, axpby, skip, &ha);
      Span sh(prof, PH_HALO);
      BG_TRY(dh->exchange(xv, skip));
    }
    return launch_spmv(ctx, A, alpha, xv, beta, bv, y, axpby, skip);
  }
  int dots(const void* Vp, int j, const void* w, int fin, int i, int mk, int nr) {
    if (!dh) return launch_dots(ctx, S->dtype, n, Vp, ld, j, w, st, fin, i, mk, o.orth, nr);
    // peer-memory transport: the all-reduce and the scalar epilogue run inside the dots kernel
    if (const PeerAR* par = dh->peer_ar()) return launch_dots(ctx, S->dtype, n, Vp, ld, j, w, st, fin, i, mk, o.orth, nr, nullptr, par);
    double* red = S->red.as<double>();
    BG_TRY(launch_dots(ctx, S->dtype, n, Vp, ld, j, w, st, fin, i, mk, o.orth, nr, red));
    BG_TRY(dh->allreduce(red, j + 1));
    return launch_finalize(ctx, st, fin, i, mk, o.orth, nr, j, red);
  }
  char* col(int k) const { return S->V.as<char>() + (size_t)k * ld * (mixed ? sizeof(float) : es); }
  // the vector the operator is applied to in inner iteration i: basis column i, or its working copy (mixed basis)
  void* vin(int i) const { return mixed ? S->vcur.p : (void*)col(i); }
  struct Span {
    Prof* p; int ph, a;
    Span(Prof* p_, int ph_) : p(p_), ph(ph_), a(p_ ? p_->mark() : 0) {}
    ~Span() { if (p) { p->spans[ph].emplace_back(a, p->mark()); p->calls[ph]++; } }
  };

  int apply_left(const void* in, void* out, const int* sk) { return precond_apply(ctx, M, split ? BG_PC_LEFT : BG_PC_FULL, in, out, sk, permuted); }

  // r = M_L^-1 (b - A x); beta, convergence test, v0 = r/beta
  int residual(int fin) {
    { Span sp(prof, PH_SPMV); BG_TRY(spmv(-1.0, x, 1.0, b, S->rr.p, true)); }
    { Span sp(prof, PH_TRSV); BG_TRY(apply_left(S->rr.p, S->r.p, skip)); }
    Span sp(prof, PH_OTHER);
    BG_TRY(dots(nullptr, 0, S->r.p, fin, 0, 0, 0));
    if (mixed) {  // v0 = r / beta, rounded to the basis precision in both copies
      BG_TRY(launch_update(ctx, S->dtype, n, S->r.p, nullptr, ld, 0, nullptr, 0.0, S->vcur.p, st, 1, 0, 0));
      return launch_round_store(ctx, n, S->vcur.as<double>(), (float*)col(0), skip);
    }
    return launch_update(ctx, S->dtype, n, S->r.p, nullptr, ld, 0, nullptr, 0.0, col(0), st, 1, 0, 0);
  }
  int iteration(int i) {
    // w = M_L^-1 A M_R^-1 v_i   (gmres.cu:2350-2354 / 2637-2640)
    if (split) {
      { Span sp(prof, PH_TRSV); BG_TRY(precond_apply(ctx, M, BG_PC_RIGHT, vin(i), S->w.p, skip, false)); }
      { Span sp(prof, PH_SPMV); BG_TRY(spmv(1.0, S->w.p, 0.0, nullptr, S->ww.p, false)); }
    } else {
      Span sp(prof, PH_SPMV);
      BG_TRY(spmv(1.0, vin(i), 0.0, nullptr, S->ww.p, false));
    }
    { Span sp(prof, PH_TRSV); BG_TRY(apply_left(S->ww.p, S->w.p, skip)); }
    Span sp(prof, PH_ORTH);
    if (o.orth == BG_ORTH_MGS) {
      if (fused_orth && !dh && mgs_fused_supported(ctx, n, m))  // one cooperative launch for the i+1 dot/axpy pairs and the norm
        return launch_mgs(ctx, S->dtype, n, col(0), ld, i, S->w.p, col(i + 1), st, S->mgs_slots.p, ++S->mgs_seq ? S->mgs_seq : ++S->mgs_seq);  // never 0
      for (int k = 0; k <= i; ++k) {
        BG_TRY(dots(col(k), 1, S->w.p, FIN_MGS_DOT, i, k, 0));
        BG_TRY(launch_update(ctx, S->dtype, n, S->w.p, col(k), ld, 1, st.h, -1.0, S->w.p, st, 0, 0, 0));
      }
      BG_TRY(dots(nullptr, 0, S->w.p, FIN_MGS_NORM, i, 0, 0));
      return launch_update(ctx, S->dtype, n, S->w.p, nullptr, ld, 0, nullptr, 0.0, col(i + 1), st, 1, 0, 0);
    }
    const int j = i + 1;
    if (mixed)  // FP32 basis columns, FP64 everything else; the new vector goes to the basis and to its working copy
      return launch_orth_mixed(ctx, n, (const float*)col(0), ld, j, S->w.as<double>(), S->vcur.as<double>(), (float*)col(j), st, i, o.orth);
    // one cooperative launch: projections, (peer all-reduce,) Givens epilogue, v_j, conditional second pass
    if (fused_orth && (!dh || dh->peer_ar()))
      return launch_orth(ctx, S->dtype, n, col(0), ld, j, S->w.p, col(j), st, i, o.orth, dh ? dh->peer_ar() : nullptr);
    { Span sd(prof, PH_DOTS); BG_TRY(dots(col(0), j, S->w.p, FIN_CGS1, i, 0, 0)); }
    { Span su(prof, PH_UPDATE); BG_TRY(launch_update(ctx, S->dtype, n, S->w.p, col(0), ld, j, st.h, -1.0, col(j), st, 1, 0, 0)); }
    BG_TRY(dots(col(0), j, col(j), FIN_CGS2, i, 0, 1));
    return launch_update(ctx, S->dtype, n, col(j), col(0), ld, j, st.h2, -1.0, col(j), st, 1, 0, 1);
  }
  // x += V[:,0..k] y  (split: y-vector += V y ; x = M_R^-1 y-vector)   gmres.cu:2399-2400
  int apply_update(int k, int force) {
    Span sp(prof, PH_OTHER);
    const int* sk = force ? nullptr : skip;
    BG_TRY(launch_solve_y(ctx, st, k, force));
    void* tgt = split ? S->yv.p : x;
    if (mixed) BG_TRY(launch_axpy_basis(ctx, n, (const float*)col(0), ld, k + 1, st.y, (double*)tgt, st, force));
    else BG_TRY(launch_update(ctx, S->dtype, n, tgt, col(0), ld, k + 1, st.y, 1.0, tgt, st, 0, force, 0));
    if (split) BG_TRY(precond_apply(ctx, M, BG_PC_RIGHT, S->yv.p, x, sk, false));
    return BG_OK;
  }
  int poll_async(int slot) {
    BG_CUDA(cudaMemcpyAsync((char*)S->pinned + slot * 128, st.flags, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA(cudaMemcpyAsync((char*)S->pinned + slot * 128 + sizeof(int) * F_COUNT, st.scal, 8 * S_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA(cudaEventRecord(S->ev[slot], ctx->stream));
    return BG_OK;
  }
  int poll_wait(int slot, Status* out) {
    BG_CUDA(cudaEventSynchronize(S->ev[slot]));
    std::memcpy(out, (char*)S->pinned + slot * 128, sizeof(Status));
    return BG_OK;
  }
};
}  // namespace

static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// D^-1 P A P^T for the block-major order of a block-Jacobi preconditioner (new row i = old row
// perm[i], D = diag(U) in that order); built once per ordering and cached inside A.  With b scaled
// the same way the left-preconditioned operator M^-1 A and M^-1 b are unchanged, but the device
// triangular solves become unit solves (trsv_block.cu).
template <class T> static int build_permuted_host(bg_matrix_t A, const FusedBlockDev& fb) {
  const int n = A->n;
  const std::vector<int>& perm = fb.h_perm;
  std::vector<int> inv(n);
  for (int i = 0; i < n; ++i) inv[perm[i]] = i;
  const T* v = reinterpret_cast<const T*>(A->h_val.data());
  std::vector<int> rp(n + 1, 0), ci(A->h_ci.size());
  std::vector<T> vv(A->h_ci.size());
  for (int i = 0; i < n; ++i) rp[i + 1] = rp[i] + (A->h_rp[perm[i] + 1] - A->h_rp[perm[i]]);
  parallel_ranges((size_t)n, [&](size_t r0, size_t r1) {  // every permuted row writes its own segment
    std::vector<std::pair<int, T>> row;
    for (int i = (int)r0; i < (int)r1; ++i) {
      const int o = perm[i];
      row.clear();
      const double di = fb.h_dinv[i];  // row scaling D^-1 (D = diag U): the device solves are unit solves
      for (int p = A->h_rp[o]; p < A->h_rp[o + 1]; ++p) {
        const int c = A->h_ci[p];  // halo columns (>= n, row-partitioned matrices) keep their slot
        row.emplace_back(c < n ? inv[c] : c, (T)((double)v[p] * di));
      }
      std::sort(row.begin(), row.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
      int q = rp[i];
      for (auto& e : row) { ci[q] = e.first; vv[q++] = e.second; }
    }
  });
  bg_matrix_t P = new bg_matrix_s;
  P->ctx = A->ctx; P->n = n; P->ncols = A->ncols; P->nnz = A->nnz; P->dtype = A->dtype;
  int rc = matrix_build<T>(P, std::move(rp), std::move(ci), vv.data(), sizeof(T) == 4 ? BG_F32 : BG_F64, A->requested_format);
  if (rc < 0) { delete P; return rc; }
  P->h_rp.clear(); P->h_rp.shrink_to_fit(); P->h_ci.clear(); P->h_ci.shrink_to_fit(); P->h_val.clear(); P->h_val.shrink_to_fit();
  delete A->permuted;
  A->permuted = P;
  A->permuted_serial = fb.serial;
  return BG_OK;
}
// the same matrix built on the device from A's CSR arrays in HBM (permute_rows_kernel): only the new row pointers are
// computed on the host.  Square matrices only (row-partitioned ones need the per-slice halo flags of the host path).
template <class T> static int build_permuted_device(bg_matrix_t A, const FusedBlockDev& fb) {
  const int n = A->n;
  cudaStream_t s = A->ctx->stream;
  const std::vector<int>& perm = fb.h_perm;
  std::vector<int> rp((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) rp[(size_t)i + 1] = rp[i] + (A->h_rp[(size_t)perm[i] + 1] - A->h_rp[perm[i]]);
  bg_matrix_t P = new bg_matrix_s;
  P->ctx = A->ctx; P->n = n; P->ncols = A->ncols; P->nnz = A->nnz; P->dtype = A->dtype;
  P->requested_format = A->requested_format;
  int rc = BG_OK;
  do {
    DevBuf dinv;
    if ((rc = dinv.upload(fb.h_dinv.data(), sizeof(double) * (size_t)n, s)) < 0) break;
    if ((rc = upload_vec(P->rp, rp, s)) < 0) break;
    if ((rc = P->ci.alloc(sizeof(int) * (size_t)std::max<int64_t>(A->nnz, 1))) < 0) break;
    if ((rc = P->val.alloc(sizeof(T) * (size_t)std::max<int64_t>(A->nnz, 1))) < 0) break;
    if ((rc = launch_permute_rows(A->ctx, A->dtype, n, A->rp.as<int>(), A->ci.as<int>(), A->val.p, fb.perm.as<int>(), fb.iperm.as<int>(),
                                  dinv.as<double>(), P->rp.as<int>(), P->ci.as<int>(), P->val.p)) < 0) break;
    if ((rc = matrix_finish<T>(P, rp.data(), A->requested_format, nullptr, nullptr)) < 0) break;
    cudaError_t e = cudaStreamSynchronize(s);  // dinv is released at the end of this scope
    if (e != cudaSuccess) { rc = cuda_fail(e, "build_permuted_device", __FILE__, __LINE__); break; }
  } while (false);
  if (rc < 0) { delete P; return rc; }
  delete A->permuted;
  A->permuted = P;
  A->permuted_serial = fb.serial;
  return BG_OK;
}
template <class T> static int build_permuted(bg_matrix_t A, const FusedBlockDev& fb) {
  if (host_setup() || A->ncols != A->n || !fb.perm.p || !fb.iperm.p || (int)fb.h_dinv.size() != A->n) return build_permuted_host<T>(A, fb);
  return build_permuted_device<T>(A, fb);
}
static int permuted_matrix(bg_matrix_t A, const FusedBlockDev& fb, bg_matrix_t* out) {
  if (!A->permuted || A->permuted_serial != fb.serial) {
    if ((int)fb.h_perm.size() != A->n) { set_error("preconditioner ordering does not match the matrix"); return BG_ERR_ARG; }
    BG_TRY(A->dtype == BG_F32 ? build_permuted<float>(A, fb) : build_permuted<double>(A, fb));
  }
  *out = A->permuted;
  return BG_OK;
}

namespace bg {
// one-time work of the first solve with a block-Jacobi preconditioner (the block-major copy of A), done ahead of it
int prepare_solve(bg_matrix_t A, bg_precond_t M) {
  if (!A || !M) return BG_ERR_ARG;
  if (M->kind != BG_PRECOND_LU || M->schedule != BG_TRSV_BLOCK || M->n != A->n) return BG_OK;
  BG_CUDA(cudaSetDevice(A->ctx->device));
  bg_matrix_t P;
  return permuted_matrix(A, M->fb, &P);
}
}  // namespace bg

static int gmres_solve_impl(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                            const bg_gmres_opts* opts, bg_gmres_result* res, Prof* prof, DistHooks* dh = nullptr) {
  if (!ctx || !S || !A || !M || !d_b || !d_x || !opts) { set_error("bg_gmres_solve: null argument"); return BG_ERR_ARG; }
  const bg_gmres_opts& o = *opts;
  if (A->n != S->n || M->n != S->n || A->dtype != S->dtype || M->dtype != S->dtype) { set_error("matrix / preconditioner / solver shapes or dtypes differ"); return BG_ERR_ARG; }
  if (o.m < 1 || o.m > S->m) { set_error("restart m=%d exceeds the solver workspace (m=%d)", o.m, S->m); return BG_ERR_ARG; }
  if (o.orth != BG_ORTH_MGS && o.m > 32) { set_error("fused classical Gram-Schmidt supports m <= 32 (use BG_ORTH_MGS beyond)"); return BG_ERR_UNSUPPORTED; }
  if (!aligned16(d_b) || !aligned16(d_x)) { set_error("device vectors must be 16-byte aligned"); return BG_ERR_ARG; }
  if (o.max_iter < 0 || !(o.tol >= 0)) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  const int64_t launches0 = ctx->launches;
  const int n = S->n, m = o.m;
  // block-Jacobi preconditioners solve in their own block-major row order: run the whole iteration
  // on P A P^T, P b, P x (residual norms and the Hessenberg matrix are permutation invariant)
  const bool permuted = M->kind == BG_PRECOND_LU && M->schedule == BG_TRSV_BLOCK;
  void* const d_x_user = d_x;
  if (permuted) {
    BG_TRY(permuted_matrix(A, M->fb, &A));
    const size_t bytes = (size_t)S->ld * dtype_size(S->dtype);
    if (S->bp.bytes < bytes) { BG_TRY(S->bp.alloc(bytes)); BG_TRY(S->xp.alloc(bytes)); }
  }
  GmresDev st = S->dev;
  st.m = m;  // H leading dimension follows the active restart length
  S->last_m = m;
  st.hist_cap = std::min(std::max(o.history_cap, 0), kHistCap);
  const int chunk = std::max(1, o.poll_chunk);

  BG_CUDA(cudaEventRecord(S->t0, ctx->stream));
  if (permuted) {
    BG_TRY(launch_gather(ctx, S->dtype, n, M->fb.perm.as<int>(), d_b, S->bp.p, M->fb.dinv.p, M->factor_dtype));
    BG_TRY(launch_gather(ctx, S->dtype, n, M->fb.perm.as<int>(), d_x, S->xp.p));
    d_b = S->bp.p;
    d_x = S->xp.p;
  }
  if (dh && !permuted) { set_error("a row-partitioned solve needs a block-Jacobi (BG_TRSV_BLOCK) ILU preconditioner"); return BG_ERR_UNSUPPORTED; }
  Driver D{ctx, S, A, M, o, d_b, d_x, n, m, S->ld, dtype_size(S->dtype), M->kind == BG_PRECOND_SPLIT, st.flags + F_CONV, st, prof, permuted, dh};
  { const char* e = getenv("BG_FUSED_ORTH"); D.fused_orth = !(e && e[0] == '0'); }
  D.mixed = S->basis_dtype != S->dtype;
  if (D.mixed && (o.orth == BG_ORTH_MGS || dh || o.m > 32)) {
    set_error("the mixed-precision basis supports the fused CGS step on one GPU (BG_ORTH_CGS_ADAPTIVE / BG_ORTH_CGS2, m <= 32)");
    return BG_ERR_UNSUPPORTED;
  }
  BG_TRY(launch_state_reset(ctx, st, o.tol, o.reorth_eta));
  // normb = || M_L^-1 b ||   (gmres.cu:2294-2296)
  BG_TRY(D.apply_left(d_b, S->bb.p, nullptr));
  BG_TRY(D.dots(nullptr, 0, S->bb.p, FIN_NORMB, 0, 0, 0));
  if (D.split) BG_TRY(precond_apply(ctx, M, BG_PC_START, d_x, S->yv.p, nullptr, false));  // gmres.cu:2300
  BG_TRY(D.residual(FIN_BETA0));

  Status stt;
  std::memset(&stt, 0, sizeof(stt));
  int total = 0, pend = -1, cur = 0;
  bool done = false;
  auto poll_point = [&]() -> int {
    BG_TRY(D.poll_async(cur));
    if (pend >= 0) {
      BG_TRY(D.poll_wait(pend, &stt));
      if (stt.flags[F_CONV]) done = true;
    }
    pend = cur;
    cur ^= 1;
    return BG_OK;
  };
  BG_TRY(poll_point());
  while (!done && total < o.max_iter) {
    const int kc = std::min(m, o.max_iter - total);
    for (int i = 0; i < kc && !done; ++i) {
      BG_TRY(D.iteration(i));
      ++total;
      if (total % chunk == 0) BG_TRY(poll_point());
    }
    if (done) break;
    // end of cycle: x update, true residual, restart vector (gmres.cu:2399-2428)
    BG_TRY(D.apply_update(kc - 1, 0));
    BG_TRY(D.residual(FIN_BETA));
    BG_TRY(poll_point());
  }
  if (pend >= 0) {
    BG_TRY(D.poll_wait(pend, &stt));
    pend = -1;
  }
  if (stt.flags[F_CONV] == 1) {
    // in-cycle convergence at inner index k_conv: the update was skipped by the flag, do it now
    BG_TRY(D.apply_update(stt.flags[F_KCONV], 1));
  }
  if (permuted) BG_TRY(launch_scatter(ctx, S->dtype, n, M->fb.perm.as<int>(), d_x, d_x_user));
  BG_CUDA(cudaEventRecord(S->t1, ctx->stream));
  BG_CUDA(cudaStreamSynchronize(ctx->stream));
  S->last_hist_len = std::min(stt.flags[F_ITERS], st.hist_cap);
  if (S->last_hist_len > 0) {
    S->last_hist.resize(S->last_hist_len);
    BG_CUDA(cudaMemcpy(S->last_hist.data(), st.hist, 8 * (size_t)S->last_hist_len, cudaMemcpyDeviceToHost));
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, S->t0, S->t1);
  const int conv = stt.flags[F_CONV];
  if (res) {
    res->flag = conv ? 0 : 1;
    res->iters_done = stt.flags[F_ITERS];
    // j was already incremented when convergence is detected at a restart boundary
    res->iters = conv == 1 ? stt.flags[F_ITERS] : conv == 2 ? (stt.flags[F_ITERS] == 0 ? 0 : stt.flags[F_ITERS] + 1) : o.max_iter;
    res->resid = stt.scal[S_RESID];
    res->restarts = stt.flags[F_RESTARTS];
    res->reorths = stt.flags[F_NREORTH];
    res->launches = ctx->launches - launches0;
    res->solve_ms = ms;
  }
  return conv ? BG_OK : BG_NOT_CONVERGED;
}

int bg::gmres_solve_dist(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                         const bg_gmres_opts* opts, bg_gmres_result* res, DistHooks* dh) {
  return gmres_solve_impl(ctx, S, A, M, d_b, d_x, opts, res, nullptr, dh);
}

extern "C" int bg_gmres_solve(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                              const bg_gmres_opts* opts, bg_gmres_result* res) {
  return gmres_solve_impl(ctx, S, A, M, d_b, d_x, opts, res, nullptr);
}

extern "C" int bg_gmres_solve_host(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* h_b,
                                   void* h_x, const bg_gmres_opts* o, bg_gmres_result* r) {
  if (!ctx || !S || !h_b || !h_x) return BG_ERR_ARG;
  // gmresInterfacePGfloat::GMRES_dev_PG (gmres_interface_pg.cu:118-129): H2D x and rhs, solve, D2H x.
  const size_t bytes = (size_t)S->n * dtype_size(S->dtype);
  DevBuf &xb = S->hx, &bb = S->hb;
  if (xb.bytes < bytes) { BG_TRY(xb.alloc(bytes + 16)); BG_TRY(bb.alloc(bytes + 16)); }
  BG_CUDA(cudaMemcpyAsync(xb.p, h_x, bytes, cudaMemcpyHostToDevice, ctx->stream));
  BG_CUDA(cudaMemcpyAsync(bb.p, h_b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = bg_gmres_solve(ctx, S, A, M, bb.p, xb.p, o, r);
  if (rc < 0) return rc;
  BG_CUDA(cudaMemcpyAsync(h_x, xb.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BG_CUDA(cudaStreamSynchronize(ctx->stream));
  return rc;
}

extern "C" int bg_solver_get_history(bg_solver_t S, double* hist, int cap, int* len) {
  if (!S) return BG_ERR_ARG;
  const int k = std::min(cap, S->last_hist_len);
  if (hist && k > 0) std::copy(S->last_hist.begin(), S->last_hist.begin() + k, hist);
  if (len) *len = S->last_hist_len;
  return BG_OK;
}
extern "C" int bg_solver_get_hessenberg(bg_solver_t S, double* H, double* s, double* cs, double* sn) {
  if (!S) return BG_ERR_ARG;
  const int m = S->last_m > 0 ? S->last_m : S->m;  // the layout of the last solve (its restart length may be below the workspace's)
  BG_CUDA(cudaStreamSynchronize(S->ctx->stream));
  if (H) BG_CUDA(cudaMemcpy(H, S->dev.H, 8 * (size_t)(m + 1) * m, cudaMemcpyDeviceToHost));
  if (s) BG_CUDA(cudaMemcpy(s, S->dev.s, 8 * (size_t)(m + 1), cudaMemcpyDeviceToHost));
  if (cs) BG_CUDA(cudaMemcpy(cs, S->dev.cs, 8 * (size_t)(m + 1), cudaMemcpyDeviceToHost));
  if (sn) BG_CUDA(cudaMemcpy(sn, S->dev.sn, 8 * (size_t)(m + 1), cudaMemcpyDeviceToHost));
  return BG_OK;
}

// one orthogonalisation step against j basis columns on the solver's own workspace (bench hook)
extern "C" BG_API int bg_solver_bench_orth(bg_context_t ctx, bg_solver_t S, int j) {
  if (!ctx || !S || j < 1 || j > std::min(S->m, 32)) return BG_ERR_ARG;
  if (S->basis_dtype != S->dtype) { set_error("bg_solver_bench_orth: not for the mixed-precision basis"); return BG_ERR_UNSUPPORTED; }
  GmresDev st = S->dev;
  BG_TRY(launch_dots(ctx, S->dtype, S->n, S->V.p, S->ld, j, S->w.p, st, FIN_NONE, 0, 0, BG_ORTH_CGS_ADAPTIVE, 0));
  return launch_update(ctx, S->dtype, S->n, S->w.p, S->V.p, S->ld, j, st.h, -1.0, S->V.as<char>() + (size_t)j * S->ld * dtype_size(S->dtype), st, 1, 0, 0);
}

extern "C" int bg_solver_mgs_fused(bg_context_t ctx, bg_solver_t S) {
  if (!ctx || !S) return BG_ERR_ARG;
  const char* e = getenv("BG_FUSED_ORTH");  // read per call, like the solve does
  return !(e && e[0] == '0') && mgs_fused_supported(ctx, S->n, S->m) ? 1 : 0;
}

extern "C" int bg_solver_profile(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                                 const bg_gmres_opts* o, bg_gmres_result* r, bg_phase_stats* out) {
  return gmres_profile_dist(ctx, S, A, M, d_b, d_x, o, r, out, nullptr);
}
int bg::gmres_profile_dist(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                           const bg_gmres_opts* o, bg_gmres_result* r, bg_phase_stats* out, DistHooks* dh) {
  if (!ctx || !S || !A || !M || !out) return BG_ERR_ARG;
  Prof prof;
  prof.pool = &S->prof_events;
  prof.stream = ctx->stream;
  bg_gmres_result rr;
  std::memset(&rr, 0, sizeof(rr));
  int rc = gmres_solve_impl(ctx, S, A, M, d_b, d_x, o, &rr, &prof, dh);
  if (rc < 0) return rc;
  if (r) *r = rr;
  double ms[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int ph = 0; ph < 7; ++ph)
    for (auto& sp : prof.spans[ph]) {
      float t = 0;
      cudaEventElapsedTime(&t, S->prof_events[sp.first], S->prof_events[sp.second]);
      ms[ph] += t;
    }
  std::memset(out, 0, sizeof(*out));
  out->spmv_ms = ms[PH_SPMV]; out->trsv_ms = ms[PH_TRSV]; out->orth_ms = ms[PH_ORTH]; out->other_ms = ms[PH_OTHER];
  out->spmv_calls = prof.calls[PH_SPMV]; out->trsv_calls = prof.calls[PH_TRSV]; out->orth_calls = prof.calls[PH_ORTH];
  out->dots_ms = ms[PH_DOTS]; out->update_ms = ms[PH_UPDATE];
  out->dots_calls = prof.calls[PH_DOTS]; out->update_calls = prof.calls[PH_UPDATE];
  out->halo_ms = ms[PH_HALO]; out->halo_calls = prof.calls[PH_HALO];
  // algorithmic bytes of the work actually performed (iters_done inner iterations)
  bg_matrix_info mi; bg_precond_info pi;
  bg_matrix_get_info(A, &mi); bg_precond_get_info(M, &pi);
  const int64_t s = dtype_size(S->dtype), n = S->n, it = rr.iters_done, m = o->m;
  const bool split = M->kind == BG_PRECOND_SPLIT;
  const int64_t nres = 1 + rr.restarts + (rr.flag == 0 && rr.iters != rr.iters_done ? 1 : 0);  // residual evaluations
  out->spmv_bytes = (it + nres) * mi.spmv_bytes;
  // a left-only application runs both factors; the split one runs L and U in separate calls
  out->trsv_bytes = (it + nres + 1) * pi.trsv_bytes + (split ? 0 : 0);
  int64_t orth = 0;
  for (int64_t t = 0; t < it; ++t) { const int64_t j = t % m + 1; orth += 2 * j * s * n + 5 * s * n; }  // SURVEY 8d B_orth(j)
  orth += (int64_t)rr.reorths * 0;  // second passes are reported through `reorths`, not counted as algorithmic bytes
  out->orth_bytes = orth;
  // first-pass kernels alone, with the bytes each launch cannot avoid: dots reads j basis columns + w,
  // update reads them again + w and writes v_j (CGS modes only; zero under strict MGS)
  if (o->orth != BG_ORTH_MGS)
    for (int64_t t = 0; t < it; ++t) { const int64_t j = t % m + 1; out->dots_bytes += (j + 1) * s * n; out->update_bytes += (j + 2) * s * n; }
  out->other_bytes = nres * (3 * s * n) + (rr.restarts + 1) * (s * n * m + 2 * s * n);
  return rc;
}
