// internal.h -- shared declarations of libb200gmres (not part of the public C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200gmres.h"
#include "../host/host_algos.h"

namespace bg {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BG_CUDA(call)                                                    \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return bg::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
#define BG_TRY(call)              \
  do {                            \
    int rc__ = (call);            \
    if (rc__ < 0) return rc__;    \
  } while (0)

inline size_t dtype_size(int dt) { return dt == BG_F64 ? 8 : 4; }

// Programmatic dependent launch: the kernel may start (launch latency, CTA prologue, loads of constant data)
// while its predecessor on the stream drains; it must execute pdl_wait() before touching anything the
// predecessor wrote.  BG_PDL=0 falls back to plain stream launches.
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// simple owning device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t b);
  int upload(const void* h, size_t b, cudaStream_t s);
  void release();
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace bg

struct bg_context_s {
  int device = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  size_t smem_optin = 0;
};

struct bg_matrix_s {
  bg_context_t ctx = nullptr;
  int n = 0, dtype = BG_F64, format = BG_SPMV_CSR_VECTOR, lanes = 4;
  int ncols = 0;  // > n for the local rows of a row-partitioned matrix (own columns, then halo slots)
  int64_t nnz = 0;
  bg::SpmvPlan plan;
  bg::DevBuf rp, ci, val;            // CSR
  bg::DevBuf s_ptr, s_col, s_val;    // sliced ELL
  bg::DevBuf s_halo;                 // per slice: 1 when one of its columns is a halo slot (ncols > n only)
  int sell_rpt = 1, sell_height = 32, sell_slices = 0;
  int sell_width = 0;  // > 0: every slice has this many entries per row (stencils): kernels skip the slice pointers
  // host CSR (compute dtype) kept so that a symmetrically permuted copy P A P^T can be built when a
  // preconditioner asks for its own row order; one cached copy per ordering serial
  std::vector<int> h_rp, h_ci;
  std::vector<unsigned char> h_val;
  int requested_format = BG_SPMV_AUTO;
  bg_matrix_s* permuted = nullptr;
  uint64_t permuted_serial = 0;
  ~bg_matrix_s() { delete permuted; }
};

// ---- device-side argument bundles --------------------------------------------------------
namespace bg {

struct TrsvXform {      // prologue / epilogue of the split preconditioner
  int pro = 0;          // 0 plain, 1 left-split (in[perm]/lscale[perm]), 2 right-split (in*middle)
  int epi = 0;          // 0 plain, 1 right-split scatter (out[ipc[r]] = x/rscale[ipc[r]])
  const int* perm = nullptr;   // permRow (pro 1) / inverse permCol (epi 1)
  const void* scale_pro = nullptr;  // lscale or middle   (vector dtype)
  const void* scale_epi = nullptr;  // rscale
};

// device copy of host_algos.h SyncFreeStream (BG_TRSV_LEVEL, any factor; trsv.cu trsv_syncfree)
struct SyncFreeDev {
  int n = 0, n_levels = 0, n_slices = 0, max_K = 0, max_level_rows = 0;
  DevBuf meta, rows, cols, vals, diag;
  DevBuf src_idx, dst_idx;           // per slot: where the right-hand side is read / the result written when the split
                                     // preconditioner's permutations are folded into the solve (empty: the row itself)
  std::vector<int> h_rows;           // host copy of rows (to build src_idx / dst_idx)
  int64_t bytes = 0;
  unsigned seq = 0;        // sequence number of the flag-in-data packets, one per launch
};
// device copy of host_algos.h LatticeGlobal (BG_TRSV_LEVEL, 5-point ILU(0) in natural order; trsv_lattice_global)
struct LatticeGlobalDev {
  bool on = false;
  int nx = 0, ny = 0, strips = 0;
  DevBuf lower, upper, udinv, bnd;  // bnd: [2][strips][ny] boundary packets (L half, U half)
  int64_t bytes = 0;
  unsigned seq = 0;
};

// device copy of host_algos.h FusedBlockStream (block-Jacobi L+U solve, trsv_block.cu)
struct FusedBlockDev {
  int n = 0, nblocks = 0, max_rows = 0, max_levels_L = 0, max_levels_U = 0, max_chunk_bytes = 0, max_chunks = 0;
  int stages = 4, warps_per_cta = 4, tab_stride = 0;
  bool lattice = false;  // index-free layout of 5-point grid tiles (host_algos.h FusedBlockStream::lattice)
  int64_t bytes = 0, total_steps = 0;
  DevBuf perm, iperm, dinv, blk, chunk_tab, stream;
  std::vector<int> h_perm;     // new -> old, kept for permuting matrices on the host
  std::vector<double> h_dinv;  // block-major 1/diag(U): the solver folds it into its copy of A and into b
  uint64_t serial = 0;      // identifies this ordering (matrix copies are cached per serial)
};

// state shared by the orthogonalisation / Givens kernels (all double, device memory)
struct GmresDev {
  int* flags;      // [0] converged (1 in-cycle, 2 at residual check) [1] k_conv [2] iters
                   // [3] reorth_flag [4] reorth_count [5] breakdown [6] restarts
  double* scal;    // [0] normb [1] beta [2] resid [3] scale [4] tol [5] eta^2 [6] ww
  double *h, *h2, *y, *H, *s, *cs, *sn, *partial, *hist;
  unsigned* ticket;
  int m, hist_cap, partial_ld;
};
enum { F_CONV = 0, F_KCONV = 1, F_ITERS = 2, F_REORTH = 3, F_NREORTH = 4, F_BREAK = 5, F_RESTARTS = 6, F_COUNT = 8 };
enum { S_NORMB = 0, S_BETA = 1, S_RESID = 2, S_SCALE = 3, S_TOL = 4, S_ETA2 = 5, S_WW = 6, S_COUNT = 8 };
enum { FIN_NONE = 0, FIN_NORMB, FIN_BETA0, FIN_BETA, FIN_CGS1, FIN_CGS2, FIN_MGS_DOT, FIN_MGS_NORM };

// ---- peer-memory transport of a row-partitioned solve (dist.cu; kernels in orth.cu) ----------
// Every rank owns a small "window" allocation that all other ranks of the box map through CUDA IPC
// (NVLink / NVSwitch peer stores).  Messages use a low-latency flag-in-data format: a double travels
// as ONE 16-byte store {lo32, seq, hi32, seq}; the receiver polls the element itself until both
// sequence words match, so no memory fence, no separate flag and no second round trip are needed.
// Both structs live in DEVICE memory of the owning rank.
constexpr int kMaxPeers = 16;   // ranks of one box
constexpr int kArLd = 40;       // elements per rank and slot of the all-reduce inbox (>= kFinMax + 2)
struct __align__(16) LL16 { unsigned lo, f0, hi, f1; };
// All-reduce fused into the last CTA of the dots / orth kernels: every rank stores its j+1 partial sums
// into the inbox of every rank (slot = parity of a device-side sequence number, row = sender), polls its
// own inbox and sums the rows in rank order -- the same bits on every rank, so stopping decisions stay
// identical without any further exchange.
struct PeerAR {
  int nranks, rank;
  unsigned long long timeout_ns; // a poll that lasts longer traps (dead peer)
  unsigned* seq;                 // local counter of executed all-reduces
  LL16* inbox[kMaxPeers];        // inbox[r]: rank r's [2][nranks][kArLd] elements (peer-mapped for r != rank)
};
// Halo exchange: one kernel stores the boundary entries of the vector entering the SpMV into the
// neighbours' halo inboxes and copies its own arrived inbox into the halo slots behind the local vector.
struct PeerHalo {
  int n_nbr, n_halo;
  unsigned long long timeout_ns;
  int send_ptr[kMaxPeers + 1];   // entries [send_ptr[k], send_ptr[k+1]) of send_idx go to neighbour k
  LL16* peer_inbox[kMaxPeers];   // first element of MY range in neighbour k's halo inbox (buffer 0)
  int peer_half[kMaxPeers];      // elements per buffer of neighbour k's inbox (= its n_halo)
  LL16* my_inbox;                // [2][n_halo]: double-buffered by the parity of the sequence number, so that a fast
                                 // neighbour's exchange s+1 cannot overwrite values of exchange s that are still unread
};
// arguments of a SpMV that carries its own halo exchange (spmv.cu, HALO variants): the first wave of CTAs
// stores this rank's boundary entries of x into the neighbours' inboxes, rows with halo columns read their
// operands straight from the own inbox (polling the sequence number carried by every element)
struct HaloArgs {
  const PeerHalo* ph = nullptr;
  const int* send_idx = nullptr;
  int n_send = 0, n_local = 0;
  unsigned seq = 0;
  const unsigned char* slice_halo = nullptr;  // per slice of the matrix: reads a halo slot (filled in by launch_spmv)
};

// ---- launchers (defined in the .cu files) --------------------------------------------------
int launch_spmv(bg_context_t ctx, bg_matrix_t A, double alpha, const void* x, double beta, const void* b,
                void* y, bool axpby, const int* skip_flag, const HaloArgs* halo = nullptr);
// true when launch_spmv can carry the halo exchange itself for this matrix (uniform-width sliced ELL)
bool spmv_supports_halo(bg_matrix_t A);
// generic CSR y = A x with independent value/vector types (used for the warm-start U*x)
int launch_csr_spmv_mixed(bg_context_t ctx, int n, const int* rp, const int* ci, const void* val, int val_dt,
                          const void* x, void* y, int vec_dt, const int* skip_flag);

// dots: partial[k] = V[:,k].w for k<j, partial[j] = w.w ; then finalize variant `fin`
// red_out != null (row-partitioned run): write this rank's sums there and skip the finalize; the caller
// allreduces them and runs launch_finalize
int launch_dots(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int j, const void* w,
                const GmresDev& st, int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth_flag,
                double* red_out = nullptr, const PeerAR* peer = nullptr);
// one halo exchange: gather + peer-store of the boundary entries of v, then wait for / unpack the own halo
// into halo_out (= v + n_local); seq identifies the exchange (same value on both sides)
int launch_halo_exchange(bg_context_t ctx, int dtype, const PeerHalo* d_ph, int n_send, int n_halo, const int* idx,
                         const void* v, void* halo_out, unsigned seq, const int* skip);
// fused CGS step (adaptive or CGS2): projections + (peer all-reduce) + Givens epilogue + v_j, one cooperative launch
int launch_orth(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int j, const void* w, void* out,
                const GmresDev& st, int inner_i, int orth_mode, const PeerAR* peer);
// strict MGS step of inner iteration i as one cooperative launch (single GPU, vectors that fit the register-resident
// layout: mgs_fused_supported); slots = mgs_slot_bytes() of zero-initialised device memory, seq = launch counter >= 1
bool mgs_fused_supported(bg_context_t ctx, int n, int m);
size_t mgs_slot_bytes(bg_context_t ctx, int m);
int launch_mgs(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int i, const void* w, void* out, const GmresDev& st,
               void* slots, unsigned seq);
// mixed-precision basis (FP64 work vectors, FP32 basis columns; orth.cu)
int launch_orth_mixed(bg_context_t ctx, int n, const float* V32, int64_t ld, int j, const double* w, double* out, float* out_b,
                      const GmresDev& st, int inner_i, int orth_mode);
int launch_round_store(bg_context_t ctx, int n, double* v, float* v32, const int* skip);
int launch_axpy_basis(bg_context_t ctx, int n, const float* V32, int64_t ld, int k, const double* y, double* x, const GmresDev& st, int force);
int launch_finalize(bg_context_t ctx, const GmresDev& st, int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth,
                    int j, const double* red_in);
int launch_pack(bg_context_t ctx, int dtype, int cnt, const int* idx, const void* v, void* buf, const int* skip);
// out = (a + sgn * sum_k coef[k] V[:,k]) * (use_scale ? scal[S_SCALE] : 1)
int launch_update(bg_context_t ctx, int dtype, int n, const void* a, const void* V, int64_t ld, int j,
                  const double* coef, double sgn, void* out, const GmresDev& st, int use_scale, int force,
                  int need_reorth_flag);
int launch_solve_y(bg_context_t ctx, const GmresDev& st, int k, int force);
int launch_state_reset(bg_context_t ctx, const GmresDev& st, double tol, double eta);

// one triangular solve in the reference's level order; xs = n flag-in-data packets (16 bytes each) of scratch
int launch_trsv_level(bg_context_t ctx, SyncFreeDev& S, int vec_dt, int fac_dt, bool unit,
                      const void* in, void* xs, void* out, const TrsvXform& xf, const int* skip_flag);
// z = U^-1 L^-1 r with the global factors of a 5-point grid (two launches); tmp = n values of the compute type
int launch_trsv_lattice_global(bg_context_t ctx, LatticeGlobalDev& S, int vec_dt, int fac_dt, const void* in, void* tmp,
                               void* out, const int* skip_flag);
int lattice_global_capacity(bg_context_t ctx, int vec_dt, int fac_dt);
// z = U^-1 L^-1 r per diagonal block.  gather = vectors are in the ORIGINAL row order (the kernel
// goes through perm); otherwise they are in the block-major order of the stream.
int launch_trsv_block(bg_context_t ctx, const FusedBlockDev& S, int vec_dt, int fac_dt, bool gather,
                      const void* in, void* out, const int* skip_flag);
// out[i] = in[perm[i]] (* scale[i] when scale != null, scale in the factor dtype)  /  out[perm[i]] = in[i]
int launch_gather(bg_context_t ctx, int dtype, int n, const int* perm, const void* in, void* out, const void* scale = nullptr,
                  int scale_dt = BG_F64);
int launch_scatter(bg_context_t ctx, int dtype, int n, const int* perm, const void* in, void* out);
// setup on the device (spmv.cu): sliced-ELL fill from the CSR arrays in HBM; rows of D^-1 P A P^T from those of A
int launch_sell_fill(bg_context_t ctx, int dtype, int n, int nslices, int slice, const int* rp, const int* ci, const void* v,
                     const int64_t* sp, int* cols, void* vals);
int launch_permute_rows(bg_context_t ctx, int dtype, int n, const int* rp_old, const int* ci_old, const void* v_old, const int* perm,
                        const int* inv, const double* dinv, const int* rp_new, int* ci_new, void* v_new);
int launch_diag_scale(bg_context_t ctx, int dtype, int n, const void* in, const void* d, void* out,
                      const int* skip_flag);
// out[perm[i]] = in[i]*scale[i]  (warm-start scatter), out[i] = in[i]/scale[i]
int launch_scatter_scale(bg_context_t ctx, int vec_dt, int n, const void* in, const void* scale, const int* perm,
                         void* out_ct);
int launch_divide_cast(bg_context_t ctx, int vec_dt, int n, const void* in_ct, const void* scale, void* out);
int launch_copy(bg_context_t ctx, int dtype, int n, const void* in, void* out, const int* skip_flag);

}  // namespace bg

struct bg_precond_s {
  bg_context_t ctx = nullptr;
  int kind = BG_PRECOND_NONE, n = 0, dtype = BG_F64, factor_dtype = BG_F64, schedule = BG_TRSV_LEVEL;
  bool unit_lower = true, use_tma = true, allow_lattice = true;
  int lg_levels = 0;
  bool device_factored = false;  // factor.cu built the streams
  int fb_chunk_bytes = 2048;
  int64_t nnzL = 0, nnzU = 0;
  bg::DevBuf dinv;                                 // DIAG, AINV (middle diagonal)
  bg_matrix_t m1 = nullptr, m2 = nullptr;          // AINV: explicit inverse factors (owned)
  ~bg_precond_s() { delete m1; delete m2; }
  bg::SyncFreeDev lsL, lsU;                        // BG_TRSV_LEVEL (any factor)
  bg::LatticeGlobalDev lg;                         // BG_TRSV_LEVEL (5-point ILU(0), natural order)
  std::vector<int> levelL, levelU;                 // getLevel of the factors (parity API)
  bg::FusedBlockDev fb;                            // BG_TRSV_BLOCK
  bg::DevBuf xsL, xsU;                             // sync-free solves: n flag-in-data packets each
  bg::DevBuf tmp, tmp2;                            // scratch vectors (compute type)
  bg::DevBuf middle, lscale, rscale, permRow, permCol, invPermCol;  // SPLIT
  bg::DevBuf u_rp, u_ci, u_val;                    // raw CSR U for the warm-start product
  int ct_dtype() const { return (dtype == BG_F64 || factor_dtype == BG_F64) ? BG_F64 : BG_F32; }
};

struct bg_solver_s {
  bg_context_t ctx = nullptr;
  int n = 0, m = 0, dtype = BG_F64;
  int last_m = 0;  // restart length of the last solve (H is laid out with ld = last_m + 1)
  int basis_dtype = BG_F64;  // storage type of V; != dtype: mixed-precision basis (FP64 work vectors, FP32 basis)
  bg::DevBuf vcur;           // mixed basis: working (dtype) copy of the newest basis vector, the SpMV's input
  int64_t ld = 0;
  bg::DevBuf V, w, ww, r, rr, bb, yv, state, hx, hb, bp, xp, red;
  bg::DevBuf mgs_slots;     // packets of the fused MGS step's grid-wide sums
  unsigned mgs_seq = 0;
  bg::GmresDev dev{};
  void* pinned = nullptr;  // host mirror of flags+scal for polling
  cudaEvent_t ev[2] = {nullptr, nullptr}, t0 = nullptr, t1 = nullptr;
  std::vector<cudaEvent_t> prof_events;
  std::vector<double> last_hist;
  int last_hist_len = 0;
};

// communication hooks of a row-partitioned solve (dist.cu); null for a single GPU
namespace bg {
struct DistHooks {
  virtual int exchange(void* vec, const int* skip) = 0;  // fill vec[n_local .. n_local+n_halo) from the neighbours
  virtual int allreduce(double* buf, int count) = 0;     // sum over ranks, in place, on the context stream
  virtual const PeerAR* peer_ar() { return nullptr; }    // non-null: dots_kernel all-reduces through peer memory itself
  // peer transport only: describe the NEXT halo exchange (consumes a sequence number) for a SpMV that carries it
  virtual bool halo_args(HaloArgs*) { return false; }
  virtual ~DistHooks() {}
};
int gmres_solve_dist(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                     const bg_gmres_opts* opts, bg_gmres_result* res, DistHooks* dh);
int gmres_profile_dist(bg_context_t ctx, bg_solver_t S, bg_matrix_t A, bg_precond_t M, const void* d_b, void* d_x,
                       const bg_gmres_opts* o, bg_gmres_result* r, bg_phase_stats* out, DistHooks* dh);
int matrix_create_internal(bg_context_t ctx, int n, int ncols, std::vector<int> rp, std::vector<int> ci,
                           const void* val, int val_dtype, int compute_dtype, int spmv_format, bg_matrix_t* out);
int solver_create_internal(bg_context_t ctx, int n, int m, int dtype, int extra_rows, bg_solver_t* out, int basis_dtype = -1);
int prepare_solve(bg_matrix_t A, bg_precond_t M);  // one-time host work of the first solve (block-major copy of A)
// factor.cu: block-Jacobi ILU(0) of the tiles of a 5-point grid factored and packed on the device (fills M->fb);
// returns 1 when the matrix is not such a grid (the caller takes the host path)
template <class T>
int lattice_tiles_factor_device(bg_context_t ctx, bg_precond_t M, int n, const int* rp, const int* ci, const void* v, int vdt,
                                const int* dims, uint64_t serial);
}  // namespace bg

struct bg_factors_s {
  int n = 0, dtype = BG_F64;
  bg::Factors<float> f32;
  bg::Factors<double> f64;
};
