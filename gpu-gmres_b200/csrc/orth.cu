// orth.cu -- Arnoldi orthogonalisation, Givens least squares and solution update, all on device.
//
// Replaces the BLAS-1 storm of the reference's inner loop (src/gmres.cu:2356-2363: (i+1) x
// {cublasSdot + blocking D2H scalar, cublasSaxpy}, cublasSnrm2, cublasScopy, cublasSscal), the
// host-side rotations (src/gmres.cu:192-216, 2365-2372) and Update_GPU (src/gmres.cu:164-189).
//
//  * dots_kernel<VT,J>: ONE pass over the Krylov basis V[:,0..j) computes all j inner products
//    V^T w and w.w with 128-bit loads, FP64 accumulators, warp-shuffle + shared-memory block
//    reduction, per-block partials and a deterministic fixed-order final reduction by the last
//    block to arrive (ticket).  That same block runs the scalar epilogue ("finalize"): the
//    Pythagorean norm, the DGKS re-orthogonalisation test, the Givens rotations, the residual
//    estimate and the convergence flag -- so no scalar ever travels to the host inside a cycle.
//  * update_kernel<VT,J>: out = (a + sgn * V[:,0..j) c) * scale in one pass (the MGS/CGS axpy
//    sweep fused with the normalisation cublasScopy+cublasSscal of the reference).
// HBM-bound: (2j+3)*s*n bytes per inner iteration (SURVEY 8d B_orth lower bound).
#include "internal.h"

namespace bg {

template <class VT> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int W = 4;
  using V = float4;
  __device__ static void unpack(const V& v, double* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  __device__ static V pack(const double* o) { return make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]); }
};
template <> struct VecIO<double> {
  static constexpr int W = 2;
  using V = double2;
  __device__ static void unpack(const V& v, double* o) { o[0] = v.x; o[1] = v.y; }
  __device__ static V pack(const double* o) { return make_double2(o[0], o[1]); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------- Givens (device)
// reference semantics: ApplyPlaneRotation / GeneratePlaneRotation, src/gmres.cu:192-216
__device__ inline void rot_apply(double& dx, double& dy, double cs, double sn) {
  const double t = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = t;
}
__device__ inline void rot_gen(double dx, double dy, double& cs, double& sn) {
  if (dy == 0.0) { cs = 1.0; sn = 0.0; }
  else if (fabs(dy) > fabs(dx)) { const double t = dx / dy; sn = 1.0 / sqrt(1.0 + t * t); cs = t * sn; }
  else { const double t = dy / dx; cs = 1.0 / sqrt(1.0 + t * t); sn = t * cs; }
}
// column i of H is complete: rotate, update s, residual estimate, convergence (gmres.cu:2365-2372)
__device__ void givens_column(const GmresDev& st, int i) {
  const int ld = st.m + 1;
  double* Hc = st.H + (size_t)i * ld;
  for (int k = 0; k < i; ++k) rot_apply(Hc[k], Hc[k + 1], st.cs[k], st.sn[k]);
  rot_gen(Hc[i], Hc[i + 1], st.cs[i], st.sn[i]);
  rot_apply(Hc[i], Hc[i + 1], st.cs[i], st.sn[i]);
  rot_apply(st.s[i], st.s[i + 1], st.cs[i], st.sn[i]);
  const double resid = fabs(st.s[i + 1]) / st.scal[S_NORMB];
  st.scal[S_RESID] = resid;
  const int it = st.flags[F_ITERS];
  if (st.hist && it < st.hist_cap) st.hist[it] = resid;
  st.flags[F_ITERS] = it + 1;
  if (resid < st.scal[S_TOL]) { st.flags[F_CONV] = 1; st.flags[F_KCONV] = i; }
}

__device__ void finalize(const GmresDev& st, int fin, int i, int mgs_k, int orth_mode, const double* red, int j) {
  const int ld = st.m + 1;
  const double ww = red[j];
  switch (fin) {
    case FIN_NORMB: {  // normb = ||M_L^-1 b||, zero -> 1 (gmres.cu:2294-2296)
      double nb = sqrt(ww);
      if (nb == 0.0) nb = 1.0;
      st.scal[S_NORMB] = nb;
    } break;
    case FIN_BETA0:
    case FIN_BETA: {  // residual check: "<=" for the initial one, "<" at restarts (gmres.cu:2312,2410)
      const double beta = sqrt(ww), resid = beta / st.scal[S_NORMB];
      st.scal[S_BETA] = beta;
      st.scal[S_RESID] = resid;
      const bool conv = fin == FIN_BETA0 ? resid <= st.scal[S_TOL] : resid < st.scal[S_TOL];
      if (conv) { st.flags[F_CONV] = 2; st.flags[F_KCONV] = -1; }
      else {
        for (int k = 0; k <= st.m; ++k) st.s[k] = 0.0;
        st.s[0] = beta;
        st.scal[S_SCALE] = 1.0 / beta;
        if (fin == FIN_BETA) st.flags[F_RESTARTS]++;
      }
    } break;
    case FIN_CGS1: {
      double sum = 0.0;
      for (int k = 0; k < j; ++k) { st.h[k] = red[k]; sum += red[k] * red[k]; }
      const double hn2 = ww - sum;
      st.scal[S_WW] = ww;
      if (orth_mode == BG_ORTH_CGS2 || !(hn2 >= st.scal[S_ETA2] * ww)) {
        st.flags[F_REORTH] = 1;   // second pass decides the column
        st.scal[S_SCALE] = 1.0;
      } else {
        st.flags[F_REORTH] = 0;
        double* Hc = st.H + (size_t)i * ld;
        for (int k = 0; k < j; ++k) Hc[k] = red[k];
        const double hn = sqrt(hn2);
        Hc[i + 1] = hn;
        if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
        givens_column(st, i);
      }
    } break;
    case FIN_CGS2: {
      double sum = 0.0;
      double* Hc = st.H + (size_t)i * ld;
      for (int k = 0; k < j; ++k) { st.h2[k] = red[k]; sum += red[k] * red[k]; Hc[k] = st.h[k] + red[k]; }
      double hn2 = ww - sum;
      if (hn2 < 0.0) hn2 = 0.0;
      const double hn = sqrt(hn2);
      Hc[i + 1] = hn;
      if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
      st.flags[F_NREORTH]++;
      givens_column(st, i);
    } break;
    case FIN_MGS_DOT: {  // H(k,i) = <w, v_k>  (gmres.cu:2357)
      st.H[(size_t)i * ld + mgs_k] = red[0];
      st.h[0] = red[0];
    } break;
    case FIN_MGS_NORM: {  // H(i+1,i) = ||w||  (gmres.cu:2360)
      const double hn = sqrt(ww);
      st.H[(size_t)i * ld + i + 1] = hn;
      if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
      givens_column(st, i);
    } break;
    default: break;
  }
}

// ---------------------------------------------------------------------------- dots
template <class VT, int J>
__global__ void __launch_bounds__(256) dots_kernel(int n, const VT* __restrict__ V, int64_t ld, int j,
                                                   const VT* __restrict__ w, GmresDev st, int fin, int inner_i,
                                                   int mgs_k, int orth_mode, int need_reorth) {
  if (st.flags[F_CONV]) return;
  if (need_reorth && !st.flags[F_REORTH]) return;
  constexpr int W = VecIO<VT>::W;
  using VV = typename VecIO<VT>::V;
  double acc[J + 1];
#pragma unroll
  for (int k = 0; k <= J; ++k) acc[k] = 0.0;
  const int64_t nvec = n / W;
  const VV* w4 = reinterpret_cast<const VV*>(w);
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nvec; c += (int64_t)gridDim.x * blockDim.x) {
    double wv[W];
    VecIO<VT>::unpack(w4[c], wv);
#pragma unroll
    for (int e = 0; e < W; ++e) acc[J] += wv[e] * wv[e];
#pragma unroll
    for (int k = 0; k < J; ++k) {
      if (k < j) {
        double vv[W];
        VecIO<VT>::unpack(*reinterpret_cast<const VV*>(V + (size_t)k * ld + c * W), vv);
#pragma unroll
        for (int e = 0; e < W; ++e) acc[k] += vv[e] * wv[e];
      }
    }
  }
  if (blockIdx.x == 0) {  // scalar tail (n not a multiple of the vector width)
    for (int64_t r = nvec * W + threadIdx.x; r < n; r += blockDim.x) {
      const double wr = (double)w[r];
      acc[J] += wr * wr;
#pragma unroll
      for (int k = 0; k < J; ++k)
        if (k < j) acc[k] += (double)V[(size_t)k * ld + r] * wr;
    }
  }
  __shared__ double sm[8][J + 1];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= J; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) sm[wid][k] = v;
  }
  __syncthreads();
  // slot j of the partial row carries w.w ; slots [0,j) the projections
  if (threadIdx.x <= j) {
    const int src = threadIdx.x == j ? J : threadIdx.x;
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += sm[q][src];
    st.partial[(size_t)blockIdx.x * st.partial_ld + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // deterministic final reduction: warp q owns slots q, q+8, ...; lanes stride over blocks
  __shared__ double red[J + 2];
  for (int k = wid; k <= j; k += 8) {
    double v = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(&st.partial[(size_t)b * st.partial_ld + k]);
    v = warp_sum(v);
    if (lane == 0) red[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    finalize(st, fin, inner_i, mgs_k, orth_mode, red, j);
    *st.ticket = 0u;
  }
}

// ---------------------------------------------------------------------------- update
template <class VT, int J>
__global__ void __launch_bounds__(256) update_kernel(int n, const VT* a, const VT* __restrict__ V,
                                                     int64_t ld, int j, const double* __restrict__ coef, double sgn,
                                                     VT* out, GmresDev st, int use_scale, int force,
                                                     int need_reorth) {
  if (!force && st.flags[F_CONV]) return;
  if (need_reorth && !st.flags[F_REORTH]) return;
  constexpr int W = VecIO<VT>::W;
  using VV = typename VecIO<VT>::V;
  double c[J > 0 ? J : 1];
#pragma unroll
  for (int k = 0; k < J; ++k) c[k] = k < j ? sgn * coef[k] : 0.0;
  const double scale = use_scale ? st.scal[S_SCALE] : 1.0;
  const int64_t nvec = n / W;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nvec; q += (int64_t)gridDim.x * blockDim.x) {
    double av[W];
    VecIO<VT>::unpack(*reinterpret_cast<const VV*>(a + q * W), av);
#pragma unroll
    for (int k = 0; k < J; ++k) {
      if (k < j) {
        double vv[W];
        VecIO<VT>::unpack(*reinterpret_cast<const VV*>(V + (size_t)k * ld + q * W), vv);
#pragma unroll
        for (int e = 0; e < W; ++e) av[e] += c[k] * vv[e];
      }
    }
#pragma unroll
    for (int e = 0; e < W; ++e) av[e] *= scale;
    *reinterpret_cast<VV*>(out + q * W) = VecIO<VT>::pack(av);
  }
  if (blockIdx.x == 0) {
    for (int64_t r = nvec * W + threadIdx.x; r < n; r += blockDim.x) {
      double v = (double)a[r];
#pragma unroll
      for (int k = 0; k < J; ++k)
        if (k < j) v += c[k] * (double)V[(size_t)k * ld + r];
      out[r] = (VT)(v * scale);
    }
  }
}

// back-substitution H y = s on one thread (Update, src/gmres.cu:92-105 / Update_GPU :172-181)
__global__ void solve_y_kernel(GmresDev st, int k, int force) {
  if (!force && st.flags[F_CONV]) return;
  if (threadIdx.x != 0) return;
  const int ld = st.m + 1;
  for (int i = 0; i <= k; ++i) st.y[i] = st.s[i];
  for (int i = k; i >= 0; --i) {
    st.y[i] /= st.H[(size_t)i * ld + i];
    for (int q = i - 1; q >= 0; --q) st.y[q] -= st.H[(size_t)i * ld + q] * st.y[i];
  }
}

__global__ void state_reset_kernel(GmresDev st, double tol, double eta) {
  const int t = threadIdx.x;
  if (t < F_COUNT) st.flags[t] = 0;
  if (t < S_COUNT) st.scal[t] = 0.0;
  if (t == 0) {
    st.flags[F_KCONV] = -1;
    st.scal[S_TOL] = tol;
    st.scal[S_ETA2] = eta * eta;
    st.scal[S_SCALE] = 1.0;
    *st.ticket = 0u;
  }
  for (int k = t; k <= st.m; k += blockDim.x) { st.s[k] = 0.0; st.cs[k] = 0.0; st.sn[k] = 0.0; }
}

template <class VT>
__global__ void diag_scale_kernel(int n, const VT* __restrict__ in, const VT* __restrict__ d, VT* __restrict__ out,
                                  const int* skip) {
  if (skip && *skip) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * d[i];
}
template <class VT>
__global__ void copy_kernel(int n, const VT* __restrict__ in, VT* __restrict__ out, const int* skip) {
  if (skip && *skip) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
// warm start of the split preconditioner (preconditioner.cu:1533-1545, 1485-1498)
template <class VT, class CT>
__global__ void scatter_scale_kernel(int n, const VT* __restrict__ in, const VT* __restrict__ scale,
                                     const int* __restrict__ perm, CT* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    VT v = in[i];
    v *= scale[i];
    out[perm[i]] = (CT)v;
  }
}
template <class VT, class CT>
__global__ void divide_cast_kernel(int n, const CT* __restrict__ in, const VT* __restrict__ scale, VT* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (VT)(in[i] / scale[i]);
}

// ---------------------------------------------------------------------------- launchers
static int vec_grid(bg_context_t ctx, int64_t n, int per_thread) {
  int64_t blocks = (n / per_thread + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}
static int round_j(int j) { return j == 0 ? 0 : j == 1 ? 1 : j <= 4 ? 4 : j <= 8 ? 8 : j <= 16 ? 16 : j <= 24 ? 24 : j <= 32 ? 32 : -1; }

template <class VT>
static int dots_dispatch(bg_context_t ctx, int n, const VT* V, int64_t ld, int j, const VT* w, const GmresDev& st,
                         int fin, int i, int mk, int om, int nr) {
  const int g = vec_grid(ctx, n, VecIO<VT>::W);
#define BG_D(JJ) dots_kernel<VT, JJ><<<g, 256, 0, ctx->stream>>>(n, V, ld, j, w, st, fin, i, mk, om, nr)
  switch (round_j(j)) {
    case 0: BG_D(0); break;
    case 1: BG_D(1); break;
    case 4: BG_D(4); break;
    case 8: BG_D(8); break;
    case 16: BG_D(16); break;
    case 24: BG_D(24); break;
    case 32: BG_D(32); break;
    default: set_error("dots: j=%d exceeds one pass (max 32 columns)", j); return BG_ERR_ARG;
  }
#undef BG_D
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_dots(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int j, const void* w, const GmresDev& st,
                int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth) {
  if (dtype == BG_F32) return dots_dispatch<float>(ctx, n, (const float*)V, ld, j, (const float*)w, st, fin, inner_i, mgs_k, orth_mode, need_reorth);
  return dots_dispatch<double>(ctx, n, (const double*)V, ld, j, (const double*)w, st, fin, inner_i, mgs_k, orth_mode, need_reorth);
}

template <class VT>
static int update_dispatch(bg_context_t ctx, int n, const VT* a, const VT* V, int64_t ld, int j, const double* coef,
                           double sgn, VT* out, const GmresDev& st, int us, int force, int nr) {
  const int g = vec_grid(ctx, n, VecIO<VT>::W);
#define BG_U(JJ) update_kernel<VT, JJ><<<g, 256, 0, ctx->stream>>>(n, a, V, ld, j, coef, sgn, out, st, us, force, nr)
  switch (round_j(j)) {
    case 0: BG_U(0); break;
    case 1: BG_U(1); break;
    case 4: BG_U(4); break;
    case 8: BG_U(8); break;
    case 16: BG_U(16); break;
    case 24: BG_U(24); break;
    case 32: BG_U(32); break;
    default: set_error("update: j=%d exceeds one pass (max 32 columns)", j); return BG_ERR_ARG;
  }
#undef BG_U
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_update(bg_context_t ctx, int dtype, int n, const void* a, const void* V, int64_t ld, int j,
                  const double* coef, double sgn, void* out, const GmresDev& st, int use_scale, int force,
                  int need_reorth) {
  if (dtype == BG_F32) return update_dispatch<float>(ctx, n, (const float*)a, (const float*)V, ld, j, coef, sgn, (float*)out, st, use_scale, force, need_reorth);
  return update_dispatch<double>(ctx, n, (const double*)a, (const double*)V, ld, j, coef, sgn, (double*)out, st, use_scale, force, need_reorth);
}

int launch_solve_y(bg_context_t ctx, const GmresDev& st, int k, int force) {
  solve_y_kernel<<<1, 32, 0, ctx->stream>>>(st, k, force);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_state_reset(bg_context_t ctx, const GmresDev& st, double tol, double eta) {
  state_reset_kernel<<<1, 256, 0, ctx->stream>>>(st, tol, eta);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_diag_scale(bg_context_t ctx, int dtype, int n, const void* in, const void* d, void* out, const int* skip) {
  const int g = vec_grid(ctx, n, 1);
  if (dtype == BG_F32) diag_scale_kernel<float><<<g, 256, 0, ctx->stream>>>(n, (const float*)in, (const float*)d, (float*)out, skip);
  else diag_scale_kernel<double><<<g, 256, 0, ctx->stream>>>(n, (const double*)in, (const double*)d, (double*)out, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_copy(bg_context_t ctx, int dtype, int n, const void* in, void* out, const int* skip) {
  const int g = vec_grid(ctx, n, 1);
  if (dtype == BG_F32) copy_kernel<float><<<g, 256, 0, ctx->stream>>>(n, (const float*)in, (float*)out, skip);
  else copy_kernel<double><<<g, 256, 0, ctx->stream>>>(n, (const double*)in, (double*)out, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_scatter_scale(bg_context_t ctx, int vec_dt, int n, const void* in, const void* scale, const int* perm,
                         void* out_ct) {
  const int g = vec_grid(ctx, n, 1);
  if (vec_dt == BG_F32) scatter_scale_kernel<float, double><<<g, 256, 0, ctx->stream>>>(n, (const float*)in, (const float*)scale, perm, (double*)out_ct);
  else scatter_scale_kernel<double, double><<<g, 256, 0, ctx->stream>>>(n, (const double*)in, (const double*)scale, perm, (double*)out_ct);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_divide_cast(bg_context_t ctx, int vec_dt, int n, const void* in_ct, const void* scale, void* out) {
  const int g = vec_grid(ctx, n, 1);
  if (vec_dt == BG_F32) divide_cast_kernel<float, double><<<g, 256, 0, ctx->stream>>>(n, (const double*)in_ct, (const float*)scale, (float*)out);
  else divide_cast_kernel<double, double><<<g, 256, 0, ctx->stream>>>(n, (const double*)in_ct, (const double*)scale, (double*)out);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

}  // namespace bg
