// spmv.cu -- y = A x (and y = alpha A x + beta b) for CSR and sliced-ELL storage, sm_100a.
//
// Replaces the reference's half-warp-per-row kernel (src/SpMV_kernel.cu:166-221: 16 lanes per row,
// 32-bit scalar loads, shared-memory round trip, serial 16-term add by lane 0) and the
// cusparseScsrmv calls (src/gmres.cu:2303,2351,2403).
//
//  * sliced ELL (stencil-like rows, chosen by the inspector): one thread owns RPT consecutive rows
//    of a 32*RPT-row slice; entry k of those rows is contiguous, so values and column indices are
//    fetched with 128-bit loads (RPT=4 for FP32, int4 + float4; RPT=2 for FP64, int2 + double2)
//    and a warp reads 512 contiguous bytes per load.  No lane is idle on 5- or 7-point rows.
//  * CSR vector: LANES (2..32, from the mean row length) lanes per row, shuffle reduction.
// x is gathered through L1/L2 (it fits the 126 MB L2 for every single-GPU config of BASELINE.json).
// HBM-bound: algorithmic bytes (s+4)*nnz + 4(n+1) + 2*s*n.
#include "internal.h"
#include "peer.cuh"

#include <cstdlib>

namespace bg {

template <class T> struct Vec;
template <> struct Vec<float> { using v4 = float4; };
template <> struct Vec<double> { using v2 = double2; };

template <class VT, class MT, int LANES, bool AXPBY>
__global__ void __launch_bounds__(256) spmv_csr_vector(int n, const int* __restrict__ rp, const int* __restrict__ ci,
                                                        const MT* __restrict__ val, const VT* __restrict__ x,
                                                        VT* __restrict__ y, double alpha, double beta,
                                                        const VT* __restrict__ b, const int* skip) {
  if (skip && *skip) return;
  using CT = decltype(MT() * VT());
  const int lane = threadIdx.x % LANES;
  const int rows_per_block = blockDim.x / LANES;
  for (int base = blockIdx.x * rows_per_block; base < n; base += gridDim.x * rows_per_block) {
    const int row = base + threadIdx.x / LANES;  // loop bound is block-uniform: shuffles stay converged
    const bool valid = row < n;
    const int lb = valid ? rp[row] : 0, ub = valid ? rp[row + 1] : 0;
    CT acc = 0;
    for (int p = lb + lane; p < ub; p += LANES) acc += val[p] * x[ci[p]];
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
    if (valid && lane == 0) {
      if (AXPBY) y[row] = (VT)(alpha * (double)acc + beta * (double)b[row]);
      else y[row] = (VT)acc;
    }
  }
}

// sliced ELL, FP32: 4 rows per thread, slice height 128
template <bool AXPBY>
__global__ void __launch_bounds__(128) spmv_sell_f32(int n, int nslices, const int64_t* __restrict__ sp,
                                                     const int* __restrict__ cols, const float* __restrict__ vals,
                                                     const float* __restrict__ x, float* __restrict__ y, float alpha,
                                                     float beta, const float* __restrict__ b, const int* skip) {
  if (skip && *skip) return;
  const int warps_per_block = blockDim.x / 32, lane = threadIdx.x & 31;
  for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < nslices; s += gridDim.x * warps_per_block) {
    const int64_t base = sp[s];
    const int width = (int)((sp[s + 1] - base) >> 7);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int4* c4 = reinterpret_cast<const int4*>(cols + base) + lane;
    const float4* v4 = reinterpret_cast<const float4*>(vals + base) + lane;
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int4 c = __ldg(c4 + k * 32);
      const float4 v = __ldg(v4 + k * 32);
      a0 += v.x * x[c.x]; a1 += v.y * x[c.y]; a2 += v.z * x[c.z]; a3 += v.w * x[c.w];
    }
    const int row = s * 128 + lane * 4;
    if (row + 3 < n) {
      float4 o = make_float4(a0, a1, a2, a3);
      if (AXPBY) {
        const float4 bb = *reinterpret_cast<const float4*>(b + row);
        o.x = alpha * o.x + beta * bb.x; o.y = alpha * o.y + beta * bb.y;
        o.z = alpha * o.z + beta * bb.z; o.w = alpha * o.w + beta * bb.w;
      }
      *reinterpret_cast<float4*>(y + row) = o;
    } else {
      const float a[4] = {a0, a1, a2, a3};
      for (int r = 0; r < 4; ++r)
        if (row + r < n) y[row + r] = AXPBY ? alpha * a[r] + beta * b[row + r] : a[r];
    }
  }
}

// sliced ELL, FP64: 2 rows per thread, slice height 64
template <bool AXPBY>
__global__ void __launch_bounds__(128) spmv_sell_f64(int n, int nslices, const int64_t* __restrict__ sp,
                                                     const int* __restrict__ cols, const double* __restrict__ vals,
                                                     const double* __restrict__ x, double* __restrict__ y,
                                                     double alpha, double beta, const double* __restrict__ b,
                                                     const int* skip) {
  if (skip && *skip) return;
  const int warps_per_block = blockDim.x / 32, lane = threadIdx.x & 31;
  for (int s = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < nslices; s += gridDim.x * warps_per_block) {
    const int64_t base = sp[s];
    const int width = (int)((sp[s + 1] - base) >> 6);
    double a0 = 0.0, a1 = 0.0;
    const int2* c2 = reinterpret_cast<const int2*>(cols + base) + lane;
    const double2* v2 = reinterpret_cast<const double2*>(vals + base) + lane;
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int2 c = __ldg(c2 + k * 32);
      const double2 v = __ldg(v2 + k * 32);
      a0 += v.x * x[c.x]; a1 += v.y * x[c.y];
    }
    const int row = s * 64 + lane * 2;
    if (row + 1 < n) {
      double2 o = make_double2(a0, a1);
      if (AXPBY) {
        const double2 bb = *reinterpret_cast<const double2*>(b + row);
        o.x = alpha * o.x + beta * bb.x; o.y = alpha * o.y + beta * bb.y;
      }
      *reinterpret_cast<double2*>(y + row) = o;
    } else if (row < n) {
      y[row] = AXPBY ? alpha * a0 + beta * b[row] : a0;
    }
  }
}

// sliced ELL with the SAME width W in every slice (stencil matrices: 5- and 7-point grids): no slice pointer
// to chase, all 2W 128-bit loads of a thread are issued before the first use, then all W*RPT gathers of x --
// two dependent memory round trips per slice instead of 2+W.  HALO = the kernel carries the halo exchange of
// a row-partitioned run (internal.h HaloArgs): columns >= n_local are read from the peer-written inbox.
template <class VT> struct SellVec;
// loads of the matrix stream as volatile asm: issued in program order, i.e. ALL of a thread's index and
// value loads leave before the first dependent gather (the compiler otherwise interleaves them to save registers)
__device__ __forceinline__ int4 ldnc(const int4* p) { int4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ int2 ldnc(const int2* p) { int2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ float4 ldnc(const float4* p) { float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ double2 ldnc(const double2* p) { double2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
template <> struct SellVec<float> {
  static constexpr int RPT = 4;
  using VI = int4; using VV = float4;
  __device__ static void idx(const VI& c, int* o) { o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
  __device__ static void val(const VV& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct SellVec<double> {
  static constexpr int RPT = 2;
  using VI = int2; using VV = double2;
  __device__ static void idx(const VI& c, int* o) { o[0] = c.x; o[1] = c.y; }
  __device__ static void val(const VV& v, double* o) { o[0] = v.x; o[1] = v.y; }
};
template <class VT, int W, bool AXPBY, bool HALO>
__global__ void __launch_bounds__(128, 8) spmv_sell_w(int n, int nslices, const int* __restrict__ cols, const VT* __restrict__ vals,
                                                   const VT* __restrict__ x, VT* __restrict__ y, VT alpha, VT beta,
                                                   const VT* __restrict__ b, const int* skip, HaloArgs h) {
  // launched with programmatic stream serialisation: everything above this line overlapped the tail of the
  // previous kernel; x, b and the skip flag are its outputs
  pdl_wait();
  pdl_launch_dependents();  // the triangular solve that follows may start its prologue (factor stream prefetch) now
  if (skip && *skip) return;
  using SV = SellVec<VT>;
  constexpr int RPT = SV::RPT, H = 32 * RPT;
  if (HALO) {
    // the first wave of CTAs (always scheduled before anything can wait) stores this rank's boundary
    // entries into the neighbours' inboxes; nothing in this phase depends on another rank
    const int pushers = min((int)gridDim.x, 148) * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < pushers) {
      const PeerHalo* P = h.ph;
      for (int e = tid; e < h.n_send; e += pushers) {
        int k = 0;
        while (e >= P->send_ptr[k + 1]) ++k;
        ll_store(P->peer_inbox[k] + (size_t)(h.seq & 1u) * P->peer_half[k] + (e - P->send_ptr[k]), (double)x[h.send_idx[e]], h.seq);
      }
    }
  }
  const int warps_per_block = blockDim.x / 32, lane = threadIdx.x & 31;
  // HALO: only the slices flagged by the host (the rank's boundary rows) read halo slots; every other slice takes the
  // plain gather without a per-operand test (the test on every operand cost 28 % of the kernel at 5 M rows per GPU).
  // The walk starts in the middle of the range: in the strip partitions the boundary rows are the first and the last
  // slices, and they are reached when the neighbours' values have had half an SpMV to arrive.
  const int rot = HALO ? nslices / 2 : 0;
  for (int s0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); s0 < nslices; s0 += gridDim.x * warps_per_block) {
    const int s = HALO ? (s0 + rot >= nslices ? s0 + rot - nslices : s0 + rot) : s0;
    const size_t base = (size_t)s * H * W;
    const typename SV::VI* cp = reinterpret_cast<const typename SV::VI*>(cols + base) + lane;
    const typename SV::VV* vp = reinterpret_cast<const typename SV::VV*>(vals + base) + lane;
    typename SV::VI cc[W];
    typename SV::VV vv[W];
#pragma unroll
    for (int k = 0; k < W; ++k) cc[k] = ldnc(cp + k * 32);
#pragma unroll
    for (int k = 0; k < W; ++k) vv[k] = ldnc(vp + k * 32);
    VT xv[W][RPT];
    if (HALO && h.slice_halo[s]) {  // warp-uniform
#pragma unroll
      for (int k = 0; k < W; ++k) {
        int ci[RPT];
        SV::idx(cc[k], ci);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          if (ci[r] >= h.n_local) xv[k][r] = (VT)ll_wait(h.ph->my_inbox + (size_t)(h.seq & 1u) * max(h.ph->n_halo, 1) + (ci[r] - h.n_local), h.seq, h.ph->timeout_ns);
          else xv[k][r] = x[ci[r]];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) {
        int ci[RPT];
        SV::idx(cc[k], ci);
#pragma unroll
        for (int r = 0; r < RPT; ++r) xv[k][r] = x[ci[r]];
      }
    }
    VT acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      VT va[RPT];
      SV::val(vv[k], va);
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc[r] += va[r] * xv[k][r];
    }
    const int row = s * H + lane * RPT;
#pragma unroll
    for (int r = 0; r < RPT; ++r)
      if (row + r < n) y[row + r] = AXPBY ? alpha * acc[r] + beta * b[row + r] : acc[r];
  }
}
template <class VT, int W>
static void sell_w_launch(bg_context_t ctx, bg_matrix_t A, int blocks, const VT* x, VT* y, double alpha, double beta, const VT* b,
                          bool axpby, const int* skip, const HaloArgs* halo) {
  const int n = A->n, ns = A->sell_slices;
  const int* c = A->s_col.as<int>();
  const VT* v = A->s_val.as<VT>();
  HaloArgs h = halo ? *halo : HaloArgs();
  h.slice_halo = A->s_halo.as<unsigned char>();
  if (halo) {
    if (axpby) launch_pdl(spmv_sell_w<VT, W, true, true>, dim3(blocks), dim3(128), 0, ctx->stream, n, ns, c, v, x, y, (VT)alpha, (VT)beta, b, skip, h);
    else launch_pdl(spmv_sell_w<VT, W, false, true>, dim3(blocks), dim3(128), 0, ctx->stream, n, ns, c, v, x, y, (VT)1, (VT)0, nullptr, skip, h);
  } else {
    if (axpby) launch_pdl(spmv_sell_w<VT, W, true, false>, dim3(blocks), dim3(128), 0, ctx->stream, n, ns, c, v, x, y, (VT)alpha, (VT)beta, b, skip, h);
    else launch_pdl(spmv_sell_w<VT, W, false, false>, dim3(blocks), dim3(128), 0, ctx->stream, n, ns, c, v, x, y, (VT)1, (VT)0, nullptr, skip, h);
  }
}
static bool sell_w_supported(int w) { return w == 3 || w == 5 || w == 7; }
bool spmv_supports_halo(bg_matrix_t A) { return A->format == BG_SPMV_SELL && sell_w_supported(A->sell_width) && A->s_halo.p != nullptr; }

template <class VT, class MT, int LANES>
static void csr_launch(bg_context_t ctx, int n, const int* rp, const int* ci, const MT* val, const VT* x, VT* y,
                       double alpha, double beta, const VT* b, bool axpby, const int* skip) {
  const int rows_per_block = 256 / LANES;
  int64_t blocks = ((int64_t)n + rows_per_block - 1) / rows_per_block;
  const int64_t cap = (int64_t)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (axpby) spmv_csr_vector<VT, MT, LANES, true><<<(int)blocks, 256, 0, ctx->stream>>>(n, rp, ci, val, x, y, alpha, beta, b, skip);
  else spmv_csr_vector<VT, MT, LANES, false><<<(int)blocks, 256, 0, ctx->stream>>>(n, rp, ci, val, x, y, alpha, beta, b, skip);
}

template <class VT, class MT>
static void csr_dispatch(bg_context_t ctx, int lanes, int n, const int* rp, const int* ci, const MT* val, const VT* x,
                         VT* y, double alpha, double beta, const VT* b, bool axpby, const int* skip) {
  switch (lanes) {
    case 2: csr_launch<VT, MT, 2>(ctx, n, rp, ci, val, x, y, alpha, beta, b, axpby, skip); break;
    case 4: csr_launch<VT, MT, 4>(ctx, n, rp, ci, val, x, y, alpha, beta, b, axpby, skip); break;
    case 8: csr_launch<VT, MT, 8>(ctx, n, rp, ci, val, x, y, alpha, beta, b, axpby, skip); break;
    case 16: csr_launch<VT, MT, 16>(ctx, n, rp, ci, val, x, y, alpha, beta, b, axpby, skip); break;
    default: csr_launch<VT, MT, 32>(ctx, n, rp, ci, val, x, y, alpha, beta, b, axpby, skip); break;
  }
}

int launch_spmv(bg_context_t ctx, bg_matrix_t A, double alpha, const void* x, double beta, const void* b, void* y,
                bool axpby, const int* skip, const HaloArgs* halo) {
  const int n = A->n;
  if (n == 0) return BG_OK;
  if (halo && !spmv_supports_halo(A)) { set_error("this matrix layout cannot carry the halo exchange inside the SpMV"); return BG_ERR_UNSUPPORTED; }
  if (A->format == BG_SPMV_SELL) {
    const int wpb = 4;
    int64_t blocks = ((int64_t)A->sell_slices + wpb - 1) / wpb;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    static const bool use_w = [] { const char* e = getenv("BG_SPMV_FIXED_WIDTH"); return !(e && e[0] == '0'); }();
    if (sell_w_supported(A->sell_width) && (use_w || halo)) {
      const int w = A->sell_width;
      if (A->dtype == BG_F32) {
        if (w == 3) sell_w_launch<float, 3>(ctx, A, (int)blocks, (const float*)x, (float*)y, alpha, beta, (const float*)b, axpby, skip, halo);
        else if (w == 5) sell_w_launch<float, 5>(ctx, A, (int)blocks, (const float*)x, (float*)y, alpha, beta, (const float*)b, axpby, skip, halo);
        else sell_w_launch<float, 7>(ctx, A, (int)blocks, (const float*)x, (float*)y, alpha, beta, (const float*)b, axpby, skip, halo);
      } else {
        if (w == 3) sell_w_launch<double, 3>(ctx, A, (int)blocks, (const double*)x, (double*)y, alpha, beta, (const double*)b, axpby, skip, halo);
        else if (w == 5) sell_w_launch<double, 5>(ctx, A, (int)blocks, (const double*)x, (double*)y, alpha, beta, (const double*)b, axpby, skip, halo);
        else sell_w_launch<double, 7>(ctx, A, (int)blocks, (const double*)x, (double*)y, alpha, beta, (const double*)b, axpby, skip, halo);
      }
    } else if (A->dtype == BG_F32) {
      if (axpby) spmv_sell_f32<true><<<(int)blocks, 128, 0, ctx->stream>>>(n, A->sell_slices, A->s_ptr.as<int64_t>(), A->s_col.as<int>(), A->s_val.as<float>(), (const float*)x, (float*)y, (float)alpha, (float)beta, (const float*)b, skip);
      else spmv_sell_f32<false><<<(int)blocks, 128, 0, ctx->stream>>>(n, A->sell_slices, A->s_ptr.as<int64_t>(), A->s_col.as<int>(), A->s_val.as<float>(), (const float*)x, (float*)y, 1.f, 0.f, nullptr, skip);
    } else {
      if (axpby) spmv_sell_f64<true><<<(int)blocks, 128, 0, ctx->stream>>>(n, A->sell_slices, A->s_ptr.as<int64_t>(), A->s_col.as<int>(), A->s_val.as<double>(), (const double*)x, (double*)y, alpha, beta, (const double*)b, skip);
      else spmv_sell_f64<false><<<(int)blocks, 128, 0, ctx->stream>>>(n, A->sell_slices, A->s_ptr.as<int64_t>(), A->s_col.as<int>(), A->s_val.as<double>(), (const double*)x, (double*)y, 1.0, 0.0, nullptr, skip);
    }
  } else {
    if (A->dtype == BG_F32)
      csr_dispatch<float, float>(ctx, A->lanes, n, A->rp.as<int>(), A->ci.as<int>(), A->val.as<float>(), (const float*)x, (float*)y, alpha, beta, (const float*)b, axpby, skip);
    else
      csr_dispatch<double, double>(ctx, A->lanes, n, A->rp.as<int>(), A->ci.as<int>(), A->val.as<double>(), (const double*)x, (double*)y, alpha, beta, (const double*)b, axpby, skip);
  }
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_csr_spmv_mixed(bg_context_t ctx, int n, const int* rp, const int* ci, const void* val, int val_dt,
                          const void* x, void* y, int vec_dt, const int* skip) {
  if (n == 0) return BG_OK;
  if (val_dt == BG_F64 && vec_dt == BG_F64)
    csr_dispatch<double, double>(ctx, 4, n, rp, ci, (const double*)val, (const double*)x, (double*)y, 1, 0, nullptr, false, skip);
  else if (val_dt == BG_F32 && vec_dt == BG_F32)
    csr_dispatch<float, float>(ctx, 4, n, rp, ci, (const float*)val, (const float*)x, (float*)y, 1, 0, nullptr, false, skip);
  else if (val_dt == BG_F32 && vec_dt == BG_F64)
    csr_dispatch<double, float>(ctx, 4, n, rp, ci, (const float*)val, (const double*)x, (double*)y, 1, 0, nullptr, false, skip);
  else {
    set_error("csr_spmv_mixed: FP64 values with FP32 vectors are staged through FP64 vectors");
    return BG_ERR_UNSUPPORTED;
  }
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}


// ------------------------------------------------------------------ small helpers behind the C ABI
// fingerprint of a device buffer (two position-weighted 64-bit sums of its 32-bit words): lets the reference-API
// shim notice that a caller refilled or re-allocated a matrix it handed over as raw device pointers
__global__ void __launch_bounds__(256) checksum_kernel(const unsigned* __restrict__ w, size_t nwords, unsigned long long* out) {
  unsigned long long s1 = 0, s2 = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long v = w[i];
    s1 += v;
    s2 += v * (2ull * i + 1ull);
  }
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, s1); atomicAdd(out + 1, s2); }
}
// x += V[:, 0..k) y   (Update_GPU, src/gmres.cu:164-189, after the host back-substitution)
template <class VT> __global__ void __launch_bounds__(256) basis_update_kernel(int n, int64_t ld, int k, const VT* __restrict__ V,
                                                                               const double* __restrict__ y, VT* x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < k; ++c) acc += (double)V[i + (size_t)c * ld] * y[c];
    x[i] = (VT)((double)x[i] + acc);
  }
}

// ---- one-time setup on the device: the sliced-ELL copy of a CSR matrix already in HBM, and its block-major reordering.
// Both used to be host loops over the whole matrix followed by an upload from pageable memory (1 s + 0.9 s at 10 M
// rows, tools/first_solve.py); the kernels move the same 0.6 GB at HBM speed and write exactly the host's arrays.

// host_algos.cpp build_sell: entry k of row s*slice + t at slice_ptr[s] + k*slice + t; padding = (own row, 0)
template <class T>
__global__ void __launch_bounds__(256) sell_fill_kernel(int n, int nslices, int slice, const int* __restrict__ rp, const int* __restrict__ ci,
                                                        const T* __restrict__ v, const int64_t* __restrict__ sp, int* __restrict__ cols,
                                                        T* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = i / slice;
  if (s >= nslices) return;
  const int t = (int)(i - s * slice);
  const int64_t base = sp[s], wid = (sp[s + 1] - base) / slice;
  const int p0 = i < n ? rp[i] : 0, len = i < n ? rp[i + 1] - p0 : 0;
  const int own = i < n ? (int)i : 0;
  for (int64_t k = 0; k < wid; ++k) {
    const int64_t q = base + k * slice + t;
    const bool in = k < len;
    cols[q] = in ? ci[p0 + k] : own;
    vals[q] = in ? v[p0 + k] : (T)0;
  }
}
int launch_sell_fill(bg_context_t ctx, int dtype, int n, int nslices, int slice, const int* rp, const int* ci, const void* v,
                     const int64_t* sp, int* cols, void* vals) {
  if (nslices <= 0) return BG_OK;
  const int64_t rows = (int64_t)nslices * slice;
  const unsigned blocks = (unsigned)((rows + 255) / 256);
  if (dtype == BG_F32) sell_fill_kernel<float><<<blocks, 256, 0, ctx->stream>>>(n, nslices, slice, rp, ci, (const float*)v, sp, cols, (float*)vals);
  else sell_fill_kernel<double><<<blocks, 256, 0, ctx->stream>>>(n, nslices, slice, rp, ci, (const double*)v, sp, cols, (double*)vals);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

// solver.cu build_permuted: row i of D^-1 P A P^T = row perm[i] of A, columns through inv (halo columns >= n keep their
// slot), values times dinv[i] in double, columns ascending (insertion sort: rows of this path are short and distinct)
template <class T>
__global__ void __launch_bounds__(256) permute_rows_kernel(int n, const int* __restrict__ rp_old, const int* __restrict__ ci_old,
                                                           const T* __restrict__ v_old, const int* __restrict__ perm,
                                                           const int* __restrict__ inv, const double* __restrict__ dinv,
                                                           const int* __restrict__ rp_new, int* __restrict__ ci_new, T* __restrict__ v_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = perm[i], a = rp_old[o], len = rp_old[o + 1] - a, q0 = rp_new[i];
  const double di = dinv[i];
  for (int k = 0; k < len; ++k) {
    const int c = ci_old[a + k];
    const int c2 = c < n ? inv[c] : c;
    const T val = (T)__dmul_rn((double)v_old[a + k], di);
    int j = k;
    while (j > 0 && ci_new[q0 + j - 1] > c2) { ci_new[q0 + j] = ci_new[q0 + j - 1]; v_new[q0 + j] = v_new[q0 + j - 1]; --j; }
    ci_new[q0 + j] = c2;
    v_new[q0 + j] = val;
  }
}
int launch_permute_rows(bg_context_t ctx, int dtype, int n, const int* rp_old, const int* ci_old, const void* v_old, const int* perm,
                        const int* inv, const double* dinv, const int* rp_new, int* ci_new, void* v_new) {
  if (n <= 0) return BG_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == BG_F32) permute_rows_kernel<float><<<blocks, 256, 0, ctx->stream>>>(n, rp_old, ci_old, (const float*)v_old, perm, inv, dinv, rp_new, ci_new, (float*)v_new);
  else permute_rows_kernel<double><<<blocks, 256, 0, ctx->stream>>>(n, rp_old, ci_old, (const double*)v_old, perm, inv, dinv, rp_new, ci_new, (double*)v_new);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

}  // namespace bg

extern "C" int bg_device_checksum(bg_context_t ctx, const void* d_ptr, size_t bytes, uint64_t* out2) {
  if (!ctx || !out2 || (bytes && !d_ptr) || (bytes & 3)) return BG_ERR_ARG;
  BG_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* d = nullptr;
  BG_CUDA(cudaMalloc(&d, 16));
  BG_CUDA(cudaMemsetAsync(d, 0, 16, ctx->stream));
  const size_t nw = bytes / 4;
  if (nw) {
    const int blocks = (int)std::min<size_t>((size_t)ctx->sm_count * 8, (nw + 255) / 256);
    bg::checksum_kernel<<<blocks, 256, 0, ctx->stream>>>((const unsigned*)d_ptr, nw, d);
    ctx->launches++;
  }
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return bg::cuda_fail(e, "bg_device_checksum", __FILE__, __LINE__);
  out2[0] = h[0]; out2[1] = h[1];
  return BG_OK;
}

extern "C" int bg_basis_update(bg_context_t ctx, int dtype, int n, int64_t ld, int k, const void* d_V, const double* h_y,
                               void* d_x) {
  if (!ctx || n < 0 || k < 0 || k > 4096 || ld < n || (k && (!d_V || !h_y)) || !d_x) return BG_ERR_ARG;
  if (n == 0 || k == 0) return BG_OK;
  BG_CUDA(cudaSetDevice(ctx->device));
  double* dy = nullptr;
  BG_CUDA(cudaMalloc(&dy, sizeof(double) * k));
  cudaError_t e = cudaMemcpyAsync(dy, h_y, sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream);
  const int blocks = std::min(ctx->sm_count * 8, (n + 255) / 256);
  if (e == cudaSuccess) {
    if (dtype == BG_F32) bg::basis_update_kernel<float><<<blocks, 256, 0, ctx->stream>>>(n, ld, k, (const float*)d_V, dy, (float*)d_x);
    else bg::basis_update_kernel<double><<<blocks, 256, 0, ctx->stream>>>(n, ld, k, (const double*)d_V, dy, (double*)d_x);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dy);
  if (e != cudaSuccess) return bg::cuda_fail(e, "bg_basis_update", __FILE__, __LINE__);
  return BG_OK;
}
