// trsv.cu -- sparse triangular solves for the ILU preconditioner, sm_100a.
//
// Replaces cusparse{S,D}csrsv_analysis/solve (src/preconditioner.cu:263,273,356-358,1313-1315,
// 1612,1637) and the nine element-wise scale/permute/convert kernels wrapped around them
// (src/preconditioner.cu:1424-1558), which are folded into the solve's prologue / epilogue here.
//
// Two schedules (DESIGN.md "SpTRSV"):
//  * LEVEL  -- the reference's global dependency levels (getLevel, src/gmres.cu:1033-1077), rows
//    stored level by level as ELL slabs.  Runs of thin levels are walked by ONE CTA with
//    __syncthreads between levels (no launch or grid barrier per level); fat levels get a launch
//    of their own.  Exact same operator as the reference; latency-bound by the dependency chain.
//  * BLOCK  -- block-Jacobi ILU, see trsv_block.cu.
#include "internal.h"

namespace bg {

// ------------------------------------------------------------------ prologue / epilogue
template <class VT> struct XformDev {
  int pro, epi;
  const int* perm_pro;
  const int* perm_epi;
  const VT* s_pro;
  const VT* s_epi;
};
template <class VT, class CT>
__device__ __forceinline__ CT load_rhs(const XformDev<VT>& xf, const VT* __restrict__ in, int r) {
  if (xf.pro == 1) {  // kernel_divide_S + kernel_perm_precondLeft_S2D (preconditioner.cu:1439-1464)
    const int p = xf.perm_pro[r];
    const VT t = in[p] / xf.s_pro[p];
    return (CT)t;
  }
  if (xf.pro == 2) {  // kernel_multiply_S2D (preconditioner.cu:1518-1531)
    const VT t = in[r] * xf.s_pro[r];
    return (CT)t;
  }
  return (CT)in[r];
}
template <class VT, class CT>
__device__ __forceinline__ void store_out(const XformDev<VT>& xf, VT* __restrict__ out, int r, CT x) {
  if (!out) return;
  if (xf.epi == 1) {  // kernel_perm_divide_precondRight_D2S (preconditioner.cu:1500-1516), as a scatter
    const int d = xf.perm_epi[r];
    VT t = (VT)x;
    t /= xf.s_epi[d];
    out[d] = t;
  } else {
    out[r] = (VT)x;
  }
}

// ------------------------------------------------------------------ LEVEL schedule
template <class FT> struct LsArgs {
  const int *lev_ptr, *lev_K, *rows, *cols;
  const int64_t* lev_off;
  const FT *vals, *diag;
};

template <class VT, class FT, class CT, bool UNIT>
__device__ __forceinline__ void ls_row(const LsArgs<FT>& a, int l, int t, int R, int K, int64_t off,
                                       const XformDev<VT>& xf, const VT* __restrict__ in, CT* xs, VT* out) {
  const int s = a.lev_ptr[l] + t;
  const int r = a.rows[s];
  CT acc = load_rhs<VT, CT>(xf, in, r);
  for (int k = 0; k < K; ++k) {
    const int c = a.cols[off + (int64_t)k * R + t];
    if (c >= 0) acc -= (CT)a.vals[off + (int64_t)k * R + t] * xs[c];
  }
  if (!UNIT) acc /= (CT)a.diag[s];
  xs[r] = acc;
  store_out<VT, CT>(xf, out, r, acc);
}

// one CTA walks levels [l0,l1): every level fits a few sweeps of the block.
// The walk is a chain of dependent L2 round trips, so everything that does not depend on the previous level is
// taken off the chain: the level table is staged in shared memory once, the row ids and column indices of level
// l+1 are fetched into registers while level l waits for its operands, and after the barrier the x gathers (the
// only loads that depend on level l) are issued together with the values, diagonal and right-hand side -- one
// round trip per level instead of three.  Rows with more than KP off-diagonals take the plain path.
constexpr int kWalkKP = 3, kWalkRP = 2;
template <class VT, class FT, class CT, bool UNIT>
__global__ void __launch_bounds__(1024) trsv_level_walk(LsArgs<FT> a, int l0, int l1, XformDev<VT> xf,
                                                        const VT* __restrict__ in, CT* xs, VT* out, const int* skip, int staged) {
  if (skip && *skip) return;
  extern __shared__ __align__(16) unsigned char lw_smem[];
  constexpr int KP = kWalkKP, RP = kWalkRP;
  const int nl = l1 - l0, tid = threadIdx.x, bd = blockDim.x;
  if (!staged) {  // level table too large for shared memory: the plain walk
    for (int l = l0; l < l1; ++l) {
      const int R = a.lev_ptr[l + 1] - a.lev_ptr[l], K = a.lev_K[l];
      const int64_t off = a.lev_off[l];
      for (int t = tid; t < R; t += bd) ls_row<VT, FT, CT, UNIT>(a, l, t, R, K, off, xf, in, xs, out);
      __syncthreads();  // orders this CTA's global writes of xs before the next level's reads
    }
    return;
  }
  int64_t* s_off = reinterpret_cast<int64_t*>(lw_smem);
  int* s_ptr = reinterpret_cast<int*>(s_off + nl);
  int* s_K = s_ptr + (nl + 1);
  for (int i = tid; i <= nl; i += bd) s_ptr[i] = a.lev_ptr[l0 + i];
  for (int i = tid; i < nl; i += bd) { s_K[i] = a.lev_K[l0 + i]; s_off[i] = a.lev_off[l0 + i]; }
  __syncthreads();
  int nr[RP], nc[RP][KP];  // row ids and columns of the NEXT level (independent of the dependency chain)
  auto is_fast = [&](int li) { return li < nl && s_K[li] <= KP && s_ptr[li + 1] - s_ptr[li] <= RP * bd; };
  auto prefetch = [&](int li) {
    const int p0 = s_ptr[li], R = s_ptr[li + 1] - p0, K = s_K[li];
    const int64_t off = s_off[li];
#pragma unroll
    for (int q = 0; q < RP; ++q) {
      const int t = tid + q * bd;
      nr[q] = -1;
      if (t < R) {
        nr[q] = a.rows[p0 + t];
#pragma unroll
        for (int k = 0; k < KP; ++k) nc[q][k] = k < K ? a.cols[off + (int64_t)k * R + t] : -1;
      }
    }
  };
  // the factor slabs stream from HBM exactly once per solve: pull the slab of a level into L2 a few levels
  // before it is needed (fire-and-forget, one prefetch per 128-byte line), so that the loads on the chain are L2 hits
  auto l2_prefetch = [&](int li) {
    if (li >= nl) return;
    const int p0 = s_ptr[li], R = s_ptr[li + 1] - p0, K = s_K[li];
    const int64_t off = s_off[li];
    for (int t = tid * 16; t < R; t += bd * 16) {
      for (int k = 0; k < K; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vals + off + (int64_t)k * R + t));
      if (!UNIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.diag + p0 + t));
    }
    for (int t = tid * 32; t < R; t += bd * 32) {
      for (int k = 0; k < K; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.cols + off + (int64_t)k * R + t));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rows + p0 + t));
    }
  };
  constexpr int kAhead = 4;
  for (int q = 0; q < kAhead; ++q) l2_prefetch(q);
  bool fast = is_fast(0);
  if (fast) prefetch(0);
  for (int li = 0; li < nl; ++li) {
    const int p0 = s_ptr[li], R = s_ptr[li + 1] - p0, K = s_K[li];
    const int64_t off = s_off[li];
    l2_prefetch(li + kAhead);
    if (fast) {
      int r[RP], c[RP][KP];
      CT acc[RP], dg[RP], xv[RP][KP], vv[RP][KP];
#pragma unroll
      for (int q = 0; q < RP; ++q) {
        r[q] = nr[q];
#pragma unroll
        for (int k = 0; k < KP; ++k) c[q][k] = nc[q][k];
      }
      // the x gathers (the chain) and every other operand of this level leave together
#pragma unroll
      for (int q = 0; q < RP; ++q) {
        if (r[q] < 0) continue;
        const int t = tid + q * bd;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          xv[q][k] = (CT)0; vv[q][k] = (CT)0;
          if (c[q][k] >= 0) { xv[q][k] = xs[c[q][k]]; vv[q][k] = (CT)a.vals[off + (int64_t)k * R + t]; }
        }
        acc[q] = load_rhs<VT, CT>(xf, in, r[q]);
        dg[q] = UNIT ? (CT)1 : (CT)a.diag[p0 + t];
      }
      const bool fast_next = is_fast(li + 1);
      if (fast_next) prefetch(li + 1);  // in flight while the gathers above return
#pragma unroll
      for (int q = 0; q < RP; ++q) {
        if (r[q] < 0) continue;
        CT v = acc[q];
#pragma unroll
        for (int k = 0; k < KP; ++k)
          if (c[q][k] >= 0) v -= vv[q][k] * xv[q][k];  // same order and operations as ls_row
        if (!UNIT) v /= dg[q];
        xs[r[q]] = v;
        store_out<VT, CT>(xf, out, r[q], v);
      }
      fast = fast_next;
    } else {
      for (int t = tid; t < R; t += bd) ls_row<VT, FT, CT, UNIT>(a, l0 + li, t, R, K, off, xf, in, xs, out);
      fast = is_fast(li + 1);
      if (fast) prefetch(li + 1);
    }
    __syncthreads();  // orders this CTA's global writes of xs before the next level's reads
  }
}
// one fat level spread over the whole GPU
template <class VT, class FT, class CT, bool UNIT>
__global__ void __launch_bounds__(256) trsv_level_wide(LsArgs<FT> a, int l, XformDev<VT> xf, const VT* __restrict__ in,
                                                       CT* xs, VT* out, const int* skip) {
  if (skip && *skip) return;
  const int R = a.lev_ptr[l + 1] - a.lev_ptr[l], K = a.lev_K[l];
  const int64_t off = a.lev_off[l];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < R; t += gridDim.x * blockDim.x)
    ls_row<VT, FT, CT, UNIT>(a, l, t, R, K, off, xf, in, xs, out);
}

template <class VT, class FT, bool UNIT>
static int level_run(bg_context_t ctx, const LevelStreamDev& S, const void* in, void* xs, void* out,
                     const TrsvXform& x, const int* skip) {
  using CT = decltype(VT() * FT());
  LsArgs<FT> a{S.lev_ptr.as<int>(), S.lev_K.as<int>(), S.rows.as<int>(), S.cols.as<int>(), S.lev_off.as<int64_t>(),
               S.vals.as<FT>(), S.diag.as<FT>()};
  XformDev<VT> xf{x.pro, x.epi, x.pro == 1 ? x.perm : nullptr, x.epi == 1 ? x.perm : nullptr,
                  (const VT*)x.scale_pro, (const VT*)x.scale_epi};
  // left-split and right-split never occur in the same solve, so one perm pointer suffices
  for (size_t g = 0; g < S.seg_begin.size(); ++g) {
    const int l0 = S.seg_begin[g], l1 = S.seg_end[g];
    if (S.seg_single[g]) {
      int maxR = 0;
      for (int l = l0; l < l1; ++l) maxR = std::max(maxR, S.h_lev_ptr[l + 1] - S.h_lev_ptr[l]);
      int threads = 32;
      while (threads < maxR && threads < 1024) threads <<= 1;
      // level table (16 bytes per level) staged in shared memory when it fits the opt-in limit
      const size_t tab = (size_t)(l1 - l0) * 16 + 16;
      const int staged = tab <= ctx->smem_optin - 4096 ? 1 : 0;
      auto kern = trsv_level_walk<VT, FT, CT, UNIT>;
      static bool attr_done[64] = {};
      if (staged && tab > 48 * 1024 && !attr_done[ctx->device & 63]) {
        BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 4096));
        attr_done[ctx->device & 63] = true;
      }
      kern<<<1, threads, staged ? tab : 0, ctx->stream>>>(a, l0, l1, xf, (const VT*)in, (CT*)xs, (VT*)out, skip, staged);
    } else {
      const int R = S.h_lev_ptr[l1] - S.h_lev_ptr[l0];
      int blocks = (R + 255) / 256;
      blocks = std::min(blocks, ctx->sm_count * 8);
      trsv_level_wide<VT, FT, CT, UNIT><<<blocks, 256, 0, ctx->stream>>>(a, l0, xf, (const VT*)in, (CT*)xs, (VT*)out, skip);
    }
    ctx->launches++;
  }
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_trsv_level(bg_context_t ctx, const LevelStreamDev& S, int vec_dt, int fac_dt, bool unit, const void* in,
                      void* xs, void* out, const TrsvXform& xf, const int* skip) {
  if (S.n == 0) return BG_OK;
#define BG_L(VT, FT) (unit ? level_run<VT, FT, true>(ctx, S, in, xs, out, xf, skip) : level_run<VT, FT, false>(ctx, S, in, xs, out, xf, skip))
  if (vec_dt == BG_F32 && fac_dt == BG_F32) return BG_L(float, float);
  if (vec_dt == BG_F32 && fac_dt == BG_F64) return BG_L(float, double);
  if (vec_dt == BG_F64 && fac_dt == BG_F32) return BG_L(double, float);
  return BG_L(double, double);
#undef BG_L
}

}  // namespace bg
