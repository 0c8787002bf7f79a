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
//  * BLOCK  -- block-Jacobi ILU: each block's triangular solve is independent, one WARP solves a
//    block entirely out of shared memory.  The block's factor stream (schedule, 16-bit local
//    columns, values, diagonal; packed level-major on the host) is staged by bulk-async copies
//    (cp.async.bulk + mbarrier transaction counts, SASS UBLKCP) into a per-warp ring, so HBM reads
//    are large, sequential and fully decoupled from the dependent FMA chain.
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

// one CTA walks levels [l0,l1): every level fits a few sweeps of the block
template <class VT, class FT, class CT, bool UNIT>
__global__ void __launch_bounds__(1024) trsv_level_walk(LsArgs<FT> a, int l0, int l1, XformDev<VT> xf,
                                                        const VT* __restrict__ in, CT* xs, VT* out, const int* skip) {
  if (skip && *skip) return;
  for (int l = l0; l < l1; ++l) {
    const int R = a.lev_ptr[l + 1] - a.lev_ptr[l], K = a.lev_K[l];
    const int64_t off = a.lev_off[l];
    for (int t = threadIdx.x; t < R; t += blockDim.x) ls_row<VT, FT, CT, UNIT>(a, l, t, R, K, off, xf, in, xs, out);
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
      trsv_level_walk<VT, FT, CT, UNIT><<<1, threads, 0, ctx->stream>>>(a, l0, l1, xf, (const VT*)in, (CT*)xs, (VT*)out, skip);
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

// ------------------------------------------------------------------ BLOCK schedule
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a scheduling bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
// global -> shared bulk-async copy completing on an mbarrier (TMA unit, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct BsArgs {
  const int *grow, *blk_row_ptr, *blk_lev_ptr, *lev_slot;
  const int64_t* blk_chunk_ptr;
  const unsigned char* stream;
  int nblocks, K, CS, max_rows, stages;
  uint32_t chunk_bytes, chunk_stride;  // bytes copied / ring slot stride (128-byte aligned)
  uint32_t warp_smem;                  // bytes of shared memory per warp
};

template <class VT, class FT, class CT, bool UNIT, bool TMA>
__global__ void __launch_bounds__(128) trsv_block_kernel(BsArgs a, const VT* __restrict__ in, VT* __restrict__ out,
                                                         const int* skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char* mine = smem + (size_t)wid * a.warp_smem;
  CT* xl = reinterpret_cast<CT*>(mine);
  const uint32_t xl_bytes = ((uint32_t)a.max_rows * sizeof(CT) + 127u) & ~127u;
  unsigned char* ring = mine + xl_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * a.chunk_stride);
  if (TMA) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) mbar_init(&bars[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const int CS = a.CS, K = a.K;
  uint32_t base_g = 0;  // ring-global index of this block's first chunk (monotone across blocks)
  for (int b = blockIdx.x * nw + wid; b < a.nblocks; b += gridDim.x * nw) {
    const int rbase = a.blk_row_ptr[b], nb = a.blk_row_ptr[b + 1] - rbase;
    const int64_t ch0 = a.blk_chunk_ptr[b];
    const int nch = (int)(a.blk_chunk_ptr[b + 1] - ch0);
    const int* levp = a.lev_slot + a.blk_lev_ptr[b];
    const int nl = a.blk_lev_ptr[b + 1] - a.blk_lev_ptr[b] - 1;
    const unsigned char* gsrc = a.stream + (size_t)ch0 * a.chunk_bytes;
    int next_issue = 0;  // next local chunk to request
    if (TMA) {
      if (lane == 0) {
        for (; next_issue < nch && next_issue < a.stages; ++next_issue) {
          const uint32_t g = base_g + next_issue, st = g % a.stages;
          mbar_expect_tx(&bars[st], a.chunk_bytes);
          bulk_g2s(ring + (size_t)st * a.chunk_stride, gsrc + (size_t)next_issue * a.chunk_bytes, a.chunk_bytes, &bars[st]);
        }
      }
      next_issue = __shfl_sync(0xffffffffu, next_issue, 0);
    }
    for (int i = lane; i < nb; i += 32) xl[i] = (CT)in[a.grow[rbase + i]];
    __syncwarp();
    int ready = 0;     // local chunks [0, ready) have landed
    int released = 0;  // local chunks [0, released) are fully consumed and their ring slots recycled
    for (int l = 0; l < nl; ++l) {
      const int s0 = levp[l], s1 = levp[l + 1];
      for (int sb = s0; sb < s1; sb += 32) {
        const int s = sb + lane;
        const int send = min(sb + 32, s1);
        if (TMA) {
          const int need = (send - 1) / CS + 1;
          for (; ready < need; ++ready) {
            const uint32_t g = base_g + ready;
            mbar_wait(&bars[g % a.stages], (g / a.stages) & 1u);
          }
        }
        if (s < s1) {
          const int c = s / CS, t = s - c * CS;
          const unsigned char* rec = TMA ? ring + (size_t)((base_g + c) % a.stages) * a.chunk_stride
                                         : gsrc + (size_t)c * a.chunk_bytes;
          const uint16_t* sc = reinterpret_cast<const uint16_t*>(rec);
          const uint16_t* cc = sc + CS;
          const FT* vv = reinterpret_cast<const FT*>(rec + (size_t)CS * (2 + 2 * K));
          const int li = sc[t];
          CT acc = xl[li];
          for (int k = 0; k < K; ++k) acc -= (CT)vv[(size_t)k * CS + t] * xl[cc[k * CS + t]];
          if (!UNIT) acc /= (CT)vv[(size_t)K * CS + t];
          xl[li] = acc;  // rows of one level never read each other (padding reads its own slot * 0)
        }
        // all lanes are done with slots < send: orders xl for the next level and lets the ring
        // slots of fully consumed chunks be refilled (a level may span more chunks than stages)
        __syncwarp();
        if (TMA) {
          const int done = send / CS;  // chunks [0,done) hold only slots < send
          if (released < done) {
            if (lane == 0) {
              for (int c = released; c < done && next_issue < nch; ++c, ++next_issue) {
                const uint32_t g = base_g + next_issue, st = g % a.stages;
                mbar_expect_tx(&bars[st], a.chunk_bytes);
                bulk_g2s(ring + (size_t)st * a.chunk_stride, gsrc + (size_t)next_issue * a.chunk_bytes, a.chunk_bytes, &bars[st]);
              }
            }
            next_issue = __shfl_sync(0xffffffffu, next_issue, 0);
            released = done;
          }
        }
      }
    }
    for (int i = lane; i < nb; i += 32) out[a.grow[rbase + i]] = (VT)xl[i];
    __syncwarp();
    if (TMA) {
      // drain: every requested chunk must have landed before the ring is reused by the next block
      for (; ready < next_issue; ++ready) {
        const uint32_t g = base_g + ready;
        mbar_wait(&bars[g % a.stages], (g / a.stages) & 1u);
      }
      base_g += (uint32_t)nch;
    }
  }
}

template <class VT, class FT, bool UNIT>
static int block_run(bg_context_t ctx, const BlockStreamDev& S, bool tma, const void* in, void* out, const int* skip) {
  using CT = decltype(VT() * FT());
  BsArgs a{};
  a.grow = S.grow.as<int>(); a.blk_row_ptr = S.blk_row_ptr.as<int>(); a.blk_lev_ptr = S.blk_lev_ptr.as<int>();
  a.lev_slot = S.lev_slot.as<int>(); a.blk_chunk_ptr = S.blk_chunk_ptr.as<int64_t>();
  a.stream = S.stream.as<unsigned char>();
  a.nblocks = S.nblocks; a.K = S.K; a.CS = S.CS; a.max_rows = S.max_rows;
  a.stages = tma ? 4 : 0;
  a.chunk_bytes = (uint32_t)S.chunk_bytes;
  a.chunk_stride = (a.chunk_bytes + 127u) & ~127u;
  const uint32_t xl_bytes = ((uint32_t)S.max_rows * sizeof(CT) + 127u) & ~127u;
  a.warp_smem = xl_bytes + (uint32_t)a.stages * a.chunk_stride + 128u;
  const int warps = 4;
  const size_t smem = (size_t)warps * a.warp_smem;
  if (smem > ctx->smem_optin) {
    set_error("block TRSV needs %zu bytes of shared memory per CTA (limit %zu): use smaller blocks", smem, ctx->smem_optin);
    return BG_ERR_UNSUPPORTED;
  }
  auto kern = tma ? trsv_block_kernel<VT, FT, CT, UNIT, true> : trsv_block_kernel<VT, FT, CT, UNIT, false>;
  static bool attr_set[2] = {false, false};  // per template instantiation and TMA variant
  if (!attr_set[tma ? 1 : 0]) {
    BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    attr_set[tma ? 1 : 0] = true;
  }
  int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (ctx->smem_optin + 1024) / (smem + 1024)));
  int blocks = std::min((S.nblocks + warps - 1) / warps, ctx->sm_count * ctas_per_sm);
  kern<<<blocks, warps * 32, smem, ctx->stream>>>(a, (const VT*)in, (VT*)out, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_trsv_block(bg_context_t ctx, const BlockStreamDev& S, int vec_dt, int fac_dt, bool unit, bool tma,
                      const void* in, void* out, const int* skip) {
  if (S.n == 0) return BG_OK;
#define BG_B(VT, FT) (unit ? block_run<VT, FT, true>(ctx, S, tma, in, out, skip) : block_run<VT, FT, false>(ctx, S, tma, in, out, skip))
  if (vec_dt == BG_F32 && fac_dt == BG_F32) return BG_B(float, float);
  if (vec_dt == BG_F32 && fac_dt == BG_F64) return BG_B(float, double);
  if (vec_dt == BG_F64 && fac_dt == BG_F32) return BG_B(double, float);
  return BG_B(double, double);
#undef BG_B
}

}  // namespace bg
