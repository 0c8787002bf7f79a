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
#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "peer.cuh"

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
// Scratch of the finalizing block: the previous rotations and the first-pass coefficients are
// fetched by many threads at once, so the serial rotation sweep of thread 0 runs out of shared
// memory instead of paying one L2 round trip per dependent global load.
constexpr int kFinMax = 32;  // restart lengths up to this use the scratch; longer ones read global memory
struct FinScratch {
  double cs[kFinMax + 1], sn[kFinMax + 1], h1[kFinMax + 1], col[kFinMax + 2];
  double s_i, normb, tol, eta2;
  int iters;
};
__device__ __forceinline__ void fin_prefetch(const GmresDev& st, FinScratch& f, int i, int tid) {
  if (st.m > kFinMax) return;
  if (tid < i) { f.cs[tid] = st.cs[tid]; f.sn[tid] = st.sn[tid]; }
  if (tid >= 32 && tid - 32 <= i) f.h1[tid - 32] = st.h[tid - 32];
  if (tid == 64) f.s_i = st.s[i];
  if (tid == 65) f.normb = st.scal[S_NORMB];
  if (tid == 66) f.tol = st.scal[S_TOL];
  if (tid == 67) f.eta2 = st.scal[S_ETA2];
  if (tid == 68) f.iters = st.flags[F_ITERS];
}
// column i of H is complete: rotate, update s, residual estimate, convergence (gmres.cu:2365-2372)
__device__ void givens_column(const GmresDev& st, int i, FinScratch* f) {
  const int ld = st.m + 1;
  double* Hc = st.H + (size_t)i * ld;
  if (f && st.m <= kFinMax) {
    // f->col holds the new column (entries 0..i+1); all operands are on chip
    double* c = f->col;
    for (int k = 0; k < i; ++k) rot_apply(c[k], c[k + 1], f->cs[k], f->sn[k]);
    double csi, sni;
    rot_gen(c[i], c[i + 1], csi, sni);
    rot_apply(c[i], c[i + 1], csi, sni);
    double si = f->s_i, si1 = 0.0;  // s[i+1] is zero before this rotation (gmres.cu:2344-2346)
    rot_apply(si, si1, csi, sni);
    for (int k = 0; k <= i + 1; ++k) Hc[k] = c[k];
    st.cs[i] = csi; st.sn[i] = sni; st.s[i] = si; st.s[i + 1] = si1;
    const double resid = fabs(si1) / f->normb;
    st.scal[S_RESID] = resid;
    const int it = f->iters;
    if (st.hist && it < st.hist_cap) st.hist[it] = resid;
    st.flags[F_ITERS] = it + 1;
    if (resid < f->tol) { st.flags[F_CONV] = 1; st.flags[F_KCONV] = i; }
    return;
  }
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

__device__ void finalize(const GmresDev& st, int fin, int i, int mgs_k, int orth_mode, const double* red, int j,
                         FinScratch* f) {
  const int ld = st.m + 1;
  const double ww = red[j];
  const bool fast = f && st.m <= kFinMax;
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
      const double eta2 = fast ? f->eta2 : st.scal[S_ETA2];
      if (orth_mode == BG_ORTH_CGS2 || !(hn2 >= eta2 * ww)) {
        st.flags[F_REORTH] = 1;   // second pass decides the column
        st.scal[S_SCALE] = 1.0;
      } else {
        st.flags[F_REORTH] = 0;
        double* Hc = fast ? f->col : st.H + (size_t)i * ld;
        for (int k = 0; k < j; ++k) Hc[k] = red[k];
        const double hn = sqrt(hn2);
        Hc[i + 1] = hn;
        if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
        givens_column(st, i, f);
      }
    } break;
    case FIN_CGS2: {
      double sum = 0.0;
      double* Hc = fast ? f->col : st.H + (size_t)i * ld;
      for (int k = 0; k < j; ++k) { st.h2[k] = red[k]; sum += red[k] * red[k]; Hc[k] = (fast ? f->h1[k] : st.h[k]) + red[k]; }
      double hn2 = ww - sum;
      if (hn2 < 0.0) hn2 = 0.0;
      const double hn = sqrt(hn2);
      Hc[i + 1] = hn;
      if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
      st.flags[F_NREORTH]++;
      givens_column(st, i, f);
    } break;
    case FIN_MGS_DOT: {  // H(k,i) = <w, v_k>  (gmres.cu:2357)
      st.H[(size_t)i * ld + mgs_k] = red[0];
      st.h[0] = red[0];
    } break;
    case FIN_MGS_NORM: {  // H(i+1,i) = ||w||  (gmres.cu:2360)
      const double hn = sqrt(ww);
      st.H[(size_t)i * ld + i + 1] = hn;
      if (hn > 0.0) st.scal[S_SCALE] = 1.0 / hn; else { st.scal[S_SCALE] = 0.0; st.flags[F_BREAK] = 1; }
      givens_column(st, i, nullptr);  // the MGS dots wrote H(0..i,i) to global memory one by one
    } break;
    default: break;
  }
}

// ---------------------------------------------------------------------------- TMA pipeline
// Both hot kernels stream a tall-skinny tile of the Krylov basis (R rows x all active columns)
// through a ring of shared-memory stages filled by the TMA unit (cp.async.bulk, one bulk copy per
// column segment, completion counted on an mbarrier), so the bytes in flight per SM are set by
// the ring (~190 KB) and not by how many loads a thread can keep in registers.  One persistent
// CTA per SM: warp 8 produces, warps 0-7 consume.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// bounded wait: a pipeline bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kConsumerWarps = 8;
constexpr int kPipeThreads = (kConsumerWarps + 1) * 32;

struct Pipe {          // geometry chosen on the host per launch
  int R;               // rows per stage (multiple of 32)
  int stages;
  uint32_t stage_bytes;
};

// producer warp: tiles t = first, first+step, ... ; column c < ncols of a tile comes from
// (c < j ? V + c*ld : extra) + t*R.  `reverse` walks the tiles from the last one down.
template <class VT>
__device__ __forceinline__ void pipe_produce(const Pipe& pp, unsigned char* ring, uint64_t* full, uint64_t* empty, int n,
                                             const VT* V, int64_t ld, int j, const VT* extra, bool reverse) {
  const int lane = threadIdx.x & 31;
  const int ntiles = (n + pp.R - 1) / pp.R;
  const int ncols = j + (extra ? 1 : 0);
  int it = 0;
  for (int t0 = blockIdx.x; t0 < ntiles; t0 += gridDim.x, ++it) {
    const int t = reverse ? ntiles - 1 - t0 : t0;
    const int s = it % pp.stages;
    const uint32_t use = (uint32_t)(it / pp.stages);
    if (use > 0) mbar_wait(&empty[s], (use & 1u) ^ 1u);
    const int rows = min(pp.R, n - t * pp.R);
    const uint32_t col_bytes = ((uint32_t)rows * sizeof(VT) + 15u) & ~15u;  // V columns and work vectors are padded to ld
    if (lane == 0) mbar_expect_tx(&full[s], col_bytes * (uint32_t)ncols);
    __syncwarp();
    unsigned char* tile = ring + (size_t)s * pp.stage_bytes;
    for (int c = lane; c < ncols; c += 32) {
      const VT* src = (c < j ? V + (size_t)c * ld : extra) + (size_t)t * pp.R;
      bulk_g2s(tile + (size_t)c * pp.R * sizeof(VT), src, col_bytes, &full[s]);
    }
  }
}

// ---------------------------------------------------------------------------- peer memory (multi-GPU)
// red[0..cnt) <- sum over the ranks, in rank order (identical bits everywhere).  Called by ALL threads
// of one CTA per rank.  Slots alternate with the parity of the sequence number: a rank can run at most
// one all-reduce ahead of the slowest one, because finishing number s needs everybody's values of s.
__device__ void peer_allreduce(const PeerAR* __restrict__ P, double* red, int cnt) {
  __shared__ unsigned s_seq;
  __shared__ double s_val[kMaxPeers * (kFinMax + 2)];
  const int R = P->nranks, me = P->rank;
  if (threadIdx.x == 0) { s_seq = *P->seq + 1u; *P->seq = s_seq; }
  __syncthreads();
  const unsigned seq = s_seq;
  const size_t slot = (size_t)(seq & 1u) * R;
  for (int t = threadIdx.x; t < R * cnt; t += blockDim.x) {
    const int r = t / cnt, k = t - r * cnt;
    ll_store(&P->inbox[r][(slot + me) * kArLd + k], red[k], seq);
  }
  // one thread per (rank, value): all packets are polled at once -- a thread walking the ranks of its value one after
  // the other paid one memory round trip per rank (5 us of the 13 us all-reduce at 8 ranks)
  for (int t = threadIdx.x; t < R * cnt; t += blockDim.x) {
    const int r = t / cnt, k = t - r * cnt;
    s_val[r * (kFinMax + 2) + k] = ll_wait(&P->inbox[me][(slot + r) * kArLd + k], seq, P->timeout_ns);
  }
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    double v = 0.0;
    for (int r = 0; r < R; ++r) v += s_val[r * (kFinMax + 2) + threadIdx.x];  // rank order: identical bits on every rank
    red[threadIdx.x] = v;
  }
  __syncthreads();
}

// One halo exchange in one kernel and without any fence: every boundary entry travels as a self-validating
// 16-byte element into the neighbour's inbox; the same kernel then polls its own inbox and writes the
// arrived values into the halo slots behind the local vector, so the SpMV that follows reads plain memory.
template <class VT>
__global__ void __launch_bounds__(256) halo_exchange_kernel(const PeerHalo* __restrict__ H, const int* __restrict__ idx,
                                                             const VT* __restrict__ v, VT* __restrict__ halo_out, unsigned seq,
                                                             const int* skip) {
  if (skip && *skip) return;
  const int nn = H->n_nbr, total = H->send_ptr[nn], nh = H->n_halo;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int e = tid; e < total; e += nt) {
    int k = 0;
    while (e >= H->send_ptr[k + 1]) ++k;
    ll_store(H->peer_inbox[k] + (size_t)(seq & 1u) * H->peer_half[k] + (e - H->send_ptr[k]), (double)v[idx[e]], seq);
  }
  const LL16* in = H->my_inbox + (size_t)(seq & 1u) * max(nh, 1);
  for (int i = tid; i < nh; i += nt) halo_out[i] = (VT)ll_wait(in + i, seq, H->timeout_ns);
}

// ---------------------------------------------------------------------------- dots
// partial[k] = V[:,k].w (k<j), partial[j] = w.w ; the last CTA reduces and runs finalize().
template <class VT>
__global__ void __launch_bounds__(kPipeThreads, 1) dots_kernel(int n, const VT* __restrict__ V, int64_t ld, int j,
                                                                const VT* __restrict__ w, GmresDev st, int fin,
                                                                int inner_i, int mgs_k, int orth_mode, int need_reorth,
                                                                Pipe pp, double* red_out, const PeerAR* peer) {
  if (st.flags[F_CONV]) return;
  if (need_reorth && !st.flags[F_REORTH]) return;
  extern __shared__ __align__(128) unsigned char dyn[];
  unsigned char* ring = dyn;
  uint64_t* full = reinterpret_cast<uint64_t*>(dyn + (size_t)pp.stages * pp.stage_bytes);
  uint64_t* empty = full + pp.stages;
  __shared__ double red_w[kConsumerWarps][5];
  __shared__ double red[kFinMax + 2];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < pp.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ncols = j + 1;
  // column groups x row groups over the 8 consumer warps; at most 5 columns per warp
  const int CW = ncols > 20 ? 8 : ncols > 10 ? 4 : ncols > 5 ? 2 : 1, RW = kConsumerWarps / CW;
  if (wid == kConsumerWarps) {
    pipe_produce<VT>(pp, ring, full, empty, n, V, ld, j, w, false);
  } else {
    const int cg = wid % CW, rg = wid / CW;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int ntiles = (n + pp.R - 1) / pp.R;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % pp.stages;
      mbar_wait(&full[s], (uint32_t)(it / pp.stages) & 1u);
      const VT* tile = reinterpret_cast<const VT*>(ring + (size_t)s * pp.stage_bytes);
      const VT* tw = tile + (size_t)j * pp.R;
      const int rows = min(pp.R, n - t * pp.R);
#pragma unroll 4
      for (int r = lane + 32 * rg; r < rows; r += 32 * RW) {
        const double wv = (double)tw[r];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const int c = cg + a * CW;
          if (c < ncols) acc[a] += (double)tile[(size_t)c * pp.R + r] * wv;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const double v = warp_sum(acc[a]);
      if (lane == 0) red_w[wid][a] = v;
    }
  }
  __syncthreads();
  // slot k of the partial row: column k = cg + a*CW, summed over the row groups
  if (threadIdx.x <= j) {
    const int cg = threadIdx.x % CW, a = threadIdx.x / CW;
    double v = 0.0;
    for (int rg = 0; rg < RW; ++rg) v += red_w[cg + CW * rg][a];
    st.partial[(size_t)blockIdx.x * st.partial_ld + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // deterministic final reduction: warp q owns slots q, q+9, q+18, q+27; lanes stride over CTAs.
  // All loads of a warp are issued before the first use (one L2 round trip instead of four).
  {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = wid + u * (kConsumerWarps + 1);
      if (k <= j)
        for (int b = lane; b < (int)gridDim.x; b += 32) v[u] += __ldcg(&st.partial[(size_t)b * st.partial_ld + k]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = wid + u * (kConsumerWarps + 1);
      const double r = warp_sum(v[u]);
      if (k <= j && lane == 0) red[k] = r;
    }
  }
  if (red_out) {
    // row-partitioned run: these are one rank's sums; the host allreduces red_out over the ranks
    // (NCCL) and finalize_kernel runs the scalar epilogue on the global values
    __syncthreads();
    if (threadIdx.x <= j) red_out[threadIdx.x] = red[threadIdx.x];
    if (threadIdx.x == 0) *st.ticket = 0u;
    return;
  }
  __shared__ FinScratch fs;
  const bool column_fin = fin == FIN_CGS1 || fin == FIN_CGS2;
  if (column_fin) fin_prefetch(st, fs, inner_i, threadIdx.x);
  if (peer) {
    // row-partitioned run: sum the per-rank values over the GPUs of the box through peer memory,
    // inside this same CTA (no NCCL call, no extra launch); the scalar epilogue then runs on
    // rank-identical numbers
    __syncthreads();
    peer_allreduce(peer, red, j + 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    finalize(st, fin, inner_i, mgs_k, orth_mode, red, j, column_fin ? &fs : nullptr);
    *st.ticket = 0u;
  }
}

// scalar epilogue of a row-partitioned dots step, after the allreduce of the per-rank sums
__global__ void finalize_kernel(GmresDev st, int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth, int j,
                                const double* __restrict__ red_in) {
  if (st.flags[F_CONV]) return;
  if (need_reorth && !st.flags[F_REORTH]) return;
  __shared__ double red[kFinMax + 2];
  __shared__ FinScratch fs;
  if (threadIdx.x <= j) red[threadIdx.x] = red_in[threadIdx.x];
  const bool column_fin = fin == FIN_CGS1 || fin == FIN_CGS2;
  if (column_fin) fin_prefetch(st, fs, inner_i, threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) finalize(st, fin, inner_i, mgs_k, orth_mode, red, j, column_fin ? &fs : nullptr);
}

// boundary values of a vector packed for the halo exchange: buf[k] = v[idx[k]]
template <class VT>
__global__ void pack_kernel(int cnt, const int* __restrict__ idx, const VT* __restrict__ v, VT* __restrict__ buf, const int* skip) {
  if (skip && *skip) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) buf[k] = v[idx[k]];
}

// ---------------------------------------------------------------------------- update
// out = (a + sgn * V[:,0..j) c) * scale, j >= 1 (the MGS/CGS axpy sweep fused with the
// normalisation cublasScopy+cublasSscal of the reference, and Update_GPU's x += V y).
template <class VT>
__global__ void __launch_bounds__(kPipeThreads, 1) update_kernel(int n, const VT* a, const VT* __restrict__ V, int64_t ld,
                                                                  int j, const double* __restrict__ coef, double sgn,
                                                                  VT* out, GmresDev st, int use_scale, int force,
                                                                  int need_reorth, Pipe pp) {
  if (!force && st.flags[F_CONV]) return;
  if (need_reorth && !st.flags[F_REORTH]) return;
  extern __shared__ __align__(128) unsigned char dyn[];
  unsigned char* ring = dyn;
  uint64_t* full = reinterpret_cast<uint64_t*>(dyn + (size_t)pp.stages * pp.stage_bytes);
  uint64_t* empty = full + pp.stages;
  __shared__ double cs[kFinMax];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < pp.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < j) cs[threadIdx.x] = sgn * coef[threadIdx.x];
  __syncthreads();
  if (wid == kConsumerWarps) {
    // rows are walked from the END: the dots kernel that ran just before finished on the high rows
    // of every basis column, so those are the lines still resident in the 126 MB L2
    pipe_produce<VT>(pp, ring, full, empty, n, V, ld, j, (const VT*)nullptr, true);
    return;
  }
  const double scale = use_scale ? st.scal[S_SCALE] : 1.0;
  const int ntiles = (n + pp.R - 1) / pp.R;
  int it = 0;
  for (int t0 = blockIdx.x; t0 < ntiles; t0 += gridDim.x, ++it) {
    const int t = ntiles - 1 - t0;
    const int s = it % pp.stages;
    const int rows = min(pp.R, n - t * pp.R);
    const size_t base = (size_t)t * pp.R;
    // the addend comes straight from global memory (it may be a caller-owned, unpadded vector);
    // its loads are issued before the wait so they overlap the bulk copies
    double av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = threadIdx.x + 256 * u;
      av[u] = r < rows ? (double)a[base + r] : 0.0;
    }
    mbar_wait(&full[s], (uint32_t)(it / pp.stages) & 1u);
    const VT* tile = reinterpret_cast<const VT*>(ring + (size_t)s * pp.stage_bytes);
    for (int r0 = 0; r0 < rows; r0 += 1024) {
      if (r0 > 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + threadIdx.x + 256 * u;
          av[u] = r < rows ? (double)a[base + r] : 0.0;
        }
      }
      for (int k = 0; k < j; ++k) {
        const double c = cs[k];
        const VT* col = tile + (size_t)k * pp.R + r0 + threadIdx.x;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (r0 + (int)threadIdx.x + 256 * u < rows) av[u] += c * (double)col[256 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + threadIdx.x + 256 * u;
        if (r < rows) out[base + r] = (VT)(av[u] * scale);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

// ---------------------------------------------------------------------------- fused orthogonalisation step
// One cooperative launch per inner iteration replaces dots_kernel + update_kernel (+ the two launches of
// the conditional second pass): the persistent CTAs stream their tiles of [V(:,0..j) | w] once for the
// projections, the last CTA to arrive reduces (all-reducing over the GPUs of the box through peer memory
// in a row-partitioned run), runs the scalar epilogue and publishes a generation number; every CTA then
// walks ITS OWN tiles again in reverse order for v_j = (w - V h) * scale.  The tiles still sitting in the
// shared-memory ring when the second phase starts are used in place ("virtual" fills: the producer only
// completes the barrier), the rest mostly hits the L2 lines the first phase left behind.  When the DGKS
// test fires (or under CGS2) the same kernel runs the second pass on its own output.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx0(uint64_t* bar) {  // completes a phase that expects no bytes
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(0u) : "memory");
}

// BT = storage type of the basis columns (V and the copy of the new vector in out_b), VT = type of w and of the working
// copy `out` of the new vector.  BT == VT: out IS the basis column and out_b is null.  BT = float with VT = double is the
// opt-in mixed-precision basis (bg_solver_create_ex): the j projections stream 4-byte columns, every sum, w and the
// new vector stay FP64, and the new vector is rounded to BT before it is stored in EITHER place, so that the vector the
// next SpMV multiplies is exactly the one the basis holds (the Arnoldi relation then holds for the stored basis).
// float -> double by re-packing the bits (normal numbers exactly; zero and denormals become ~2^-127, which is below
// anything the sums can resolve; the basis holds no infinities or NaNs while the solve is healthy)
__device__ __forceinline__ double widen_f32(float x) {
  const unsigned f = __float_as_uint(x);
  const unsigned hi = (f & 0x80000000u) | (((f & 0x7fffffffu) >> 3) + 0x38000000u);
  return __hiloint2double((int)hi, (int)(f << 29));
}
template <class VT, class BT>
__global__ void __launch_bounds__(kPipeThreads, 1) orth_kernel(int n, const BT* __restrict__ V, int64_t ld, int j,
                                                                const VT* w, VT* out, BT* out_b, GmresDev st, int inner_i,
                                                                int orth_mode, Pipe pp, const PeerAR* peer) {
  if (st.flags[F_CONV]) return;
  extern __shared__ __align__(128) unsigned char dyn[];
  unsigned char* ring = dyn;
  uint64_t* full = reinterpret_cast<uint64_t*>(dyn + (size_t)pp.stages * pp.stage_bytes);
  uint64_t* empty = full + pp.stages;
  __shared__ double red_w[kConsumerWarps][5];
  __shared__ double red[kFinMax + 2];
  __shared__ double cs[kFinMax];
  __shared__ FinScratch fs;
  __shared__ bool last;
  __shared__ int s_ctl[2];
  __shared__ double s_scale;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned* gen = st.ticket + 1;
  const unsigned gen0 = ld_acquire_gpu(gen);  // read by every CTA before it takes a ticket
  if (threadIdx.x == 0) {
    for (int s = 0; s < pp.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int S = pp.stages, G = gridDim.x;
  const int ntiles = (n + pp.R - 1) / pp.R;
  const int K = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + G - 1) / G : 0;  // tiles of this CTA
  const int ncols = j + 1;
  const int CW = ncols > 20 ? 8 : ncols > 10 ? 4 : ncols > 5 ? 2 : 1, RW = kConsumerWarps / CW;
  // per-stage parities of the fills (producer) / uses (consumers) so far: stages are not visited round-robin
  unsigned fbits = 0, fany = 0, ubits = 0;

  for (int pass = 0; pass < 2; ++pass) {
    const VT* extra = pass == 0 ? w : out;
    // ---------------- phase A: projections of `extra` on V(:,0..j) and its squared norm
    if (wid == kConsumerWarps) {
      for (int q = 0; q < K; ++q) {
        const int t = blockIdx.x + q * G, s = q % S;
        if ((fany >> s) & 1u) mbar_wait(&empty[s], ((fbits >> s) & 1u) ^ 1u);
        const int rows = min(pp.R, n - t * pp.R);
        const uint32_t cb_v = ((uint32_t)rows * sizeof(BT) + 15u) & ~15u, cb_w = ((uint32_t)rows * sizeof(VT) + 15u) & ~15u;
        if (lane == 0) mbar_expect_tx(&full[s], cb_v * (uint32_t)j + cb_w);
        __syncwarp();
        unsigned char* tile = ring + (size_t)s * pp.stage_bytes;
        for (int c = lane; c < ncols; c += 32) {
          if (c < j) bulk_g2s(tile + (size_t)c * pp.R * sizeof(BT), V + (size_t)c * ld + (size_t)t * pp.R, cb_v, &full[s]);
          else bulk_g2s(tile + (size_t)j * pp.R * sizeof(BT), extra + (size_t)t * pp.R, cb_w, &full[s]);
        }
        fbits ^= 1u << s; fany |= 1u << s;
      }
    } else {
      const int cg = wid % CW, rg = wid / CW;
      double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      for (int q = 0; q < K; ++q) {
        const int t = blockIdx.x + q * G, s = q % S;
        mbar_wait(&full[s], (ubits >> s) & 1u);
        const BT* tile = reinterpret_cast<const BT*>(ring + (size_t)s * pp.stage_bytes);
        const VT* tw = reinterpret_cast<const VT*>(ring + (size_t)s * pp.stage_bytes + (size_t)j * pp.R * sizeof(BT));
        const int rows = min(pp.R, n - t * pp.R);
        if constexpr (sizeof(BT) == 4) {
          // FP32 basis columns (mixed basis, and the all-FP32 solver): FP32 products summed in FP32 over this tile's rows
          // (<= 128 per thread), folded into the FP64 accumulators once per tile; w.w stays FP64.  A float -> double
          // conversion per ELEMENT (F2F.F64.F32) would make the step slower than the all-FP64 one; the rounding errors
          // of the products average out over the rows (the reference's cublasSdot sums everything in FP32).
          float a32[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
          for (int r = lane + 32 * rg; r < rows; r += 32 * RW) {
            const double wv = (double)tw[r];
            const float wf = (float)wv;
#pragma unroll
            for (int a = 0; a < 5; ++a) {
              const int c = cg + a * CW;
              if (c < j) a32[a] = fmaf((float)tile[(size_t)c * pp.R + r], wf, a32[a]);
              else if (c == j) acc[a] += wv * wv;
            }
          }
#pragma unroll
          for (int a = 0; a < 5; ++a) acc[a] += (double)a32[a];
        } else {  // uniform tile: column j is w itself
#pragma unroll 4
          for (int r = lane + 32 * rg; r < rows; r += 32 * RW) {
            const double wv = (double)tw[r];
#pragma unroll
            for (int a = 0; a < 5; ++a) {
              const int c = cg + a * CW;
              if (c < ncols) acc[a] += (double)tile[(size_t)c * pp.R + r] * wv;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        ubits ^= 1u << s;
      }
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        const double v = warp_sum(acc[a]);
        if (lane == 0) red_w[wid][a] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x <= j) {
      const int cg = threadIdx.x % CW, a = threadIdx.x / CW;
      double v = 0.0;
      for (int rg = 0; rg < RW; ++rg) v += red_w[cg + CW * rg][a];
      st.partial[(size_t)blockIdx.x * st.partial_ld + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    const unsigned gen_next = gen0 + 1u + (unsigned)pass;
    if (last) {
      __threadfence();
      double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = wid + u * (kConsumerWarps + 1);
        if (k <= j)
          for (int b = lane; b < (int)gridDim.x; b += 32) v[u] += __ldcg(&st.partial[(size_t)b * st.partial_ld + k]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = wid + u * (kConsumerWarps + 1);
        const double r = warp_sum(v[u]);
        if (k <= j && lane == 0) red[k] = r;
      }
      fin_prefetch(st, fs, inner_i, threadIdx.x);
      __syncthreads();
      if (peer) peer_allreduce(peer, red, j + 1);
      if (threadIdx.x == 0) {
        finalize(st, pass == 0 ? FIN_CGS1 : FIN_CGS2, inner_i, 0, orth_mode, red, j, &fs);
        *st.ticket = 0u;
        __threadfence();
        st_release_gpu(gen, gen_next);
      }
    }
    // ---------------- everybody waits for the coefficients of this pass
    if (threadIdx.x == 0) {
      for (unsigned spin = 0; ld_acquire_gpu(gen) != gen_next; ++spin)
        if (spin > (1u << 28)) __trap();
      s_ctl[0] = __ldcg(&st.flags[F_CONV]);
      s_ctl[1] = __ldcg(&st.flags[F_REORTH]);
      s_scale = __ldcg(&st.scal[S_SCALE]);
    }
    __syncthreads();
    if (s_ctl[0]) return;  // converged in this iteration: v_j is not needed (the separate kernels skip it too)
    const bool again = pass == 0 && s_ctl[1] != 0;
    if (threadIdx.x < j) cs[threadIdx.x] = -__ldcg(&(pass == 0 ? st.h : st.h2)[threadIdx.x]);
    __syncthreads();
    // ---------------- phase B: out = (extra - V c) * scale over the same tiles, last tile first
    if (wid == kConsumerWarps) {
      for (int q = 0; q < K; ++q) {
        const int qq = K - 1 - q, t = blockIdx.x + qq * G, s = qq % S;
        if ((fany >> s) & 1u) mbar_wait(&empty[s], ((fbits >> s) & 1u) ^ 1u);
        if (q < S) {
          // still resident from phase A: complete the barrier without moving a byte
          if (lane == 0) mbar_arrive_tx0(&full[s]);
        } else {
          const int rows = min(pp.R, n - t * pp.R);
          const uint32_t cb_v = ((uint32_t)rows * sizeof(BT) + 15u) & ~15u, cb_w = ((uint32_t)rows * sizeof(VT) + 15u) & ~15u;
          if (lane == 0) mbar_expect_tx(&full[s], cb_v * (uint32_t)j + cb_w);
          __syncwarp();
          unsigned char* tile = ring + (size_t)s * pp.stage_bytes;
          for (int c = lane; c < ncols; c += 32) {
            if (c < j) bulk_g2s(tile + (size_t)c * pp.R * sizeof(BT), V + (size_t)c * ld + (size_t)t * pp.R, cb_v, &full[s]);
            else bulk_g2s(tile + (size_t)j * pp.R * sizeof(BT), extra + (size_t)t * pp.R, cb_w, &full[s]);
          }
        }
        fbits ^= 1u << s; fany |= 1u << s;
      }
    } else {
      const double scale = s_scale;
      for (int q = 0; q < K; ++q) {
        const int qq = K - 1 - q, t = blockIdx.x + qq * G, s = qq % S;
        mbar_wait(&full[s], (ubits >> s) & 1u);
        const BT* tile = reinterpret_cast<const BT*>(ring + (size_t)s * pp.stage_bytes);
        const VT* tw = reinterpret_cast<const VT*>(ring + (size_t)s * pp.stage_bytes + (size_t)j * pp.R * sizeof(BT));
        const int rows = min(pp.R, n - t * pp.R);
        const size_t base = (size_t)t * pp.R;
        for (int r0 = 0; r0 < rows; r0 += 1024) {
          double av[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + threadIdx.x + 256 * u;
            av[u] = r < rows ? (double)tw[r] : 0.0;
          }
          if constexpr (sizeof(BT) == 4) {
            // FP32 basis columns: w - V h cancels down to the (often much smaller) new direction, so the products and the sum
            // must be exact to FP64 relative to w: FP64 FMAs on widened columns.  The widening is done with integer
            // instructions (widen_f32): the conversion instruction F2F.F64.F32 is issued at a fraction of the FMA rate
            // and made this step slower than the all-FP64 one (377 vs 360 us at 10 M rows).
            for (int k = 0; k < j; ++k) {
              const double c = cs[k];
              const BT* col = tile + (size_t)k * pp.R + r0 + threadIdx.x;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (r0 + (int)threadIdx.x + 256 * u < rows) av[u] += c * widen_f32((float)col[256 * u]);
            }
          } else {
            for (int k = 0; k < j; ++k) {
              const double c = cs[k];
              const BT* col = tile + (size_t)k * pp.R + r0 + threadIdx.x;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (r0 + (int)threadIdx.x + 256 * u < rows) av[u] += c * (double)col[256 * u];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + threadIdx.x + 256 * u;
            if (r < rows) {
              if constexpr (sizeof(BT) == sizeof(VT)) {
                out[base + r] = (VT)(av[u] * scale);
              } else {  // rounded to the basis precision in both copies
                const BT vb = (BT)(av[u] * scale);
                out[base + r] = (VT)vb;
                out_b[base + r] = vb;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        ubits ^= 1u << s;
      }
    }
    if (!again) return;
    // second pass reads `out` through the bulk-copy (async) proxy: order the generic stores before it
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- fused strict MGS step
// The reference's orthogonalisation (src/gmres.cu:2356-2362): for k = 0..i { H(k,i) = <w, v_k>; w -= H(k,i) v_k },
// H(i+1,i) = ||w||, v_{i+1} = w / H(i+1,i) -- i+1 DEPENDENT dot/axpy pairs.  As separate kernels that is 2(i+2)
// launches per inner iteration, each a few microseconds of work at power-grid sizes (337 us per iteration on the
// 1.7 M-row grid, five times the fused CGS step).  Here the whole step is ONE cooperative launch:
//  * CTA c owns rows [c*rows_per, (c+1)*rows_per); its slice of w lives in REGISTERS (1024 threads x 12 values)
//    for the whole step and is written once, as v_{i+1};
//  * the slices of v_k arrive by bulk-async copies into two shared-memory buffers, v_{k+2} being fetched while pair
//    k+1 is computed, so every basis column is read from HBM exactly once and the copy hides behind the exchange;
//  * the grid-wide sum of a pair is an all-to-all of flag-in-data packets: every CTA stores its partial sum (value +
//    launch sequence number in one 16-byte store) and polls the packets of all CTAs -- one L2 round trip after the
//    last store instead of an atomic ticket plus a broadcast -- and adds them in CTA order, so every CTA holds the same
//    bits and the result does not depend on timing;
//  * CTA 0 keeps the column of H and runs the Givens / residual / convergence epilogue.
// Sums are accumulated in double whatever VT is; for VT = float w is rounded to float after every axpy, as the
// separate kernels (and the reference's float w) do.
constexpr int kMgsThreads = 1024, kMgsEPT = 12;
#ifdef BG_MGS_TIMING
// development aid (make dbg): clock64 stamps (SM cycles) of thread 0 of CTA g_mgs_probe, per pair k < 40:
// [0] tile there  [1] local sum done  [2] own packet stored  [3] all packets read  [4] pair done
__device__ unsigned long long g_mgs_t[40 * 5];
__device__ int g_mgs_probe = 0;
// stamps go to shared memory and are flushed at the end of the kernel: a global store in front of the packet store
// delayed the latter by a store round trip
#define MGS_STAMP(k, w) do { if (threadIdx.x == 0 && (k) < 40) s_mgs_t[(k) * 5 + (w)] = (unsigned long long)clock64(); } while (0)  // SM clock: one CTA, one SM
#else
#define MGS_STAMP(k, w) do {} while (0)
#endif
struct MgsArgs {
  LL16* slots;   // [(m + 2)][G] packets
  unsigned seq;  // launch sequence number (never 0)
  int rows_per;  // rows per CTA (multiple of 32)
  unsigned long long timeout_ns;
};
template <class VT>
__global__ void __launch_bounds__(kMgsThreads, 1) mgs_kernel(int n, const VT* __restrict__ V, int64_t ld, int i, const VT* __restrict__ w_in,
                                                             VT* __restrict__ out, GmresDev st, MgsArgs a) {
  if (st.flags[F_CONV]) return;  // written by an earlier launch: uniform over the grid
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t full[2];
  __shared__ double red_w[32];
  __shared__ double s_sum;
  __shared__ FinScratch fs;
#ifdef BG_MGS_TIMING
  __shared__ unsigned long long s_mgs_t[40 * 5];
#endif
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x;
  const int r0 = blockIdx.x * a.rows_per, rows = max(0, min(a.rows_per, n - r0));
  const uint32_t buf_bytes = ((uint32_t)a.rows_per * sizeof(VT) + 127u) & ~127u;
  const uint32_t bytes = ((uint32_t)rows * sizeof(VT) + 15u) & ~15u;  // columns are padded to ld: the tail stays in bounds
  const int ncol = i + 1;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {  // one thread
    if (rows > 0) {
      mbar_expect_tx(&full[k & 1], bytes);
      bulk_g2s(dyn + (size_t)(k & 1) * buf_bytes, V + (size_t)k * ld + r0, bytes, &full[k & 1]);
    }
  };
  if (tid == 0) { issue(0); if (ncol > 1) issue(1); }
  double wr[kMgsEPT];
#pragma unroll
  for (int e = 0; e < kMgsEPT; ++e) {
    const int r = tid + e * kMgsThreads;
    wr[e] = r < rows ? (double)w_in[r0 + r] : 0.0;
  }
  if (blockIdx.x == 0) fin_prefetch(st, fs, i, tid);
  const bool fast = st.m <= kFinMax;

  // grid-wide sum of one value per thread; identical bits in every thread of every CTA
  auto grid_sum = [&](double p, int k) -> double {
    p = warp_sum(p);
    if (lane == 0) red_w[wid] = p;
    __syncthreads();
    if (wid == 0) {
      const double mine = warp_sum(red_w[lane]);
      MGS_STAMP(k, 1);  // (after the first use of data guarded by the barrier: BAR.SYNC.DEFER_BLOCKING lets the warp run on)
      LL16* row = a.slots + (size_t)k * G;
      if (lane == 0) ll_store_gpu(row + blockIdx.x, mine, a.seq);
      MGS_STAMP(k, 2);
      constexpr int PER = 5;  // up to 160 CTAs (launcher)
      LL16 pk[PER];
      bool need[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) { need[u] = lane + 32 * u < G; pk[u] = LL16{0u, 0u, 0u, 0u}; }
      unsigned long long t0 = 0;
      for (unsigned spin = 0;; ++spin) {  // warp-uniform poll of everybody's packet
        bool all = true;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          if (need[u]) {
            const LL16* q = row + lane + 32 * u;
            pk[u] = ll_load_gpu(q);
            need[u] = !(pk[u].f0 == a.seq && pk[u].f1 == a.seq);
          }
          all = all && !need[u];
        }
        if (__all_sync(0xffffffffu, all)) break;
        if ((spin & 4095u) == 4095u) {
          const unsigned long long now = global_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > a.timeout_ns) __trap();
        }
      }
      MGS_STAMP(k, 3);
      double acc = 0.0;
#pragma unroll
      for (int u = 0; u < PER; ++u)
        if (lane + 32 * u < G) acc += __hiloint2double((int)pk[u].hi, (int)pk[u].lo);
      acc = warp_sum(acc);
      if (lane == 0) s_sum = acc;
    }
    __syncthreads();
    return s_sum;
  };

  for (int k = 0; k < ncol; ++k) {
    // one lane per warp waits for the tile: 1024 threads polling one mbarrier serialise on it (the probe showed the
    // CTA spending ~3000 cycles per pair between the tile's arrival and the end of the local dot products)
    if (rows > 0) {
      if (lane == 0) mbar_wait(&full[k & 1], (uint32_t)(k >> 1) & 1u);
      __syncwarp();
    }
    MGS_STAMP(k, 0);
    const VT* vk = reinterpret_cast<const VT*>(dyn + (size_t)(k & 1) * buf_bytes);
    double p = 0.0;
#pragma unroll
    for (int e = 0; e < kMgsEPT; ++e) {
      const int r = tid + e * kMgsThreads;
      if (r < rows) p += (double)vk[r] * wr[e];
    }
    const double h = grid_sum(p, k);  // H(k,i) = <w, v_k>  (gmres.cu:2357)
    if (blockIdx.x == 0 && tid == 0) {
      if (fast) fs.col[k] = h; else st.H[(size_t)i * (st.m + 1) + k] = h;
    }
#pragma unroll
    for (int e = 0; e < kMgsEPT; ++e) {
      const int r = tid + e * kMgsThreads;
      if (r < rows) wr[e] = (double)(VT)(wr[e] - h * (double)vk[r]);  // w -= H(k,i) v_k  (gmres.cu:2358)
    }
    __syncthreads();  // everybody is done with this buffer
    if (tid == 0 && k + 2 < ncol) issue(k + 2);
    MGS_STAMP(k, 4);
  }
  double p = 0.0;
#pragma unroll
  for (int e = 0; e < kMgsEPT; ++e) p += wr[e] * wr[e];
  const double ww = grid_sum(p, ncol);
  const double hn = sqrt(ww);  // H(i+1,i) = ||w||  (gmres.cu:2360)
  const double scale = hn > 0.0 ? 1.0 / hn : 0.0;
#pragma unroll
  for (int e = 0; e < kMgsEPT; ++e) {
    const int r = tid + e * kMgsThreads;
    if (r < rows) out[r0 + r] = (VT)(wr[e] * scale);  // v_{i+1} = w / H(i+1,i)  (gmres.cu:2361-2362)
  }
#ifdef BG_MGS_TIMING
  if ((int)blockIdx.x == g_mgs_probe && tid == 0)
    for (int q = 0; q < 200; ++q) g_mgs_t[q] = s_mgs_t[q];
#endif
  if (blockIdx.x == 0 && tid == 0) {
    st.scal[S_SCALE] = scale;
    if (!(hn > 0.0)) st.flags[F_BREAK] = 1;
    if (fast) {
      fs.col[i + 1] = hn;
      givens_column(st, i, &fs);
    } else {
      st.H[(size_t)i * (st.m + 1) + i + 1] = hn;
      givens_column(st, i, nullptr);
    }
  }
}

// out = a * scale (no basis columns: v0 = r/beta, gmres.cu:2342-2343)
template <class VT>
__global__ void __launch_bounds__(256) scale_copy_kernel(int n, const VT* a, VT* out, GmresDev st, int use_scale, int force) {
  if (!force && st.flags[F_CONV]) return;
  const double scale = use_scale ? st.scal[S_SCALE] : 1.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (VT)((double)a[i] * scale);
}

// back-substitution H y = s on one thread (Update, src/gmres.cu:92-105 / Update_GPU :172-181)
__global__ void solve_y_kernel(GmresDev st, int k, int force) {
  if (!force && st.flags[F_CONV]) return;
  const int ld = st.m + 1;
  if (st.m <= kFinMax) {  // whole triangle on chip: parallel fetch, serial solve out of shared memory
    __shared__ double Hs[(kFinMax + 1) * kFinMax], ys[kFinMax + 1];
    for (int t = threadIdx.x; t < ld * (k + 1); t += blockDim.x) Hs[t] = st.H[t];
    for (int t = threadIdx.x; t <= k; t += blockDim.x) ys[t] = st.s[t];
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = k; i >= 0; --i) {
        ys[i] /= Hs[i * ld + i];
        for (int q = i - 1; q >= 0; --q) ys[q] -= Hs[i * ld + q] * ys[i];
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t <= k; t += blockDim.x) st.y[t] = ys[t];
    return;
  }
  if (threadIdx.x != 0) return;
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

// block-major renumbering of the solver's vectors (trsv_block.cu): out[i] = in[perm[i]] / out[perm[i]] = in[i]
template <class VT, class ST>
__global__ void gather_kernel(int n, const int* __restrict__ perm, const VT* __restrict__ in, VT* __restrict__ out,
                              const ST* __restrict__ scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = scale ? (VT)((double)in[perm[i]] * (double)scale[i]) : in[perm[i]];
}
template <class VT>
__global__ void scatter_kernel(int n, const int* __restrict__ perm, const VT* __restrict__ in, VT* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[perm[i]] = in[i];
}

// ---------------------------------------------------------------------------- launchers
static int vec_grid(bg_context_t ctx, int64_t n, int per_thread, int ctas_per_sm = 4) {
  int64_t blocks = (n / per_thread + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ring geometry: `stages` stages of about `stage_kb` KB each, R rows x ncols columns per stage
static Pipe pipe_geometry(int ncols, size_t elem) {
  static const int stage_kb = [] { const char* e = getenv("BG_ORTH_STAGE_KB"); int v = e ? atoi(e) : 0; return v >= 4 && v <= 104 ? v : 100; }();
  static const int stages = [] { const char* e = getenv("BG_ORTH_STAGES"); int v = e ? atoi(e) : 0; return v >= 2 && v <= 8 ? v : 2; }();
  Pipe p;
  int R = (int)((size_t)stage_kb * 1024 / ((size_t)std::max(1, ncols) * elem)) / 32 * 32;
  p.R = std::max(32, std::min(R, 4096));
  p.stages = stages;
  p.stage_bytes = (uint32_t)(((size_t)std::max(1, ncols) * p.R * elem + 127) / 128 * 128);
  return p;
}
template <class K> static int pipe_smem_attr(K kernel, bg_context_t ctx) {
  static bool done[64] = {};  // per kernel instantiation and per device (the attribute is per device)
  const int d = ctx->device & 63;
  if (!done[d]) {
    BG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 8192));
    done[d] = true;
  }
  return BG_OK;
}

template <class VT>
static int dots_dispatch(bg_context_t ctx, int n, const VT* V, int64_t ld, int j, const VT* w, const GmresDev& st,
                         int fin, int i, int mk, int om, int nr, double* red_out, const PeerAR* peer) {
  if (j < 0 || j > kFinMax) { set_error("dots: j=%d exceeds one pass (max %d columns)", j, kFinMax); return BG_ERR_ARG; }
  const Pipe pp = pipe_geometry(j + 1, sizeof(VT));
  const size_t smem = (size_t)pp.stages * pp.stage_bytes + 2 * sizeof(uint64_t) * pp.stages;
  BG_TRY(pipe_smem_attr(dots_kernel<VT>, ctx));
  const int ntiles = (n + pp.R - 1) / pp.R;
  const int g = std::max(1, std::min(ntiles, ctx->sm_count));
  dots_kernel<VT><<<g, kPipeThreads, smem, ctx->stream>>>(n, V, ld, j, w, st, fin, i, mk, om, nr, pp, red_out, peer);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_dots(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int j, const void* w, const GmresDev& st,
                int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth, double* red_out, const PeerAR* peer) {
  if (dtype == BG_F32) return dots_dispatch<float>(ctx, n, (const float*)V, ld, j, (const float*)w, st, fin, inner_i, mgs_k, orth_mode, need_reorth, red_out, peer);
  return dots_dispatch<double>(ctx, n, (const double*)V, ld, j, (const double*)w, st, fin, inner_i, mgs_k, orth_mode, need_reorth, red_out, peer);
}
int launch_halo_exchange(bg_context_t ctx, int dtype, const PeerHalo* d_ph, int n_send, int n_halo, const int* idx,
                         const void* v, void* halo_out, unsigned seq, const int* skip) {
  const int work = std::max(n_send, n_halo);
  const int g = std::max(1, std::min((work + 255) / 256, ctx->sm_count));
  if (dtype == BG_F32) halo_exchange_kernel<float><<<g, 256, 0, ctx->stream>>>(d_ph, idx, (const float*)v, (float*)halo_out, seq, skip);
  else halo_exchange_kernel<double><<<g, 256, 0, ctx->stream>>>(d_ph, idx, (const double*)v, (double*)halo_out, seq, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_finalize(bg_context_t ctx, const GmresDev& st, int fin, int inner_i, int mgs_k, int orth_mode, int need_reorth, int j,
                    const double* red_in) {
  finalize_kernel<<<1, 128, 0, ctx->stream>>>(st, fin, inner_i, mgs_k, orth_mode, need_reorth, j, red_in);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_pack(bg_context_t ctx, int dtype, int cnt, const int* idx, const void* v, void* buf, const int* skip) {
  if (cnt <= 0) return BG_OK;
  const int g = std::max(1, std::min((cnt + 255) / 256, ctx->sm_count * 4));
  if (dtype == BG_F32) pack_kernel<float><<<g, 256, 0, ctx->stream>>>(cnt, idx, (const float*)v, (float*)buf, skip);
  else pack_kernel<double><<<g, 256, 0, ctx->stream>>>(cnt, idx, (const double*)v, (double*)buf, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

template <class VT>
static int update_dispatch(bg_context_t ctx, int n, const VT* a, const VT* V, int64_t ld, int j, const double* coef,
                           double sgn, VT* out, const GmresDev& st, int us, int force, int nr) {
  if (j < 0 || j > kFinMax) { set_error("update: j=%d exceeds one pass (max %d columns)", j, kFinMax); return BG_ERR_ARG; }
  if (j == 0) {
    scale_copy_kernel<VT><<<vec_grid(ctx, n, 1, 8), 256, 0, ctx->stream>>>(n, a, out, st, us, force);
  } else {
    const Pipe pp = pipe_geometry(j, sizeof(VT));
    const size_t smem = (size_t)pp.stages * pp.stage_bytes + 2 * sizeof(uint64_t) * pp.stages;
    BG_TRY(pipe_smem_attr(update_kernel<VT>, ctx));
    const int ntiles = (n + pp.R - 1) / pp.R;
    const int g = std::max(1, std::min(ntiles, ctx->sm_count));
    update_kernel<VT><<<g, kPipeThreads, smem, ctx->stream>>>(n, a, V, ld, j, coef, sgn, out, st, us, force, nr, pp);
  }
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

// ring geometry of the fused step: same stage budget, rows chosen so that every CTA owns the same number of tiles
// j basis columns of elem_b bytes + one work column of elem_w bytes per row
static Pipe orth_geometry(int j, size_t elem_b, size_t elem_w, int n, int G) {
  const size_t row_bytes = (size_t)j * elem_b + elem_w;
  Pipe p = pipe_geometry(1, row_bytes);
  const int64_t per = ((int64_t)n + G - 1) / G;                  // rows per CTA
  const int64_t k = std::max<int64_t>(1, (per + p.R - 1) / p.R); // tiles per CTA
  int R = (int)(((per + k - 1) / k + 31) / 32 * 32);
  p.R = std::max(32, std::min(R, p.R));
  p.stage_bytes = (uint32_t)((row_bytes * p.R + 127) / 128 * 128);
  return p;
}
template <class VT, class BT>
static int orth_dispatch(bg_context_t ctx, int n, const BT* V, int64_t ld, int j, const VT* w, VT* out, BT* out_b, const GmresDev& st,
                         int i, int om, const PeerAR* peer) {
  if (j < 1 || j > kFinMax) { set_error("orth: j=%d outside 1..%d", j, kFinMax); return BG_ERR_ARG; }
  const int G = std::max(1, std::min((n + 31) / 32, ctx->sm_count));
  Pipe pp = orth_geometry(j, sizeof(BT), sizeof(VT), n, G);
  const size_t smem = (size_t)pp.stages * pp.stage_bytes + 2 * sizeof(uint64_t) * pp.stages;
  auto kern = orth_kernel<VT, BT>;
  BG_TRY(pipe_smem_attr(kern, ctx));
  GmresDev stv = st;
  void* args[] = {(void*)&n, (void*)&V, (void*)&ld, (void*)&j, (void*)&w, (void*)&out, (void*)&out_b, (void*)&stv, (void*)&i, (void*)&om, (void*)&pp, (void*)&peer};
  // cooperative launch: the CTAs wait for each other inside the kernel, so co-residency must be guaranteed
  BG_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(G), dim3(kPipeThreads), args, smem, ctx->stream));
  ctx->launches++;
  return BG_OK;
}
#ifdef BG_MGS_TIMING
extern "C" BG_API int bg_debug_mgs_times(int probe_cta, unsigned long long* out200) {
  if (probe_cta >= 0) return cudaMemcpyToSymbol(g_mgs_probe, &probe_cta, sizeof(int)) == cudaSuccess ? 0 : -1;
  return cudaMemcpyFromSymbol(out200, g_mgs_t, sizeof(unsigned long long) * 200) == cudaSuccess ? 0 : -1;
}
#endif
// rows per CTA of the fused MGS step, or 0 when the vector does not fit the register-resident layout
static int mgs_rows_per(bg_context_t ctx, int n, int G) {
  const int per = (((n + G - 1) / G) + 31) / 32 * 32;
  return per <= kMgsThreads * kMgsEPT ? per : 0;
}
bool mgs_fused_supported(bg_context_t ctx, int n, int m) {
  if (n < 1) return false;
  const int G = std::max(1, std::min((n + 31) / 32, ctx->sm_count));
  return G <= 160 && mgs_rows_per(ctx, n, G) > 0;
}
size_t mgs_slot_bytes(bg_context_t ctx, int m) { return (size_t)(m + 2) * 256 * sizeof(LL16); }
template <class VT>
static int mgs_dispatch(bg_context_t ctx, int n, const VT* V, int64_t ld, int i, const VT* w, VT* out, const GmresDev& st, void* slots,
                        unsigned seq) {
  const int G = std::max(1, std::min((n + 31) / 32, ctx->sm_count));
  MgsArgs a{(LL16*)slots, seq, mgs_rows_per(ctx, n, G), 5ull * 1000000000ull};
  if (a.rows_per == 0 || G > 160) { set_error("fused MGS: n=%d does not fit", n); return BG_ERR_UNSUPPORTED; }
  const size_t smem = 2 * (((size_t)a.rows_per * sizeof(VT) + 127) & ~(size_t)127);
  BG_TRY(pipe_smem_attr(mgs_kernel<VT>, ctx));
  GmresDev stv = st;
  void* args[] = {(void*)&n, (void*)&V, (void*)&ld, (void*)&i, (void*)&w, (void*)&out, (void*)&stv, (void*)&a};
  // cooperative launch: the CTAs exchange partial sums inside the kernel, so co-residency must be guaranteed
  BG_CUDA(cudaLaunchCooperativeKernel((const void*)mgs_kernel<VT>, dim3(G), dim3(kMgsThreads), args, smem, ctx->stream));
  ctx->launches++;
  return BG_OK;
}
int launch_mgs(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int i, const void* w, void* out, const GmresDev& st,
               void* slots, unsigned seq) {
  if (dtype == BG_F32) return mgs_dispatch<float>(ctx, n, (const float*)V, ld, i, (const float*)w, (float*)out, st, slots, seq);
  return mgs_dispatch<double>(ctx, n, (const double*)V, ld, i, (const double*)w, (double*)out, st, slots, seq);
}

int launch_orth(bg_context_t ctx, int dtype, int n, const void* V, int64_t ld, int j, const void* w, void* out, const GmresDev& st,
                int inner_i, int orth_mode, const PeerAR* peer) {
  if (dtype == BG_F32) return orth_dispatch<float, float>(ctx, n, (const float*)V, ld, j, (const float*)w, (float*)out, nullptr, st, inner_i, orth_mode, peer);
  return orth_dispatch<double, double>(ctx, n, (const double*)V, ld, j, (const double*)w, (double*)out, nullptr, st, inner_i, orth_mode, peer);
}
// FP64 work vectors, FP32 basis storage: V32 = basis (ld floats per column), out = working copy of the new vector,
// out_b = its column in the basis
int launch_orth_mixed(bg_context_t ctx, int n, const float* V32, int64_t ld, int j, const double* w, double* out, float* out_b,
                      const GmresDev& st, int inner_i, int orth_mode) {
  return orth_dispatch<double, float>(ctx, n, V32, ld, j, w, out, out_b, st, inner_i, orth_mode, nullptr);
}

// mixed-precision basis, restart terms (once per cycle: plain grid-stride kernels)
//   round_store: v <- (double)(float)v, v32 <- (float)v        (v0 = r / beta)
//   axpy_basis:  x += sum_k y[k] V32[:,k]                       (x += V y, Update_GPU)
__global__ void round_store_kernel(int n, double* __restrict__ v, float* __restrict__ v32, const int* skip) {
  if (skip && *skip) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float f = (float)v[i];
    v32[i] = f;
    v[i] = (double)f;
  }
}
__global__ void __launch_bounds__(256) axpy_basis_kernel(int n, const float* __restrict__ V32, int64_t ld, int k, const double* __restrict__ y,
                                                         double* __restrict__ x, GmresDev st, int force) {
  if (!force && st.flags[F_CONV]) return;
  __shared__ double ys[kFinMax + 1];
  if ((int)threadIdx.x < k) ys[threadIdx.x] = y[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = x[i];
#pragma unroll 8
    for (int c = 0; c < k; ++c) acc += ys[c] * (double)V32[(size_t)c * ld + i];
    x[i] = acc;
  }
}
int launch_round_store(bg_context_t ctx, int n, double* v, float* v32, const int* skip) {
  round_store_kernel<<<vec_grid(ctx, n, 1, 8), 256, 0, ctx->stream>>>(n, v, v32, skip);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_axpy_basis(bg_context_t ctx, int n, const float* V32, int64_t ld, int k, const double* y, double* x, const GmresDev& st, int force) {
  if (k < 1 || k > kFinMax + 1) { set_error("axpy_basis: k=%d", k); return BG_ERR_ARG; }
  axpy_basis_kernel<<<vec_grid(ctx, n, 1, 8), 256, 0, ctx->stream>>>(n, V32, ld, k, y, x, st, force);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_solve_y(bg_context_t ctx, const GmresDev& st, int k, int force) {
  solve_y_kernel<<<1, 256, 0, ctx->stream>>>(st, k, force);
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
int launch_gather(bg_context_t ctx, int dtype, int n, const int* perm, const void* in, void* out, const void* scale, int scale_dt) {
  const int g = vec_grid(ctx, n, 1);
  if (dtype == BG_F32 && scale_dt == BG_F32) gather_kernel<float, float><<<g, 256, 0, ctx->stream>>>(n, perm, (const float*)in, (float*)out, (const float*)scale);
  else if (dtype == BG_F32) gather_kernel<float, double><<<g, 256, 0, ctx->stream>>>(n, perm, (const float*)in, (float*)out, (const double*)scale);
  else if (scale_dt == BG_F32) gather_kernel<double, float><<<g, 256, 0, ctx->stream>>>(n, perm, (const double*)in, (double*)out, (const float*)scale);
  else gather_kernel<double, double><<<g, 256, 0, ctx->stream>>>(n, perm, (const double*)in, (double*)out, (const double*)scale);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
int launch_scatter(bg_context_t ctx, int dtype, int n, const int* perm, const void* in, void* out) {
  const int g = vec_grid(ctx, n, 1);
  if (dtype == BG_F32) scatter_kernel<float><<<g, 256, 0, ctx->stream>>>(n, perm, (const float*)in, (float*)out);
  else scatter_kernel<double><<<g, 256, 0, ctx->stream>>>(n, perm, (const double*)in, (double*)out);
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
