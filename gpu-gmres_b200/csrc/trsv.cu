// trsv.cu -- sparse triangular solves for the ILU preconditioner, sm_100a.
//
// Replaces cusparse{S,D}csrsv_analysis/solve (src/preconditioner.cu:263,273,356-358,1313-1315,
// 1612,1637) and the nine element-wise scale/permute/convert kernels wrapped around them
// (src/preconditioner.cu:1424-1558), which are folded into the solve's prologue / epilogue here.
//
// Schedules (DESIGN.md "SpTRSV"):
//  * LEVEL  -- the reference's operator: the GLOBAL factors, dependencies = the level sets of getLevel
//    (src/gmres.cu:1033-1077).  Two all-SM kernels, no launch and no grid barrier per level:
//      - trsv_syncfree (any factor): rows sorted by level and cut into 32-row slices that never span two
//        levels; persistent warps take slices in order and wait for exactly the rows they read through
//        flag-in-data packets in L2 (value + sequence number in one 16-byte store), so a level costs one
//        L2 hand-off instead of a launch / grid barrier, and independent chains overlap.
//      - trsv_lattice_global (ILU(0) factors of a 5-point grid in natural order, detected on the host): one
//        CTA per 32-column strip; its chain warp walks the anti-diagonal wavefront (south / north = own previous
//        result, west / east = neighbour lane by shuffle) on operands that loader warps stage in shared memory,
//        a storer warp writes the results back; only the strip boundary column crosses L2.  No indices are
//        stored; a level costs a shuffle and two FMAs.
//    The device schedules follow the STRUCTURAL levels (every stored entry is a dependency); getLevel itself, which
//    stops at / skips entries below 1e-9, stays the parity artefact (bg_precond_get_levels).
//  * BLOCK  -- block-Jacobi ILU, see trsv_block.cu.
#include <type_traits>

#include "internal.h"
#include "peer.cuh"

namespace bg {

// ------------------------------------------------------------------ prologue / epilogue
// The split preconditioner's element-wise kernels, folded into the solve.  The permutations are composed with the
// stream's slot order on the host (SyncFreeDev::src_idx / dst_idx), so the device sees one index per slot.
//   pro 1: rhs = in[src] / s_pro[src]   kernel_divide_S + kernel_perm_precondLeft_S2D (preconditioner.cu:1439-1464)
//   pro 2: rhs = in[src] * s_pro[src]   kernel_multiply_S2D (preconditioner.cu:1518-1531)
//   epi 1: out[dst] = x / s_epi[dst]    kernel_perm_divide_precondRight_D2S (preconditioner.cu:1500-1516), as a scatter
template <class VT> struct XformDev {
  int pro, epi;
  const int* src_idx;  // per slot; null: the row itself
  const int* dst_idx;
  const VT* s_pro;
  const VT* s_epi;
};

// ------------------------------------------------------------------ sync-free LEVEL schedule (any factor)
__device__ __forceinline__ LL16 ll_load(const LL16* p) { return ll_load_gpu(p); }  // hand-offs stay inside one GPU
__device__ __forceinline__ bool ll_ok(const LL16& v, unsigned seq) { return v.f0 == seq && v.f1 == seq; }
__device__ __forceinline__ double ll_val(const LL16& v) { return __hiloint2double((int)v.hi, (int)v.lo); }

template <class FT> struct SfArgs {
  const int4* meta;  // host_algos.h SyncFreeStream::meta
  const int *rows, *cols;
  const FT *vals, *diag;
  int n_slices;
  unsigned seq;
  unsigned long long timeout_ns;
};

// One warp per slice (<= 32 rows of ONE level), slices taken in ascending order by a fixed set of resident warps:
// a warp only ever waits for rows / levels of earlier slices, which are finished or held by a running warp, so the
// walk cannot deadlock as long as the whole grid is resident (launcher: occupancy query).
//
// A warp issues in order and stalls at the first USE of a load that has not landed, so the kernel is written as a
// software pipeline in which every load is issued one stage before its first use:
//   slice q+2G : metadata (one 16-byte load)
//   slice q+G  : row id, source / destination index, diagonal, first group of columns and values   (needs metadata)
//   slice q    : right-hand side and scales (need the indices), then per group of KP entries: packets of this group
//                and columns / values of the next one in flight together, FMAs of this group behind the poll
// so a slice costs one round trip per group and nothing else; measured before the pipeline: 2.2 us from "slice taken"
// to "operands loaded" and 1.4 us from "dependencies arrived" to "published" (the epilogue's dependent loads), on top
// of K/4 round trips.
//  * Rows are right-aligned in the slice and their entries ordered by the operand's level (host packer): the operands
//    that arrive last are in the LAST group of every lane, so when the warp runs ahead of the dependency front (the
//    normal case on a thin level) only one group lies on the level's critical path.  With left-aligned rows the warp
//    walked all K/4 groups after the front arrived: 10 us per level on the long rows of ILU++ factors.
//  * The packet loads of a group leave together and the warp re-polls, AS A WHOLE, only the ones that were not ready,
//    until a warp vote says every lane has its operands: 32 lanes spinning in private loops are executed one
//    divergent path after the other, which measured 10 us per dependency hop instead of one L2 round trip
//    (tools/probe/relay_probe.cu, profiles/r02_syncfree_probe.txt).
//  * Tried and dropped: gating the warps ahead of the front behind a per-level "complete" packet (one lane polling one
//    address) to keep their packet polls off L2.  The completion signal (atomic count + packet) arrives later than the
//    operands themselves: 346 us -> 446..542 us per L+U solve on the ILU++ factors of the 1304^2 grid.
// Per row: acc = rhs; acc -= v_k x_k (k in the stored order of the stream); acc /= diag.
constexpr int kSfKP = 4;
#ifdef BG_SF_TIMING
// development aid (make dbg): %globaltimer stamps of lane 0 of the warps that solve slices [g_sf_q0, g_sf_q0 + kSfCap):
// [0] slice taken  [1] right-hand side loaded  [2] all dependencies arrived  [3] result published  (+ SM id)
constexpr int kSfCap = 65536;
__device__ unsigned long long g_sf_t[kSfCap * 5];
__device__ int g_sf_q0 = 0;
#define SF_STAMP(k) do { if (lane == 0 && q >= g_sf_q0 && q < g_sf_q0 + kSfCap) g_sf_t[(q - g_sf_q0) * 5 + (k)] = global_ns(); } while (0)
#else
#define SF_STAMP(k) do {} while (0)
#endif
template <class VT, class FT, class CT, bool UNIT>
__global__ void __launch_bounds__(128, 5) trsv_syncfree(SfArgs<FT> a, XformDev<VT> xf, const VT* __restrict__ in, LL16* xs,
                                                     VT* out, const int* skip) {
  pdl_wait();
  if (skip && *skip) return;
  constexpr int KP = kSfKP;
  const int lane = threadIdx.x & 31;
  const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), G = (int)((gridDim.x * blockDim.x) >> 5);
  if (g >= a.n_slices) return;
  const int last = a.n_slices;  // meta[n_slices] is a copy of the last slice: prefetches past the end stay in bounds

  // stage "row" of a slice: everything that needs only the metadata
  struct Row { int r, src, dst; CT dg; int c[KP]; CT v[KP]; };
  auto load_row = [&](const int4& m, Row& w) {
    const int R = m.y & 255, K = m.y >> 8;
    const bool mine = lane < R;
    const int slot = m.x + (mine ? lane : 0);
    w.r = a.rows[slot];
    w.src = xf.src_idx ? xf.src_idx[slot] : w.r;
    w.dst = xf.dst_idx ? xf.dst_idx[slot] : w.r;
    w.dg = UNIT ? (CT)1 : (CT)a.diag[slot];
    const int64_t base = (int64_t)m.w * 32 + lane;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const bool has = mine && j < K;
      w.c[j] = has ? a.cols[base + (int64_t)j * 32] : -1;
      w.v[j] = has ? (CT)a.vals[base + (int64_t)j * 32] : (CT)0;
    }
  };

  int4 m_cur = a.meta[g];
  int4 m_nxt = a.meta[min(g + G, last)];
  Row w_cur;
  load_row(m_cur, w_cur);
  for (int q = g; q < a.n_slices; q += G) {
    SF_STAMP(0);
    const int R = m_cur.y & 255, K = m_cur.y >> 8;
    const bool mine = lane < R;  // lanes beyond the slice stay in the loops with nothing to wait for
    const int64_t base = (int64_t)m_cur.w * 32 + lane;
    const int r = w_cur.r, dst = w_cur.dst;
    const CT dg = w_cur.dg;
    // this slice: right-hand side and scales (in flight during the polls)
    const VT rin = in[w_cur.src];
    const VT spro = xf.pro ? xf.s_pro[w_cur.src] : (VT)1;
    const VT sepi = xf.epi ? xf.s_epi[dst] : (VT)1;
    int c[KP];
    CT v[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) { c[j] = w_cur.c[j]; v[j] = w_cur.v[j]; }
    // the warp's next slice: row stage; the one after: metadata
    const int4 m_nn = a.meta[min(q + 2 * G, last)];
    Row w_nxt;
    load_row(m_nxt, w_nxt);
    auto rhs = [&]() -> CT {
      if (xf.pro == 1) { const VT t = rin / spro; return (CT)t; }
      if (xf.pro == 2) { const VT t = rin * spro; return (CT)t; }
      return (CT)rin;
    };
#ifdef BG_SF_TIMING
    if (rin == (VT)1.2345e30) return;  // forces the operand loads to complete before the stamp
    SF_STAMP(1);
    if (lane == 0 && q >= g_sf_q0 && q < g_sf_q0 + kSfCap) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_sf_t[(q - g_sf_q0) * 5 + 4] = sm; }
#endif
    CT acc = (CT)0;
    if (K == 0) acc = rhs();
    for (int k0 = 0; k0 < K; k0 += KP) {
      int cn[KP];
      CT vn[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) {  // next group: in flight while this one is polled
        const bool has = mine && k0 + KP + j < K;
        cn[j] = has ? a.cols[base + (int64_t)(k0 + KP + j) * 32] : -1;
        vn[j] = has ? (CT)a.vals[base + (int64_t)(k0 + KP + j) * 32] : (CT)0;
      }
      LL16 pk[KP];
      bool need[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        need[j] = c[j] >= 0;
        pk[j] = LL16{0u, 0u, 0u, 0u};
      }
      unsigned long long t_start = 0;
      for (unsigned spin = 0;; ++spin) {  // warp-uniform: leaves when every lane has all its operands
        bool all = true;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          if (need[j]) {  // predicated reload of the packets still missing
            pk[j] = ll_load(xs + c[j]);
            need[j] = !ll_ok(pk[j], a.seq);
          }
          all = all && !need[j];
        }
        if (__all_sync(0xffffffffu, all)) break;
        if ((spin & 4095u) == 4095u) {
          const unsigned long long now = global_ns();
          if (t_start == 0) t_start = now;
          else if (now - t_start > a.timeout_ns) __trap();
        }
      }
      if (k0 == 0) acc = rhs();
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        if (c[j] >= 0) acc -= v[j] * (CT)ll_val(pk[j]);
        c[j] = cn[j];
        v[j] = vn[j];
      }
    }
    SF_STAMP(2);
    if (!UNIT) acc /= dg;
    if (mine) {
      ll_store_gpu(xs + r, (double)acc, a.seq);
      if (out) {
        if (xf.epi == 1) { VT t = (VT)acc; t /= sepi; out[dst] = t; }
        else out[dst] = (VT)acc;
      }
    }
    SF_STAMP(3);
    m_cur = m_nxt;
    m_nxt = m_nn;
    w_cur = w_nxt;
  }
}
#ifdef BG_SF_TIMING
}  // namespace bg
extern "C" BG_API int bg_debug_sf_times(int q0, unsigned long long* out) {
  if (q0 >= 0) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, bg::g_sf_t) != cudaSuccess || cudaMemset(p, 0, sizeof(unsigned long long) * bg::kSfCap * 5) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(bg::g_sf_q0, &q0, sizeof(int)) == cudaSuccess ? 0 : -1;
  }
  return cudaMemcpyFromSymbol(out, bg::g_sf_t, sizeof(unsigned long long) * bg::kSfCap * 5) == cudaSuccess ? 0 : -1;
}
namespace bg {
#endif

// ------------------------------------------------------------------ global lattice wavefront (5-point ILU(0))
__device__ __forceinline__ uint32_t lg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES> __device__ __forceinline__ void cp_async(void* dst, const void* src) {
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lg_smem_u32(dst)), "l"(src) : "memory");
  else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(lg_smem_u32(dst)), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lg_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class FT> struct LgArgs {
  const FT* pair;  // host_algos.h LatticeGlobal::lower / upper
  const FT* dinv;  // LatticeGlobal::udinv (upper only)
  LL16* bnd;       // [strips][ny] boundary column of every strip (flag-in-data packets)
  int nx, ny, strips;
  unsigned seq;
  unsigned long long timeout_ns;
};
// 8 steps per block: 4 were measured 30 % slower (the per-block cost -- progress polls, publication, the burst of
// operand loads -- is ~480 cycles), 16 do not fit the shallow ring and lengthen every strip-to-strip offset by 16 steps
constexpr int kLgBS = 8;  // steps per block: granularity of staging, boundary hand-off and result flush
template <class TIN, class TOUT, class FT, bool LOWER> constexpr size_t lg_smem_bytes(int NR) {
  return (size_t)NR * 32 * ((LOWER ? 2 : 3) * sizeof(FT) + sizeof(TIN) + sizeof(TOUT)) + (size_t)NR * 8 + 64;
}
// progress counters in shared memory: the producer publishes "blocks finished" behind a CTA fence, the consumer
// polls it (one 30-cycle LDS when the block is already there, an mbarrier try_wait costs three times that)
// Hand-offs between the warps of one CTA go through shared memory only: data (st.shared / cp.async completed by
// wait_group) first, then the volatile progress word, both from the same warp after a __syncwarp, and the consumer
// reads the word before the data.  Shared-memory accesses of a warp are performed in program order, so no memory
// fence is needed -- a MEMBAR.CTA here measured ~900 cycles per 8-step block on the chain warp, four times the
// recurrence itself (tools/lg_probe.py) -- only the compiler must not reorder the accesses.
__device__ __forceinline__ void lg_fence_cta() { asm volatile("" ::: "memory"); }
__device__ __forceinline__ void lg_publish(volatile int* p, int v) {
  lg_fence_cta();
  *p = v;
}
template <bool SLEEP> __device__ __forceinline__ void lg_wait_progress(volatile int* p, int need, unsigned long long timeout_ns) {
  unsigned long long t0 = 0;
  for (unsigned spin = 0; *p < need; ++spin) {
    if (SLEEP) __nanosleep(32);
    if ((spin & 4095u) == 4095u) {  // a scheduling bug or a dead upstream strip must trap, never hang the GPU
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) __trap();
    }
  }
  lg_fence_cta();
}

#ifdef BG_SF_TIMING
// development aid (make dbg): clock64 accounting of strip g_lg_probe's warps: [0] chain total [1] chain waiting
// [2] chain waiting on the loader [3] ... on the boundary [4] ... on the storer [5] loader total [6] loader flow-control
// [7] loader wait_group [8] storer total [9] storer waiting [10] boundary total [11] boundary polling
__device__ long long g_lg_t[16];
__device__ int g_lg_probe = 0;
#define LG_T0(v) long long v = clock64()
#define LG_ACC(k, v) do { lg_acc[k] += clock64() - (v); } while (0)   // registers; flushed once at the end of the warp's role
#define LG_DECL long long lg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define LG_FLUSH do { if (s == g_lg_probe && lane == 0) for (int k_ = 0; k_ < 12; ++k_) if (lg_acc[k_]) atomicAdd((unsigned long long*)&g_lg_t[k_], (unsigned long long)lg_acc[k_]); } while (0)
#else
#define LG_T0(v) do {} while (0)
#define LG_ACC(k, v) do {} while (0)
#define LG_DECL do {} while (0)
#define LG_FLUSH do {} while (0)
#endif
// the 8 steps of one block on the chain warp.  The rings are indexed by STEP (the loader puts lane x's element of
// time row t where step t + lag will look for it), so the operands of the block's 8 steps sit at constant offsets
// from one base address that never wraps inside a block: no address arithmetic on the chain warp.
// INTERIOR: every lane is inside [0, ny) for all 8 steps and the strip is full, so the operands need no masking.
template <class TIN, class TOUT, class FT, class CT, bool LOWER, int NR, bool INTERIOR, class P>
__device__ __forceinline__ void lg_chain_block(const P* __restrict__ pr, const FT* __restrict__ dv, const TIN* __restrict__ rh,
                                               TOUT* __restrict__ ot, const double* __restrict__ bd, int t0, bool bnd, bool colok,
                                               int ny, bool is_edge, CT& zprev, LL16* edge_out, unsigned seq) {
  constexpr int BS = kLgBS;
  CT ca[BS], cb[BS], r[BS], be[BS];
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    const P c = pr[j * 32];
    CT rr = (CT)rh[j * 32];
    if (!LOWER) rr *= (CT)dv[j * 32];
    if (INTERIOR) {
      ca[j] = (CT)c.x; cb[j] = (CT)c.y; r[j] = rr;
    } else {
      const bool act = colok && (unsigned)(t0 + j) < (unsigned)ny;
      ca[j] = act ? (CT)c.x : (CT)0;
      cb[j] = act ? (CT)c.y : (CT)0;
      r[j] = act ? rr : (CT)0;
    }
    be[j] = bnd ? (CT)bd[j] : (CT)0;  // same address for all lanes: broadcast (rows past ny-1 are never used by an active lane)
  }
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    CT zn = LOWER ? __shfl_up_sync(0xffffffffu, zprev, 1) : __shfl_down_sync(0xffffffffu, zprev, 1);
    zn = is_edge ? be[j] : zn;
    const CT tt = r[j] - cb[j] * zprev;  // south / north: own previous result
    zprev = tt - ca[j] * zn;             // west / east: neighbour lane (or the upstream strip)
    ot[j * 32] = (TOUT)zprev;            // steps outside [0, ny) are never flushed
    // the strip's last lane hands its value to the downstream strip the moment it exists (edge_out: this lane's
    // packet of time row t0 + j, null on every other lane); waiting for the storer's block flush kept every strip
    // ~8 steps plus a shared-memory hand-off further behind its neighbour
    if (edge_out && (INTERIOR || (unsigned)(t0 + j) < (unsigned)ny)) ll_store_gpu(edge_out + (LOWER ? j : -j), (double)zprev, seq);
  }
}

// One CTA (= one SM) per strip of 32 grid columns, four specialised warps.  "Time row" t = y (LOWER, strips left
// to right) or ny-1-y (UPPER, strips right to left); lane x of the CHAIN warp reaches time row t at step t + lag,
// lag = x (LOWER) / 31-x (UPPER).  A lone warp issues about one dependent instruction every four cycles, so
// everything that is not the recurrence itself is moved off the chain warp:
//   warps 3..6 LOADERS  stage the coefficient pairs, 1/diag and the right-hand side of 8 time rows per block (two
//                    rows each) with coalesced cp.async copies into a ring of NR rows, several blocks ahead of the
//                    chain.  Four of them because a lone warp needs ~900 cycles to issue a block's 24 copies with
//                    their address arithmetic, three times what the chain needs to consume it (ncu: the chain
//                    polled 22 times per block behind a single loader)
//   warp 1 BOUNDARY  polls the upstream strip's boundary column (flag-in-data packets in L2, warp-uniform vote
//                    loop) and drops the values into shared memory
//   warp 0 CHAIN     per step: operands from shared memory (its own column) into registers a block at a time,
//                    west/east value by shuffle (edge lane: the boundary value), two FMAs, result to a ring
//   warp 2 STORER    once all 32 lanes have passed a row: one coalesced 256-byte store of the row, and the edge
//                    column of the block goes to the downstream strip as 8 packets
// Hand-offs are "blocks finished" counters in shared memory: [0] chain, [1] storer, [3] boundary, [4..7] loaders.
constexpr int kLgLoaders = 4, kLgThreads = 32 * (3 + kLgLoaders);
template <class TIN, class TOUT, class FT, class CT, bool LOWER, int NR>
__global__ void __launch_bounds__(kLgThreads) trsv_lattice_global(LgArgs<FT> a, const TIN* __restrict__ in, TOUT* __restrict__ out,
                                                           const int* skip) {
  constexpr int BS = kLgBS;
  constexpr int AHEAD = (NR - 31) / BS - 1;  // blocks a producer may run ahead of the consumer that frees its ring rows
  constexpr int INFLIGHT = AHEAD > 8 ? 6 : 2;  // cp.async groups (blocks) the loader keeps in flight
  static_assert((NR & (NR - 1)) == 0 && AHEAD >= 2 && INFLIGHT < AHEAD, "ring too small for the lane skew");
  extern __shared__ __align__(16) unsigned char lg_smem[];
  using P = typename std::conditional<sizeof(FT) == 8, double2, float2>::type;
  unsigned char* sp = lg_smem;
  P* s_pair = reinterpret_cast<P*>(sp); sp += (size_t)NR * 32 * sizeof(P);
  FT* s_dinv = reinterpret_cast<FT*>(sp); sp += LOWER ? 0 : (size_t)NR * 32 * sizeof(FT);
  TIN* s_rhs = reinterpret_cast<TIN*>(sp); sp += (size_t)NR * 32 * sizeof(TIN);
  TOUT* s_out = reinterpret_cast<TOUT*>(sp); sp += (size_t)NR * 32 * sizeof(TOUT);
  double* s_bnd = reinterpret_cast<double*>(sp); sp += (size_t)NR * 8;
  volatile int* prog = reinterpret_cast<volatile int*>(sp);
  pdl_wait();
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < 8) prog[threadIdx.x] = 0;
  __syncthreads();
  const int s = LOWER ? (int)blockIdx.x : a.strips - 1 - (int)blockIdx.x;  // CTAs in dependency order
  const int X = 32 * s + lane, nx = a.nx, ny = a.ny;
  const bool colok = X < nx;
  const int nblk = (ny + 31 + BS - 1) / BS;
  const bool has_up = LOWER ? s > 0 : s < a.strips - 1, has_down = LOWER ? s < a.strips - 1 : s > 0;
  const int edge_in = LOWER ? 0 : 31, edge_out = LOWER ? 31 : 0;
  const unsigned long long tmo = a.timeout_ns;
  auto yrow = [&](int t) { return LOWER ? t : ny - 1 - t; };
  LG_DECL;

  if (wid >= 3) {  // ---------------------------------------------------------------- LOADERS
    constexpr int NL = kLgLoaders, RPL = BS / NL;  // rows of each block per loader warp
    const int lw = wid - 3;
    const P* gpair = reinterpret_cast<const P*>(a.pair) + (size_t)s * ny * 32 + lane;
    const FT* gdinv = a.dinv + (size_t)s * ny * 32 + lane;
    const int lag = LOWER ? lane : 31 - lane;  // ring row = the STEP at which this lane consumes the element
    LG_T0(tl);
    for (int lb = 0; lb < nblk; ++lb) {
      { LG_T0(tf); if (lb > AHEAD) lg_wait_progress<false>(&prog[0], lb - AHEAD, tmo); LG_ACC(6, tf); }  // the ring rows about to be overwritten are dead
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        const int t = lb * BS + lw * RPL + j;
        if (t < ny) {
          const int y = yrow(t), slot = ((t + lag) & (NR - 1)) * 32 + lane;
          cp_async<sizeof(P)>(s_pair + slot, gpair + (size_t)y * 32);
          if (!LOWER) cp_async<sizeof(FT)>(s_dinv + slot, gdinv + (size_t)y * 32);
          if (colok) cp_async<sizeof(TIN)>(s_rhs + slot, in + (size_t)y * nx + X);
        }
      }
      cp_async_commit();
      { LG_T0(tw); cp_async_wait<INFLIGHT>(); LG_ACC(7, tw); }  // block lb - INFLIGHT (and everything older) has landed for this lane
      __syncwarp();               // ... and for every lane
      if (lane == 0 && lb >= INFLIGHT) lg_publish(&prog[4 + lw], lb - INFLIGHT + 1);
    }
    cp_async_wait<0>();
    __syncwarp();
    if (lane == 0) lg_publish(&prog[4 + lw], nblk);
    LG_ACC(5, tl);
    if (lw == 0) LG_FLUSH;
  } else if (wid == 1) {  // --------------------------------------------------------- BOUNDARY IN
    if (has_up) {
      const LL16* bnd_in = a.bnd + (size_t)(LOWER ? s - 1 : s + 1) * ny;
      for (int b = 0; b * BS < ny; ++b) {
        if (b > AHEAD) lg_wait_progress<false>(&prog[0], b - AHEAD, tmo);
        const int t = b * BS + lane;  // the edge lane has no lag: it reaches time row t at step t
        bool need = lane < BS && t < ny;
        LL16 pk = {0u, 0u, 0u, 0u};
        unsigned long long t0 = 0;
        // (three staggered polls in flight instead of one were measured: the priming delay costs more than the finer
        // detection grain buys, 499 -> 691 us at 1304^2)
        for (unsigned spin = 0;; ++spin) {  // warp-uniform poll
          if (need) { pk = ll_load(bnd_in + yrow(t)); need = !ll_ok(pk, a.seq); }
          if (__all_sync(0xffffffffu, !need)) break;
          if ((spin & 4095u) == 4095u) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > tmo) __trap();
          }
        }
        if (lane < BS && t < ny) s_bnd[t & (NR - 1)] = ll_val(pk);
        __syncwarp();
        if (lane == 0) lg_publish(&prog[3], b + 1);
      }
    }
  } else if (wid == 0) {  // --------------------------------------------------------- CHAIN
    const int lag = LOWER ? lane : 31 - lane;
    CT zprev = (CT)0;
    LG_T0(tc);
    for (int blk = 0; blk < nblk; ++blk) {
      const bool bnd = has_up && blk * BS < ny;
      LG_T0(tw);
#ifdef BG_SF_TIMING
      { LG_T0(t2); while (min(min(prog[4], prog[5]), min(prog[6], prog[7])) <= blk) {} LG_ACC(2, t2); }
      { LG_T0(t3); while (bnd && prog[3] <= blk) {} LG_ACC(3, t3); }
      { LG_T0(t4); while (blk > AHEAD && prog[1] < blk - AHEAD) {} LG_ACC(4, t4); }
#endif
      {  // operands staged, boundary values here, result ring rows about to be reused have left: one poll loop, one fence
        unsigned long long tw = 0;
        for (unsigned spin = 0;; ++spin) {
          int4 pl;  // the four loaders' counters in one 16-byte shared load
          asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(pl.x), "=r"(pl.y), "=r"(pl.z), "=r"(pl.w) : "r"(lg_smem_u32(const_cast<int*>(prog + 4))) : "memory");
          const int p2 = min(min(pl.x, pl.y), min(pl.z, pl.w)), p3 = bnd ? prog[3] : blk + 1, p1 = blk > AHEAD ? prog[1] : blk;
          if (p2 > blk && p3 > blk && p1 >= blk - AHEAD) break;
          if ((spin & 4095u) == 4095u) {
            const unsigned long long now = global_ns();
            if (tw == 0) tw = now;
            else if (now - tw > tmo) __trap();
          }
        }
        lg_fence_cta();
      }
      LG_ACC(1, tw);
      const int t0 = blk * BS - lag;  // this lane's time row at the first step of the block
      const int base = ((blk * BS) & (NR - 1)) * 32 + lane;
      const double* bd = s_bnd + ((blk * BS) & (NR - 1));
      // packet of this lane's first time row of the block in the strip's boundary column (edge lane of a strip with a
      // downstream neighbour only); rows run upwards for L (y = t) and downwards for U (y = ny-1-t)
      LL16* eo = (has_down && lane == edge_out) ? a.bnd + (size_t)s * ny + (LOWER ? t0 : ny - 1 - t0) : nullptr;
      if (blk * BS - 31 >= 0 && blk * BS + BS - 1 < ny && 32 * s + 31 < nx)
        lg_chain_block<TIN, TOUT, FT, CT, LOWER, NR, true>(s_pair + base, s_dinv + base, s_rhs + base, s_out + base, bd, t0, bnd, colok, ny, lane == edge_in, zprev, eo, a.seq);
      else
        lg_chain_block<TIN, TOUT, FT, CT, LOWER, NR, false>(s_pair + base, s_dinv + base, s_rhs + base, s_out + base, bd, t0, bnd, colok, ny, lane == edge_in, zprev, eo, a.seq);
      __syncwarp();
      if (lane == 0) lg_publish(&prog[0], blk + 1);
    }
    LG_ACC(0, tc);
    LG_FLUSH;
  } else if (wid == 2) {  // --------------------------------------------------------- STORER
    const int lag = LOWER ? lane : 31 - lane;
    LG_T0(ts);
    for (int b = 0; b < nblk; ++b) {
      { LG_T0(tw); lg_wait_progress<false>(&prog[0], b + 1, tmo); LG_ACC(9, tw); }
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const int tf = b * BS - 31 + j;
        if (colok && (unsigned)tf < (unsigned)ny) out[(size_t)yrow(tf) * nx + X] = s_out[((tf + lag) & (NR - 1)) * 32 + lane];
      }
      __syncwarp();
      if (lane == 0) lg_publish(&prog[1], b + 1);
    }
    LG_ACC(8, ts);
    LG_FLUSH;
  }
}
#ifdef BG_SF_TIMING
}  // namespace bg
extern "C" BG_API int bg_debug_lg_times(int probe, long long* out16) {
  if (probe >= 0) { long long z[16] = {0}; cudaMemcpyToSymbol(bg::g_lg_t, z, sizeof(z)); return cudaMemcpyToSymbol(bg::g_lg_probe, &probe, sizeof(int)) == cudaSuccess ? 0 : -1; }
  return cudaMemcpyFromSymbol(out16, bg::g_lg_t, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
namespace bg {
#endif

template <class VT, class FT, int NR>
static int lattice_launch(bg_context_t ctx, LatticeGlobalDev& S, const void* in, void* tmp, void* out, const int* skip) {
  using CT = decltype(VT() * FT());
  const unsigned long long tmo = 5ull * 1000000000ull;
  auto kl = trsv_lattice_global<VT, CT, FT, CT, true, NR>;
  auto ku = trsv_lattice_global<CT, VT, FT, CT, false, NR>;
  const size_t sl = lg_smem_bytes<VT, CT, FT, true>(NR), su = lg_smem_bytes<CT, VT, FT, false>(NR);
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    BG_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl));
    BG_CUDA(cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)su));
    attr_done[ctx->device & 63] = true;
  }
  LgArgs<FT> al{S.lower.as<FT>(), nullptr, S.bnd.as<LL16>(), S.nx, S.ny, S.strips, ++S.seq, tmo};
  BG_CUDA(launch_pdl(kl, dim3(S.strips), dim3(kLgThreads), sl, ctx->stream, al, (const VT*)in, (CT*)tmp, skip));
  LgArgs<FT> au{S.upper.as<FT>(), S.udinv.as<FT>(), S.bnd.as<LL16>() + (size_t)S.strips * S.ny, S.nx, S.ny, S.strips, S.seq, tmo};
  BG_CUDA(launch_pdl(ku, dim3(S.strips), dim3(kLgThreads), su, ctx->stream, au, (const CT*)tmp, (VT*)out, skip));
  ctx->launches += 2;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}
// deep ring (one strip per SM) while the strips fit the SMs, a shallower one beyond (several strips per SM)
template <class VT, class FT>
static int lattice_run(bg_context_t ctx, LatticeGlobalDev& S, const void* in, void* tmp, void* out, const int* skip) {
  if (S.strips <= ctx->sm_count) return lattice_launch<VT, FT, 128>(ctx, S, in, tmp, out, skip);
  return lattice_launch<VT, FT, 64>(ctx, S, in, tmp, out, skip);
}

int launch_trsv_lattice_global(bg_context_t ctx, LatticeGlobalDev& S, int vec_dt, int fac_dt, const void* in, void* tmp,
                               void* out, const int* skip) {
  if (vec_dt == BG_F32 && fac_dt == BG_F32) return lattice_run<float, float>(ctx, S, in, tmp, out, skip);
  if (vec_dt == BG_F32 && fac_dt == BG_F64) return lattice_run<float, double>(ctx, S, in, tmp, out, skip);
  if (vec_dt == BG_F64 && fac_dt == BG_F32) return lattice_run<double, float>(ctx, S, in, tmp, out, skip);
  return lattice_run<double, double>(ctx, S, in, tmp, out, skip);
}
// strips that can be resident at once (one CTA per strip, shared memory bound): beyond that the general kernel runs
int lattice_global_capacity(bg_context_t ctx, int vec_dt, int fac_dt) {
  const size_t ct = (vec_dt == BG_F64 || fac_dt == BG_F64) ? 8 : 4, ft = dtype_size(fac_dt);
  const size_t su = (size_t)64 * 32 * (3 * ft + 2 * ct) + 64 * 8 + 64 + 1024;  // the shallow ring
  const int per_sm = (int)std::min<size_t>(16, (size_t)(227 * 1024) / su);
  return ctx->sm_count * std::max(per_sm, 1);
}

template <class VT, class FT, bool UNIT>
static int syncfree_run(bg_context_t ctx, SyncFreeDev& S, const void* in, void* xs, void* out, const TrsvXform& x,
                        const int* skip) {
  using CT = decltype(VT() * FT());
  SfArgs<FT> a{S.meta.as<int4>(), S.rows.as<int>(), S.cols.as<int>(), S.vals.as<FT>(), S.diag.as<FT>(), S.n_slices, ++S.seq,
               5ull * 1000000000ull};
  if ((x.pro == 1 && !S.src_idx.p) || (x.epi == 1 && !S.dst_idx.p)) { set_error("trsv_syncfree: permutation not folded into the stream"); return BG_ERR_ARG; }
  XformDev<VT> xf{x.pro, x.epi, x.pro == 1 ? S.src_idx.as<int>() : nullptr, x.epi == 1 ? S.dst_idx.as<int>() : nullptr,
                  (const VT*)x.scale_pro, (const VT*)x.scale_epi};
  // Resident warps.  The grid must be resident as a whole (warps wait for each other), hence the occupancy query.
  // Beyond that: deep enough to have the operands of the coming levels loaded before the dependency front arrives
  // and shallow enough that the warps polling ahead of the front do not flood L2.
  auto kern = trsv_syncfree<VT, FT, CT, UNIT>;
  static int per_sm_cache[64] = {};
  int& per_sm = per_sm_cache[ctx->device & 63];
  if (per_sm == 0) {
    BG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
    if (per_sm < 1) { set_error("trsv_syncfree: no resident block"); return BG_ERR_CUDA; }
  }
  static const int depth = [] { const char* e = getenv("BG_SF_DEPTH"); const int d = e ? atoi(e) : 16; return d > 0 ? d : 16; }();
  const double per_level = (double)S.n_slices / std::max(1, S.n_levels);
  const int max_warps = ctx->sm_count * std::min(per_sm, 8) * 4;
  int warps = (int)std::min<double>((double)max_warps, std::max<double>(ctx->sm_count, per_level * depth));
  warps = std::min(warps, std::max(1, S.n_slices));
  const int threads = warps <= 4 * ctx->sm_count ? 32 : 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  BG_CUDA(launch_pdl(kern, dim3(blocks), dim3(threads), 0, ctx->stream, a, xf, (const VT*)in, (LL16*)xs, (VT*)out, skip));
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

int launch_trsv_level(bg_context_t ctx, SyncFreeDev& S, int vec_dt, int fac_dt, bool unit, const void* in,
                      void* xs, void* out, const TrsvXform& xf, const int* skip) {
  if (S.n == 0) return BG_OK;
#define BG_L(VT, FT) (unit ? syncfree_run<VT, FT, true>(ctx, S, in, xs, out, xf, skip) : syncfree_run<VT, FT, false>(ctx, S, in, xs, out, xf, skip))
  if (vec_dt == BG_F32 && fac_dt == BG_F32) return BG_L(float, float);
  if (vec_dt == BG_F32 && fac_dt == BG_F64) return BG_L(float, double);
  if (vec_dt == BG_F64 && fac_dt == BG_F32) return BG_L(double, float);
  return BG_L(double, double);
#undef BG_L
}

}  // namespace bg
