// trsv_block.cu -- fused forward+backward sparse triangular solve of a block-Jacobi ILU, sm_100a.
//
// Performance-mode replacement of the two cusparse?csrsv_solve calls per preconditioner
// application (src/preconditioner.cu:263,273 / 1612,1637).  The global level schedule of the
// reference's factors (trsv.cu) has thousands of dependent levels on grid matrices and is bound
// by that chain, not by HBM; here the factors are block-diagonal (the north-star's per-partition
// block-Jacobi ILU applied inside one GPU as well), so every block is an independent problem:
//
//  * ONE warp solves one block: z = U_b^-1 L_b^-1 r_b, both solves back to back with the
//    intermediate vector never leaving shared memory (one kernel instead of two, no temporary
//    vector in HBM).
//  * rows are renumbered block-major on the host, so r_b and z_b are contiguous in HBM
//    (coalesced 8 KB reads/writes); the public entry point offers a gather/scatter variant for
//    callers whose vectors are in the original order.
//  * the factor stream of a block (schedule, 16-bit local columns, values, reciprocal diagonal;
//    packed step by step on the host, host_algos.h FusedBlockStream) is staged by bulk-async copies
//    (cp.async.bulk.shared::cluster.global + mbarrier complete_tx, SASS UBLKCP) into a per-warp
//    ring of `stages` slots, so HBM reads are large and sequential and run ahead of the dependent
//    FMA chain.  Every piece of per-step metadata lives in shared memory; the metadata of step s+1
//    is fetched into registers before the warp barrier of step s.
//  * 12 resident warps per SM (3 CTAs x 4 warps) give 1776 concurrent block solves on 148 SMs.
//
//  * both factors are rescaled on the host (Lh = D^-1 L D, Ut = D^-1 U, D = diag U) so that the
//    device performs two UNIT solves on d .* r: no division and no per-row diagonal in the stream.
//
// Bytes per row (5-point ILU(0), FP64): L 24 B + U 24 B + 1/diag 8 B + rhs 8 B + result 8 B.
//
//  * LATTICE layout (host_algos.h; chosen by the packer when every block is a tx x ty grid tile, tx <= 32, whose
//    factors couple only west/south and east/north neighbours, i.e. ILU(0) of a 5-point stencil): the stream
//    holds NO indices, two values per row per factor (32 B per row in FP64).  Lane x owns column x of the tile and
//    walks it along the anti-diagonal wavefront: its south (north) operand is its own previous result, kept in
//    a register, the west (east) one arrives from lane x-1 (x+1) by one warp shuffle.  A step is a shuffle and
//    two FMAs instead of ~40 instructions of index decoding, shared-memory gathers and a warp barrier, which
//    takes the block solve from latency-bound (~190 ns per step) to HBM-bound.
#include "internal.h"

namespace bg {

namespace {

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

#ifdef BG_TRSV_TIMING
// development aid: %globaltimer stamps of the warp that solves block `g_trsv_probe` (bg_debug_trsv_times)
__device__ unsigned long long g_trsv_t[96];
__device__ int g_trsv_probe = 0;
__device__ __forceinline__ unsigned long long gt_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef BG_TRSV_TIMING_CHUNKS
#define TSTAMP(k) do { if (b == g_trsv_probe && lane == 0 && (k) < 96) g_trsv_t[(k)] = gt_ns(); } while (0)
#else
#define TSTAMP(k) do { if ((k) < 8 && b == g_trsv_probe && lane == 0) g_trsv_t[(k)] = gt_ns(); } while (0)
#endif
#else
#define TSTAMP(k) do {} while (0)
#endif

struct FbArgs {
  const BlockDesc* blk;
  const uint32_t* chunk_tab;  // [nblocks][tab_stride]: chunk start offsets (16-byte units) of a block + end
  const unsigned char* stream;
  const int* perm;
  const void* dinv;     // FT[n], block-major (gather variant only: the solver folds it into A and b)
  int nblocks, max_rows, stages, tab_stride;
  uint32_t slot_bytes;  // ring slot stride (multiple of 128)
  uint32_t warp_smem;   // bytes of shared memory per warp (multiple of 128)
  uint32_t tab_bytes;   // LATTICE: zero page at the end of the warp's area (multiple of 128)
};

template <class FT> struct PairOf;
template <> struct PairOf<float> { using type = float2; };
template <> struct PairOf<double> { using type = double2; };

// metadata of one slot: row id + K <= 7 local columns in one or two 8-byte quads, K values
template <class FT, class CT, int K>
__device__ __forceinline__ void load_slot(int t, int nsA, const uint2* quads, const FT* vals, uint2& q0, uint2& q1, CT* v) {
  q0 = quads[t];
  if (K > 3) q1 = quads[nsA + t];
  using P = typename PairOf<FT>::type;
#pragma unroll
  for (int p = 0; p < K / 2; ++p) {
    const P pr = reinterpret_cast<const P*>(vals)[(size_t)p * nsA + t];
    v[2 * p] = (CT)pr.x;
    v[2 * p + 1] = (CT)pr.y;
  }
  if (K & 1) v[K - 1] = (CT)vals[(size_t)(K / 2) * 2 * nsA + t];
}

// all steps of one chunk whose rows have at most K <= 7 off-diagonals.  Branch-free: lanes beyond the
// step's row count read a (clamped) later slot and only their store is predicated off; the slot of
// step s+1 is fetched into registers before the warp barrier of step s.
template <class FT, class CT, int K>
__device__ __forceinline__ void run_steps_fast(int lane, int ns, int nst, int nsA, const unsigned char* cnts,
                                               const uint2* quads, const FT* vals, CT* xl) {
  uint2 qn0, qn1 = make_uint2(0u, 0u);
  CT vn[K > 0 ? K : 1];
  const int last = ns - 1;
  int s0 = 0, cnt = cnts[0];
  load_slot<FT, CT, K>(min(lane, last), nsA, quads, vals, qn0, qn1, vn);
  for (int s = 0; s < nst; ++s) {
    const uint2 q0 = qn0, q1 = qn1;
    CT v[K > 0 ? K : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = vn[k];
    const bool act = lane < cnt;
    s0 += cnt;
    cnt = cnts[s + 1];  // one readable byte past the last count (alignment padding)
    load_slot<FT, CT, K>(min(s0 + lane, last), nsA, quads, vals, qn0, qn1, vn);
    const uint32_t li = q0.x & 0xffffu;
    CT acc = xl[li];
    if (K > 0) acc -= v[0] * xl[q0.x >> 16];
    if (K > 1) acc -= v[1] * xl[q0.y & 0xffffu];
    if (K > 2) acc -= v[2] * xl[q0.y >> 16];
    if (K > 3) acc -= v[3] * xl[q1.x & 0xffffu];
    if (K > 4) acc -= v[4] * xl[q1.x >> 16];
    if (K > 5) acc -= v[5] * xl[q1.y & 0xffffu];
    if (K > 6) acc -= v[6] * xl[q1.y >> 16];
    if (act) xl[li] = acc;  // rows of one step never read each other
    __syncwarp();           // orders this step's writes before the next step's reads
  }
}

// any K: no register prefetch (wide ILUT rows)
template <class FT, class CT>
__device__ __forceinline__ void run_steps_generic(int lane, int K, int nst, int nsA, const unsigned char* cnts,
                                                  const uint2* quads, const FT* vals, CT* xl) {
  const uint16_t* q16 = reinterpret_cast<const uint16_t*>(quads);
  int s0 = 0;
  for (int s = 0; s < nst; ++s) {
    const int cnt = cnts[s];
    if (lane < cnt) {
      const int t = s0 + lane;
      const uint32_t li = q16[(size_t)t * 4];
      CT acc = xl[li];
      for (int k = 0; k < K; ++k) {
        const int e = k + 1;
        const uint32_t c = q16[((size_t)(e >> 2) * nsA + t) * 4 + (e & 3)];
        const FT v = (k == K - 1 && (K & 1)) ? vals[(size_t)(K / 2) * 2 * nsA + t] : vals[((size_t)(k >> 1) * nsA + t) * 2 + (k & 1)];
        acc -= (CT)v * xl[c];
      }
      xl[li] = acc;
    }
    __syncwarp();
    s0 += cnt;
  }
}

// one chunk of the LATTICE layout: whole anti-diagonals of a tx x ty tile, (west|east, south|north) value pairs.
// The wavefront step is the dependency chain of the whole solve (2 (tx + ty - 1) steps per tile, one warp), and a
// lone warp issues one instruction every few cycles, so the step is cut down to what the recurrence needs: every
// lane keeps RUNNING shared-memory addresses of its next right-hand side entry (+- tx elements per step) and of its
// next pair, and its tile row y (+- 1); "active" is one unsigned compare on y.  On anti-diagonal d the active lanes
// are [xlo_d, xlo_d + cnt_d) and lane x finds its pair at index u_d + x - xlo_d of the factor's stream (u_d = pairs
// of the earlier diagonals), so every lane's pair pointer advances by the same warp-uniform amount
// cnt_d - (xlo_next - xlo_d) = min(d, tx-1) + 1 - max(0, d_next - (ty-1)), three integer operations.  A lane outside
// the wavefront reads its operands from 16 zero bytes, so its z is 0 and only the store is predicated; a lane's
// neighbour is only consulted while that neighbour's value is the one the recurrence asks for (lane x is active on
// diagonal d exactly when lane x-1 was on d-1; the tile's first column / row carry zero coefficients).
// Per step: 2 SHFL, 2 DFMA, 2 LDS (right-hand side, pair), 1 STS, two selects, a compare and a few integer adds.
// The kernel is bound by the shared-memory pipe (ncu: 12 warps x ~10 wavefronts per step), not by issue or HBM.
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <class CT> __device__ __forceinline__ CT lds_ct(uint32_t a);
template <> __device__ __forceinline__ double lds_ct<double>(uint32_t a) { return lds_f64(a); }
template <> __device__ __forceinline__ float lds_ct<float>(uint32_t a) { return lds_f32(a); }
__device__ __forceinline__ void sts_ct(uint32_t a, double v) { sts_f64(a, v); }
__device__ __forceinline__ void sts_ct(uint32_t a, float v) { sts_f32(a, v); }
template <class CT, class FT> __device__ __forceinline__ void lds_pair(uint32_t a, CT& x, CT& y);
template <> __device__ __forceinline__ void lds_pair<double, double>(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
template <> __device__ __forceinline__ void lds_pair<double, float>(uint32_t a, double& x, double& y) {
  float fx, fy;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(fx), "=f"(fy) : "r"(a));
  x = (double)fx; y = (double)fy;
}
template <> __device__ __forceinline__ void lds_pair<float, float>(uint32_t a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}

template <class FT, class CT, bool LOWER>
__device__ __forceinline__ void lattice_steps(int lane, int nlev, int e0, const typename PairOf<FT>::type* __restrict__ pairs,
                                              int tx, int ty, CT* xl, CT& zprev_io, const int* zero16) {
  using P = typename PairOf<FT>::type;
  const int D = tx + ty - 2;
  const int d0 = LOWER ? e0 : D - e0;
  const uint32_t zero_a = smem_u32(zero16);
  const uint32_t uty = (uint32_t)ty;
  const int dxa = (LOWER ? tx : -tx) * (int)sizeof(CT), dy = LOWER ? 1 : -1;
  int y = lane < tx ? d0 - lane : -0x40000000;  // lanes beyond the tile never become active
  uint32_t pa = smem_u32(pairs) + (uint32_t)((lane - max(0, d0 - (ty - 1))) * (int)sizeof(P));
  uint32_t xa = smem_u32(xl) + (uint32_t)((lane + tx * (d0 - lane)) * (int)sizeof(CT));
  // operands of a step; a lane outside the wavefront reads zeros, so its z is 0 and nothing has to be predicated
  // but the store
  auto fetch = [&](bool act, CT& r, CT& ca, CT& cb) {
    r = lds_ct<CT>(act ? xa : zero_a);
    lds_pair<CT, FT>(act ? pa : zero_a, ca, cb);
  };
  CT z = zprev_io;
  auto chain = [&](bool act, uint32_t xs, CT r, CT ca, CT cb) {
    // neighbour lane's previous result (west / east), own previous result (south / north)
    const CT zn = LOWER ? __shfl_up_sync(0xffffffffu, z, 1) : __shfl_down_sync(0xffffffffu, z, 1);
    const CT t = r - cb * z;
    z = t - ca * zn;
    if (act) sts_ct(xs, z);
  };
  bool act = (uint32_t)y < uty;
  CT r, ca, cb;
  fetch(act, r, ca, cb);
  // warp-uniform, pre-scaled by the pair size: current diagonal, tx-1, and (ty-1) shifted by the step direction
  constexpr int PS = (int)sizeof(P);
  int dS = d0 * PS;
  const int tx1S = (tx - 1) * PS, offS = (LOWER ? ty - 2 : ty) * PS;  // d_next - (ty-1) = d - (ty-2) going up, d - ty going down
#pragma unroll 4
  for (int s = 0; s + 1 < nlev; ++s) {
    const uint32_t xs = xa;
    const bool act_s = act;
    // inc = min(d, tx-1) + 1 - max(0, d_next - (ty-1)),  d_next = d +- 1
    const int inc = min(dS, tx1S) + PS - max(0, dS - offS);
    dS += LOWER ? PS : -PS;
    pa += (uint32_t)inc; xa += (uint32_t)dxa; y += dy;
    act = (uint32_t)y < uty;
    CT r2, ca2, cb2;
    fetch(act, r2, ca2, cb2);  // next step's operands: in flight during this step's chain
    chain(act_s, xs, r, ca, cb);
    r = r2; ca = ca2; cb = cb2;
  }
  chain(act, xa, r, ca, cb);  // the successor of the chunk's last step sits in another ring slot
  zprev_io = z;
}
template <class FT, class CT>
__device__ __forceinline__ void run_lattice_chunk(int lane, const unsigned char* rec, int tx, int ty, CT* xl, CT& zprev, const int* zero16) {
  using P = typename PairOf<FT>::type;
  const uint2 h = *reinterpret_cast<const uint2*>(rec);
  const int nlev = (int)(h.x & 0xffffu), e0 = (int)(h.x >> 16);
  const P* pairs = reinterpret_cast<const P*>(rec + 16);
  if (e0 == 0) zprev = (CT)0;  // first chunk of a factor
  if (nlev < 1) return;
  if ((h.y & 0xffffu) != 0) lattice_steps<FT, CT, true>(lane, nlev, e0, pairs, tx, ty, xl, zprev, zero16);
  else lattice_steps<FT, CT, false>(lane, nlev, e0, pairs, tx, ty, xl, zprev, zero16);
}

// GATHER: vectors in the caller's original row order, r scaled by 1/diag(U) while loading.
// !GATHER: vectors in block-major order and already row-scaled (the solver runs on D^-1 P A P^T,
//          D^-1 P b), so a block's right-hand side / result is one bulk copy each way.
template <class VT, class FT, class CT, bool GATHER, bool LATTICE>
__global__ void __launch_bounds__(128) trsv_block_fused(FbArgs a, const VT* __restrict__ in, VT* __restrict__ out,
                                                        const int* skip) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char* mine = smem + (size_t)wid * a.warp_smem;
  CT* xl = reinterpret_cast<CT*>(mine);
  const uint32_t xl_bytes = ((uint32_t)a.max_rows * sizeof(CT) + 127u) & ~127u;
  uint32_t* coff = reinterpret_cast<uint32_t*>(mine + xl_bytes);
  const uint32_t coff_bytes = ((uint32_t)a.tab_stride * 4u + 127u) & ~127u;
  unsigned char* ring = mine + xl_bytes + coff_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * a.slot_bytes);
  uint64_t* bar_in = bars + a.stages;
  int* zero16 = reinterpret_cast<int*>(mine + a.warp_smem - a.tab_bytes);  // LATTICE: operands of lanes outside the wavefront
  if (LATTICE && lane < 4) zero16[lane] = 0;
  const FT* dinv = reinterpret_cast<const FT*>(a.dinv);
  constexpr bool kBulkVec = !GATHER && sizeof(VT) == sizeof(CT);  // rhs/result move as bulk copies
  if (lane == 0) {
    for (int s = 0; s <= a.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#ifdef BG_TRSV_TIMING
  if ((int)(blockIdx.x * nw + wid) == g_trsv_probe && lane == 0) g_trsv_t[7] = gt_ns();  // kernel entry of the probed warp
#endif
  // Launched with programmatic stream serialisation: until pdl_wait() below, only constant data (block
  // descriptors, chunk tables, the factor stream) is touched, so the first chunks of every warp are already on
  // their way while the SpMV that produces `in` is still running.
  bool waited = false;
  uint32_t st = 0;      // ring slot of the next chunk to consume (runs on across the blocks of this warp)
  uint32_t phases = 0;  // bit s: parity of the next completion of slot s
  uint32_t nin = 0;     // bulk loads of a right-hand side so far (phase of bar_in)
  const uint32_t nst = (uint32_t)a.stages;
  for (int b = blockIdx.x * nw + wid; b < a.nblocks; b += gridDim.x * nw) {
    // descriptor and chunk table are fetched in one round trip (the table has a fixed stride)
    TSTAMP(0);
    const int4 d0 = __ldg(reinterpret_cast<const int4*>(a.blk + b));
    for (int i = lane; i < a.tab_stride; i += 32) coff[i] = __ldg(a.chunk_tab + (size_t)b * a.tab_stride + i);
    const int row0 = d0.x, nb = d0.y, nch = d0.w;
    int tx = 0, ty = 0;
    if (LATTICE) {
      const int4 d1 = __ldg(reinterpret_cast<const int4*>(a.blk + b) + 1);
      tx = d1.z; ty = d1.w;
    }
    CT zprev = (CT)0;
    __syncwarp();
    TSTAMP(1);
    auto issue = [&](int c, uint32_t slot) {  // lane 0 only
      const uint32_t o0 = coff[c], bytes = (coff[c + 1] - o0) * 16u;
      mbar_expect_tx(&bars[slot], bytes);
      bulk_g2s(ring + (size_t)slot * a.slot_bytes, a.stream + (size_t)o0 * 16u, bytes, &bars[slot]);
    };
    const bool bulk_vec = kBulkVec && (((uint32_t)row0 * sizeof(VT)) & 15u) == 0 && (((uint32_t)nb * sizeof(VT)) & 15u) == 0 && nb > 0;
    if (lane == 0) {  // factor stream: constant data, may precede pdl_wait
      uint32_t sl = st;
      for (int c = 0; c < nch && c < a.stages; ++c) { issue(c, sl); sl = sl + 1 == nst ? 0u : sl + 1; }
    }
    if (!waited) {
      pdl_wait();
      waited = true;
      if (skip && *skip) {  // a skipped launch must still drain the bulk copies it started
        if (lane == 0) {
          uint32_t sl = st;
          for (int c = 0; c < nch && c < a.stages; ++c) { mbar_wait(&bars[sl], (phases >> sl) & 1u); sl = sl + 1 == nst ? 0u : sl + 1; }
        }
        return;
      }
    }
    if (lane == 0) {
      if (kBulkVec) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous result has left xl
      if (bulk_vec) {
        mbar_expect_tx(bar_in, (uint32_t)nb * sizeof(VT));
        bulk_g2s(xl, in + row0, (uint32_t)nb * sizeof(VT), bar_in);
      }
    }
    __syncwarp();
    TSTAMP(2);
    if (bulk_vec) {
      mbar_wait(bar_in, nin & 1u);
      ++nin;
      TSTAMP(3);
    } else {
      for (int i = lane; i < nb; i += 32) {
        if (GATHER) xl[i] = (CT)in[__ldg(a.perm + row0 + i)] * (CT)__ldg(dinv + row0 + i);
        else xl[i] = (CT)in[row0 + i];
      }
      __syncwarp();
    }
    for (int c = 0; c < nch; ++c, st = st + 1 == nst ? 0u : st + 1) {
      mbar_wait(&bars[st], (phases >> st) & 1u);
      phases ^= 1u << st;
      TSTAMP(8 + 2 * c);
      const unsigned char* rec = ring + (size_t)st * a.slot_bytes;
      if (LATTICE) {
        // a lane's next read of xl (its own column) follows its own write: no warp barrier per step; the
        // next chunk's first operands may have been written by this lane only
        run_lattice_chunk<FT, CT>(lane, rec, tx, ty, xl, zprev, zero16);
        __syncwarp();  // every lane is past its last read of this ring slot
        TSTAMP(9 + 2 * c);
        if (lane == 0 && c + a.stages < nch) issue(c + a.stages, st);
        continue;
      }
      const uint2 h = *reinterpret_cast<const uint2*>(rec);
      const int ns = (int)(h.x & 0xffffu), K = (int)(h.x >> 16), nst = (int)(h.y & 0xffffu), nsA = (int)(h.y >> 16);
      const unsigned char* cnts = rec + 16;
      const int stA = (nst + 16) & ~15;
      const uint2* quads = reinterpret_cast<const uint2*>(rec + 16 + stA);
      const FT* vals = reinterpret_cast<const FT*>(rec + 16 + stA + 8 * (size_t)((K + 4) >> 2) * nsA);
      switch (K) {  // uniform per chunk
        case 0: run_steps_fast<FT, CT, 0>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 1: run_steps_fast<FT, CT, 1>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 2: run_steps_fast<FT, CT, 2>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 3: run_steps_fast<FT, CT, 3>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 4: run_steps_fast<FT, CT, 4>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 5: run_steps_fast<FT, CT, 5>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 6: run_steps_fast<FT, CT, 6>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        case 7: run_steps_fast<FT, CT, 7>(lane, ns, nst, nsA, cnts, quads, vals, xl); break;
        default: run_steps_generic<FT, CT>(lane, K, nst, nsA, cnts, quads, vals, xl); break;
      }
      // every lane is past its last read of this ring slot: refill it
      TSTAMP(9 + 2 * c);
      if (lane == 0 && c + a.stages < nch) issue(c + a.stages, st);
    }
    TSTAMP(4);
    if (bulk_vec) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of xl -> async proxy
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + row0), "r"(smem_u32(xl)),
                     "r"((uint32_t)nb * (uint32_t)sizeof(VT))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      for (int i = lane; i < nb; i += 32) out[GATHER ? __ldg(a.perm + row0 + i) : row0 + i] = (VT)xl[i];
    }
    __syncwarp();
    TSTAMP(5);
  }
#ifdef BG_TRSV_TIMING
  if (blockIdx.x * nw + wid == g_trsv_probe && lane == 0) { if (kBulkVec) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); g_trsv_t[6] = gt_ns(); }
#endif
  if (!waited) pdl_wait();  // a warp without any block still orders itself behind the previous kernel
  if (kBulkVec && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // results are in HBM before exit
}

template <class VT, class FT>
int fused_run(bg_context_t ctx, const FusedBlockDev& S, bool gather, const void* in, void* out, const int* skip) {
  using CT = decltype(VT() * FT());
  FbArgs a{};
  a.blk = S.blk.as<BlockDesc>(); a.chunk_tab = S.chunk_tab.as<uint32_t>(); a.stream = S.stream.as<unsigned char>();
  a.perm = S.perm.as<int>();
  a.dinv = S.dinv.p;
  a.nblocks = S.nblocks; a.max_rows = S.max_rows; a.stages = S.stages; a.tab_stride = S.tab_stride;
  a.slot_bytes = ((uint32_t)S.max_chunk_bytes + 127u) & ~127u;
  const uint32_t xl_bytes = ((uint32_t)S.max_rows * sizeof(CT) + 127u) & ~127u;
  const uint32_t coff_bytes = ((uint32_t)S.tab_stride * 4u + 127u) & ~127u;
  a.tab_bytes = S.lattice ? 128u : 0u;
  a.warp_smem = xl_bytes + coff_bytes + (uint32_t)a.stages * a.slot_bytes + 128u + a.tab_bytes;
  const int warps = S.warps_per_cta;
  const size_t smem = (size_t)warps * a.warp_smem;
  if (smem > ctx->smem_optin) {
    set_error("block TRSV needs %zu bytes of shared memory per CTA (limit %zu): use smaller blocks", smem, ctx->smem_optin);
    return BG_ERR_UNSUPPORTED;
  }
  auto kern = S.lattice ? (gather ? trsv_block_fused<VT, FT, CT, true, true> : trsv_block_fused<VT, FT, CT, false, true>)
                        : (gather ? trsv_block_fused<VT, FT, CT, true, false> : trsv_block_fused<VT, FT, CT, false, false>);
  static size_t attr_set[64][4] = {};  // per template instantiation and per device (the attribute is per device)
  const int ai = (gather ? 1 : 0) + (S.lattice ? 2 : 0), dv = ctx->device & 63;
  if (attr_set[dv][ai] < smem) {
    BG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    attr_set[dv][ai] = ctx->smem_optin;
  }
  // resident CTAs per SM: 228 KB per SM, 1 KB reserved per CTA
  int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(228 * 1024) / (smem + 1024)));
  ctas_per_sm = std::min(ctas_per_sm, 2048 / (warps * 32));
  const int blocks = std::max(1, std::min((S.nblocks + warps - 1) / warps, ctx->sm_count * ctas_per_sm));
  BG_CUDA(launch_pdl(kern, dim3(blocks), dim3(warps * 32), smem, ctx->stream, a, (const VT*)in, (VT*)out, skip));
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  return BG_OK;
}

}  // namespace

#ifdef BG_TRSV_TIMING
extern "C" BG_API int bg_debug_trsv_times(int probe_block, unsigned long long* out96) {
  if (probe_block >= 0) return cudaMemcpyToSymbol(g_trsv_probe, &probe_block, sizeof(int)) == cudaSuccess ? 0 : -1;
  return cudaMemcpyFromSymbol(out96, g_trsv_t, sizeof(unsigned long long) * 96) == cudaSuccess ? 0 : -1;
}
#endif

int launch_trsv_block(bg_context_t ctx, const FusedBlockDev& S, int vec_dt, int fac_dt, bool gather, const void* in,
                      void* out, const int* skip) {
  if (S.n == 0) return BG_OK;
  if (vec_dt == BG_F32 && fac_dt == BG_F32) return fused_run<float, float>(ctx, S, gather, in, out, skip);
  if (vec_dt == BG_F32 && fac_dt == BG_F64) return fused_run<float, double>(ctx, S, gather, in, out, skip);
  if (vec_dt == BG_F64 && fac_dt == BG_F32) return fused_run<double, float>(ctx, S, gather, in, out, skip);
  return fused_run<double, double>(ctx, S, gather, in, out, skip);
}

}  // namespace bg
