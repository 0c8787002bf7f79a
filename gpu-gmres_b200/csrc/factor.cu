// factor.cu -- ILU(0) factorisation ON THE DEVICE for 5-point grids, sm_100a.
//
// Replaces, for the block-Jacobi tiles of a 2-D grid, the host pipeline  leftILU  ->  splitLU_csr  ->  stream
// packing  (reference: src/leftILU.cu:219-298 level loop + :649-702 sparseTriSolve_V2; here bg_host_ilu0 +
// pack_fused_block_stream, 2.9 s at 10 M rows, 98 % of a one-shot DC solve) by one kernel that produces the packed
// LATTICE stream of trsv_block_fused directly in HBM.
//
// ILU(0) of a 5-point stencil has no fill and only the diagonal changes:
//     l_s = a_s / d_south,  l_w = a_w / d_west,   d = (a_d - l_s * a_n(south)) - l_w * a_e(west)
// i.e. the same anti-diagonal wavefront as the triangular solve: one warp per tile, lane x owns column x, its south
// operand is its own previous node, its west operand arrives by shuffle.  The operations, their order and their
// rounding (separate multiply and subtract, no FMA: the host code is compiled with -ffp-contract=off) are exactly
// those of bg_host_ilu0 (row-wise IKJ, lower entries in ascending column order: south, then west), and the scaling
// of the stream values repeats pack_fused_block_stream's expressions, so the device-built preconditioner is
// bit-identical to the host-built one (tests/test_gpu_parity.py::test_device_factorisation_equals_host).
#include <algorithm>
#include <cstring>

#include "internal.h"

namespace bg {

namespace {

constexpr int kMaxDiag = 64;     // anti-diagonals of a tile: tx + ty - 1 <= 63 for tiles up to 32 x 32
constexpr int kMaxShapeChunks = 32;
struct TileShape {               // layout of one tile's stream region; identical for every tile of the same size
  int tx, ty, nchunks, bytes;    // bytes of the region (L chunks then U chunks)
  int offL[kMaxDiag], offU[kMaxDiag];        // byte offset of the first pair of anti-diagonal dd
  int hdr_off[kMaxShapeChunks];              // chunk headers: offset, {nlev, e0, lower}
  unsigned short hdr_nlev[kMaxShapeChunks], hdr_e0[kMaxShapeChunks], hdr_lower[kMaxShapeChunks];
};
struct TileArgs {
  const int *rp, *ci;
  const void* val;
  int val_f64;
  int nx, ny, tile_x, tile_y, tiles_x, tiles_y;
  const TileShape* shapes;       // [4]: (full|edge in x) + 2 (full|edge in y)
  const unsigned long long* tile_base;  // [ntiles] byte offset of the tile's region in the stream
  const int* tile_row0;          // [ntiles] first block-major row
  unsigned char* stream;
  void* dinv;                    // T[n], block-major
  int* status;                   // [0] pattern mismatch  [1] vanishing pivot  [2..3] nnzL-n, nnzU-n counters
};

template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <class T> __device__ __forceinline__ T sub_rn(T a, T b);
template <> __device__ __forceinline__ double sub_rn<double>(double a, double b) { return __dsub_rn(a, b); }
template <> __device__ __forceinline__ float sub_rn<float>(float a, float b) { return __fsub_rn(a, b); }
template <class T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <class T> __device__ __forceinline__ bool tiny(T v) { return (v < 0 ? -v : v) < (T)kEps; }

// one warp per tile
template <class T>
__global__ void __launch_bounds__(128) ilu0_lattice_tiles(TileArgs a) {
  const int lane = threadIdx.x & 31;
  const int tile = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (tile >= a.tiles_x * a.tiles_y) return;
  const int bx = tile % a.tiles_x, by = tile / a.tiles_x;
  const int X0 = bx * a.tile_x, Y0 = by * a.tile_y;
  const int tx = min(a.tile_x, a.nx - X0), ty = min(a.tile_y, a.ny - Y0);
  const TileShape& sh = a.shapes[(tx == a.tile_x ? 0 : 1) + (ty == a.tile_y ? 0 : 2)];
  unsigned char* reg = a.stream + a.tile_base[tile];
  T* dinv = reinterpret_cast<T*>(a.dinv) + a.tile_row0[tile];
  for (int c = lane; c < sh.nchunks; c += 32) {  // chunk headers
    unsigned short* h = reinterpret_cast<unsigned short*>(reg + sh.hdr_off[c]);
    h[0] = sh.hdr_nlev[c]; h[1] = sh.hdr_e0[c]; h[2] = sh.hdr_lower[c]; h[3] = 0;
  }
  const int x = lane, nx = a.nx;
  T d_prev = (T)1, an_prev = (T)0, ae_prev = (T)0;  // this lane's previous node (south of the current one)
  int bad = 0, zero_piv = 0, cntL = 0, cntU = 0;
  const int D = tx + ty - 1;
  for (int dd = 0; dd < D; ++dd) {
    const int y = dd - x;
    const bool act = x < tx && y >= 0 && y < ty;
    T a_w = 0, a_s = 0, a_d = 0, a_e = 0, a_n = 0;
    bool has_d = false;
    if (act) {
      const int r = (Y0 + y) * nx + X0 + x;
      const int p0 = a.rp[r], p1 = a.rp[r + 1];
      for (int p = p0; p < p1; ++p) {
        const int c = a.ci[p];
        const T v = a.val_f64 ? (T)reinterpret_cast<const double*>(a.val)[p] : (T)reinterpret_cast<const float*>(a.val)[p];
        if (c == r) { a_d = v; has_d = true; }
        else if (c == r - 1 && X0 + x > 0) { if (x > 0) a_w = v; }           // west: dropped when it leaves the tile
        else if (c == r + 1 && X0 + x < nx - 1) { if (x < tx - 1) a_e = v; }
        else if (c == r - nx) { if (y > 0) a_s = v; }
        else if (c == r + nx) { if (y < ty - 1) a_n = v; }
        else bad = 1;  // not a 5-point row of this grid
      }
      if (!has_d) bad = 1;
    }
    // west neighbour's pivot and east coefficient: lane x-1 computed them in the previous step
    const T d_w = __shfl_up_sync(0xffffffffu, d_prev, 1), ae_w = __shfl_up_sync(0xffffffffu, ae_prev, 1);
    if (act) {
      const T d_s = d_prev;
      // bg_host_ilu0 (host_algos.cpp ilu0): lik = a_ik / u_kk (0 for a vanishing pivot), a_ii -= lik * u_ki, k ascending
      T l_s = (T)0, l_w = (T)0, d = a_d;
      if (y > 0 && a_s != (T)0) { l_s = tiny(d_s) ? (T)0 : div_rn(a_s, d_s); d = sub_rn(d, mul_rn(l_s, an_prev)); }
      if (x > 0 && a_w != (T)0) { l_w = tiny(d_w) ? (T)0 : div_rn(a_w, d_w); d = sub_rn(d, mul_rn(l_w, ae_w)); }
      if (tiny(d)) zero_piv = 1;
      // splitLU_csr drops |v| < 1e-9 entries of the factored matrix from the pattern
      const bool kw = x > 0 && a_w != (T)0 && !tiny(l_w), ks = y > 0 && a_s != (T)0 && !tiny(l_s);
      const bool ke = a_e != (T)0 && !tiny(a_e), kn = a_n != (T)0 && !tiny(a_n);
      cntL += (kw ? 1 : 0) + (ks ? 1 : 0);
      cntU += (ke ? 1 : 0) + (kn ? 1 : 0);
      // pack_fused_block_stream: Lh = D^-1 L D, Ut = D^-1 U, dinv = 1/d, all through double
      const double dr = (double)d, di = 1.0 / dr;
      const int xlo = max(0, dd - (ty - 1));
      T* pl = reinterpret_cast<T*>(reg + sh.offL[dd]) + 2 * (x - xlo);
      T* pu = reinterpret_cast<T*>(reg + sh.offU[dd]) + 2 * (x - xlo);
      pl[0] = kw ? (T)((double)l_w * di * (double)d_w) : (T)0;
      pl[1] = ks ? (T)((double)l_s * di * (double)d_s) : (T)0;
      pu[0] = ke ? (T)((double)a_e * di) : (T)0;
      pu[1] = kn ? (T)((double)a_n * di) : (T)0;
      dinv[x + tx * y] = (T)di;
      d_prev = d; an_prev = a_n; ae_prev = a_e;  // the elimination uses every stored in-tile entry, tiny or not
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  zero_piv = __any_sync(0xffffffffu, zero_piv);
  for (int o = 16; o > 0; o >>= 1) { cntL += __shfl_xor_sync(0xffffffffu, cntL, o); cntU += __shfl_xor_sync(0xffffffffu, cntU, o); }
  if (lane == 0) {
    if (bad) atomicExch(&a.status[0], 1);
    if (zero_piv) atomicExch(&a.status[1], 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.status + 2), (unsigned long long)cntL);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.status + 4), (unsigned long long)cntU);
  }
}

// stream layout of a tx x ty tile: the chunking rule of pack_fused_block_stream's emit_lattice
template <class T> static void shape_layout(int tx, int ty, int chunk_target, TileShape& s, std::vector<int>& chunk_bytes) {
  std::memset(&s, 0, sizeof(s));
  s.tx = tx; s.ty = ty;
  chunk_bytes.clear();
  const int D = tx + ty - 1;
  int pos = 0;
  for (int lower = 1; lower >= 0; --lower) {
    int e = 0;
    while (e < D) {
      int nlev = 0, rows = 0;
      const int e0 = e;
      const int c = (int)chunk_bytes.size();
      s.hdr_off[c] = pos; s.hdr_e0[c] = (unsigned short)e0; s.hdr_lower[c] = (unsigned short)lower;
      while (e < D) {
        const int dd = lower ? e : D - 1 - e;
        const int sz = std::min(dd, tx - 1) - std::max(0, dd - (ty - 1)) + 1;
        if (nlev > 0 && 16 + (rows + sz) * 2 * (int)sizeof(T) > chunk_target) break;
        (lower ? s.offL : s.offU)[dd] = pos + 16 + rows * 2 * (int)sizeof(T);
        rows += sz; ++nlev; ++e;
      }
      s.hdr_nlev[c] = (unsigned short)nlev;
      const int bytes = (16 + rows * 2 * (int)sizeof(T) + 15) / 16 * 16;
      chunk_bytes.push_back(bytes);
      pos += bytes;
    }
  }
  s.nchunks = (int)chunk_bytes.size();
  s.bytes = pos;
}

}  // namespace

// Builds M->fb (the device copy of a FusedBlockStream in LATTICE layout) for the tiles dims = {nx, ny, 1, tile_x,
// tile_y, 1} of a 5-point grid matrix given as host CSR.  Returns BG_OK, a negative error, or 1 when the matrix is
// not such a grid / a pivot vanishes (the caller then takes the host path, which handles every case).
template <class T>
int lattice_tiles_factor_device(bg_context_t ctx, bg_precond_t M, int n, const int* rp, const int* ci, const void* v, int vdt,
                                const int* dims, uint64_t serial) {
  const int nx = dims[0], ny = dims[1], tile_x = dims[3], tile_y = dims[4];
  if (nx < 1 || ny < 1 || (int64_t)nx * ny != n || tile_x < 1 || tile_x > 32 || tile_y < 1 || tile_x + tile_y - 1 > kMaxDiag - 1) return 1;
  if ((int64_t)rp[n] > 5LL * n) return 1;
  cudaStream_t st = ctx->stream;
  const int tiles_x = (nx + tile_x - 1) / tile_x, tiles_y = (ny + tile_y - 1) / tile_y, ntiles = tiles_x * tiles_y;
  // ---- layout (host, integer work only)
  TileShape shapes[4];
  std::vector<int> cb[4];
  const int ex = nx - (tiles_x - 1) * tile_x, ey = ny - (tiles_y - 1) * tile_y;  // sizes of the last tile column / row
  for (int k = 0; k < 4; ++k) {
    shape_layout<T>((k & 1) ? ex : tile_x, (k & 2) ? ey : tile_y, std::max(512, M->fb_chunk_bytes), shapes[k], cb[k]);
    if (shapes[k].nchunks > kMaxShapeChunks) return 1;
  }
  FusedBlockDev& d = M->fb;
  std::vector<BlockDesc> blk(ntiles);
  std::vector<unsigned long long> tile_base(ntiles);
  std::vector<int> tile_row0(ntiles);
  std::vector<uint32_t> chunk_off(1, 0u);
  d.max_rows = d.max_levels_L = d.max_levels_U = d.max_chunk_bytes = d.max_chunks = 0;
  d.total_steps = 0;
  int row0 = 0;
  unsigned long long pos = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int bx = t % tiles_x, by = t / tiles_x;
    const int k = (bx == tiles_x - 1 && ex != tile_x ? 1 : 0) + (by == tiles_y - 1 && ey != tile_y ? 2 : 0);
    const TileShape& s = shapes[k];
    BlockDesc& b = blk[t];
    b.row0 = row0; b.nrows = s.tx * s.ty;
    b.chunk0 = (int)chunk_off.size() - 1; b.nchunks = s.nchunks;
    b.levelsL = b.levelsU = s.tx + s.ty - 1;
    b.pad0 = s.tx; b.pad1 = s.ty;
    tile_base[t] = pos; tile_row0[t] = row0;
    for (int c = 0; c < s.nchunks; ++c) {
      pos += (unsigned long long)cb[k][c];
      chunk_off.push_back((uint32_t)(pos / 16));
      d.max_chunk_bytes = std::max(d.max_chunk_bytes, cb[k][c]);
    }
    row0 += b.nrows;
    d.max_rows = std::max(d.max_rows, b.nrows);
    d.max_levels_L = std::max(d.max_levels_L, b.levelsL);
    d.max_levels_U = d.max_levels_L;
    d.max_chunks = std::max(d.max_chunks, b.nchunks);
    d.total_steps += 2 * (int64_t)(s.tx + s.ty - 1);
    if (pos / 16 > 0xFFFFFFF0ull) return 1;
  }
  const size_t stream_bytes = (size_t)pos + 16;
  // block-major order: tiles row by row, nodes of a tile row by row (ascending old id inside a block)
  d.h_perm.resize(n);
  std::vector<int> inv(n);
  {
    int q = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int X0 = (t % tiles_x) * tile_x, Y0 = (t / tiles_x) * tile_y, tx = blk[t].pad0, ty = blk[t].pad1;
      for (int y = 0; y < ty; ++y)
        for (int x = 0; x < tx; ++x) { const int r = (Y0 + y) * nx + X0 + x; d.h_perm[q] = r; inv[r] = q++; }
    }
  }
  d.n = n; d.nblocks = ntiles; d.lattice = true; d.serial = serial;
  d.tab_stride = (d.max_chunks + 1 + 3) / 4 * 4;
  std::vector<uint32_t> tab((size_t)ntiles * d.tab_stride, 0u);
  for (int t = 0; t < ntiles; ++t)
    for (int c = 0; c <= blk[t].nchunks; ++c) tab[(size_t)t * d.tab_stride + c] = chunk_off[blk[t].chunk0 + c];
  // ---- device: matrix in, stream out
  DevBuf d_rp, d_ci, d_val, d_shapes, d_base, d_row0, d_status;
  BG_TRY(d_rp.upload(rp, sizeof(int) * ((size_t)n + 1), st));
  BG_TRY(d_ci.upload(ci, sizeof(int) * (size_t)rp[n], st));
  BG_TRY(d_val.upload(v, dtype_size(vdt) * (size_t)rp[n], st));
  BG_TRY(d_shapes.upload(shapes, sizeof(shapes), st));
  BG_TRY(d_base.upload(tile_base.data(), sizeof(unsigned long long) * ntiles, st));
  BG_TRY(d_row0.upload(tile_row0.data(), sizeof(int) * ntiles, st));
  BG_TRY(d_status.alloc(32));
  BG_CUDA(cudaMemsetAsync(d_status.p, 0, 32, st));
  BG_TRY(d.stream.alloc(stream_bytes));
  BG_CUDA(cudaMemsetAsync(d.stream.p, 0, stream_bytes, st));
  BG_TRY(d.dinv.alloc(sizeof(T) * (size_t)n));
  TileArgs a{d_rp.as<int>(), d_ci.as<int>(), d_val.p, vdt == BG_F64 ? 1 : 0, nx, ny, tile_x, tile_y, tiles_x, tiles_y,
             d_shapes.as<TileShape>(), d_base.as<unsigned long long>(), d_row0.as<int>(), d.stream.as<unsigned char>(), d.dinv.p,
             d_status.as<int>()};
  ilu0_lattice_tiles<T><<<(ntiles + 3) / 4, 128, 0, st>>>(a);
  ctx->launches++;
  BG_CUDA(cudaGetLastError());
  BG_TRY(d.perm.upload(d.h_perm.data(), sizeof(int) * (size_t)n, st));
  BG_TRY(d.iperm.upload(inv.data(), sizeof(int) * (size_t)n, st));
  BG_TRY(d.blk.upload(blk.data(), sizeof(BlockDesc) * blk.size(), st));
  BG_TRY(d.chunk_tab.upload(tab.data(), sizeof(uint32_t) * tab.size(), st));
  int status[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  BG_CUDA(cudaMemcpyAsync(status, d_status.p, 32, cudaMemcpyDeviceToHost, st));
  std::vector<T> hd(n);
  BG_CUDA(cudaMemcpyAsync(hd.data(), d.dinv.p, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, st));
  BG_CUDA(cudaStreamSynchronize(st));
  if (status[0] || status[1]) {  // not a 5-point grid in this numbering, or a vanishing pivot: the host path decides
    for (DevBuf* b : {&d.stream, &d.dinv, &d.perm, &d.iperm, &d.blk, &d.chunk_tab}) b->release();
    d.h_perm.clear(); d.h_dinv.clear(); d.lattice = false; d.nblocks = 0; d.n = 0;
    return 1;
  }
  d.h_dinv.assign(hd.begin(), hd.end());
  unsigned long long cl, cu;
  std::memcpy(&cl, status + 2, 8);
  std::memcpy(&cu, status + 4, 8);
  M->nnzL = (int64_t)cl + n;
  M->nnzU = (int64_t)cu + n;
  M->unit_lower = true;
  d.bytes = (int64_t)stream_bytes + (int64_t)blk.size() * sizeof(BlockDesc) + (int64_t)tab.size() * 4;
  return BG_OK;
}
template int lattice_tiles_factor_device<float>(bg_context_t, bg_precond_t, int, const int*, const int*, const void*, int, const int*, uint64_t);
template int lattice_tiles_factor_device<double>(bg_context_t, bg_precond_t, int, const int*, const int*, const void*, int, const int*, uint64_t);

}  // namespace bg
