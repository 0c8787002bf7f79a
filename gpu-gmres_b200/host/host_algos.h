// host_algos.h -- host-side setup for the B200 GMRES(m)+ILU path: ILU(0)/ILUT factorisation,
// level sets, row/block partitioning, and the packers that turn CSR factors into the streams the
// triangular-solve kernels consume.  Pure C++ (no CUDA): linked into libb200gmres.so and usable
// without a GPU (bg_host_* entry points).
#pragma once
#include <cstdint>
#include <vector>

#include <algorithm>
#include <thread>

namespace bg {

// Setup of 10 M-row problems is bound by first-touch page faults and plain copies of several hundred megabytes:
// independent index ranges are handed to a few host threads.  fn(begin, end) must only write its own range.
template <class F> inline void parallel_ranges(size_t n, F fn, size_t min_parallel = (size_t)1 << 20) {
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t nt = n < min_parallel ? 1 : std::min<size_t>(8, hw ? hw : 1);
  if (nt <= 1) { fn((size_t)0, n); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + nt - 1) / nt;
  for (size_t t = 0; t < nt; ++t) {
    const size_t b = t * chunk, e = std::min(n, b + chunk);
    if (b < e) th.emplace_back([=] { fn(b, e); });
  }
  for (auto& x : th) x.join();
}


// |v| < 1e-9 counts as a structural zero (reference: src/defs.h:45-47)
constexpr float kEps = 1e-9f;
template <class T> inline bool is_zero(T v) { return (v < 0 ? -v : v) < (T)kEps; }

template <class T> struct Csr {
  int n = 0;
  std::vector<int> rp, ci;
  std::vector<T> v;
};

// L: strictly lower entries (ascending column) then the diagonal LAST; U: diagonal FIRST then
// ascending columns (reference layout, src/leftILU.cu:481-541).
template <class T> struct Factors {
  Csr<T> L, U;
  int zero_pivots = 0;
  bool unit_lower = true;
};

// generateLevel, src/leftILU.cu:339-368
void factor_levels(int n, const int* rp, const int* ci, const float* v, int* level);
// getLevel, src/gmres.cu:1033-1077
template <class T> void trsv_levels(bool lower, int n, const int* rp, const int* ci, const T* v, int* level);
// bucketing of src/leftILU.cu:41-49,188-192; returns number of levels
int bucket_levels(int n, const int* level, std::vector<int>& levelPtr, std::vector<int>& nodes);

// ILU(0), row-wise IKJ.  Performs, for every entry, the same subtractions in the same order as the
// reference's left-looking column sweep (cpuSequentialTriSolve, src/leftILU.cu:769-825), so in
// the same precision the factors are bit-identical.  block (may be null): keep only entries whose
// row and column share a block (block-Jacobi ILU).
template <class T>
void ilu0(int n, const int* rp, const int* ci, const T* v, const int* block, Factors<T>& F);
// ILU(k), level-of-fill rule of ITSOL's lofC (src/iluk.cpp:193-335, vendored by the reference); lfil = 0 is ilu0
template <class T>
void iluk(int n, const int* rp, const int* ci, const T* v, int lfil, const int* block, Factors<T>& F);
// ILUT(tau, lfil): dual-threshold incomplete LU (Saad), same output layout, L unit lower.
template <class T>
void ilut(int n, const int* rp, const int* ci, const T* v, double tau, int lfil, const int* block,
          Factors<T>& F);

// Sparse approximate inverse by left-looking biconjugation with dropping (Benzi-Tuma AINV; the role of CUSP's
// nonsym_bridson_ainv(A, droptol) in MyAINV::Initilize, src/preconditioner.cu:140-201):  A^-1 ~ Z D^-1 W^T.
// Output in the factor container: F.L = D^-1 W^T (CSR, applied first), F.U = Z (CSR, applied second).
// Entries of z_i / w_i below droptol in magnitude are dropped as they are created (the unit diagonal is kept).
template <class T>
void ainv_biconjugation(int n, const int* rp, const int* ci, const T* v, double droptol, Factors<T>& F);

// partition / partition3 post-processing, src/partition.cpp:22-99, src/partition3.cpp:68-117
void partition_ref(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, int64_t nNodes,
                   int64_t* part_size, int64_t* node_part, int64_t* pinv, int64_t* q);
void partition3_post(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, const int64_t* part,
                     int64_t* part_size, int64_t* node_part, int64_t* pinv, int64_t* q);
// one block per row, no separator.  method: 0 contiguous, 1 METIS k-way on the off-diagonal
// pattern (graph built as src/partition3.cpp:49-58), 2/3 structured tiles from dims.
// Returns the number of blocks actually produced (<= nblocks for tiles it is the tile count).
int block_partition(int method, int n, const int* rp, const int* ci, int nblocks, const int* dims,
                    int* block_of_row);

// ------------------------------------------------------------------ device stream packers
// Structural triangular-solve levels: like trsv_levels but over EVERY stored off-diagonal entry, whatever its
// value (getLevel stops at / skips |v| < 1e-9, so a factor holding an explicit tiny entry would be scheduled
// below a row it still reads).  This is the dependency spec the device schedules execute; trsv_levels stays
// the parity artefact.  The two agree whenever the factor stores no |v| < 1e-9 entry (always true for
// bg_host_ilu0 / splitLU_csr output).
void trsv_levels_structural(bool lower, int n, const int* rp, const int* ci, int* level);

// Sync-free stream (BG_TRSV_LEVEL, general factors; trsv.cu trsv_syncfree): rows sorted by (structural level,
// number of 4-entry groups descending, row id) -- slot s holds row rows[s] -- and cut into SLICES of <= 32 consecutive slots that never span two
// levels, so a slice (= one warp) only ever waits for rows of earlier slices.  Slice q is described by
// meta[4q..4q+3] = { first slot, R | K << 8, level, off32 }: R rows, K entries per row, and off-diagonal entry k of
// its t-th row sits at cols/vals[(off32 + k) * 32 + t] (column -1 = padding; the lane stride is always 32, so every
// group of a slice is one aligned 128/256-byte line).  Rows shorter than K are RIGHT-aligned (padding first) and a
// row's entries are stored by ascending (level of the operand's row, column): the operands that arrive last sit in
// the LAST entries of every row of the slice, so a warp that runs ahead of the dependency front has consumed
// everything else when they appear and the level's critical path is one group, not K/4.
// diag[s] per slot (1 for unit-lower).  level_slices[l] = slices of level l.
template <class T> struct SyncFreeStream {
  int n = 0, n_levels = 0, n_slices = 0, max_K = 0, max_level_rows = 0;
  std::vector<int> level;        // per row: getLevel (parity API)
  std::vector<int> sched_level;  // per row: structural level the slices follow
  std::vector<int> meta;         // 4 * (n_slices + 1); the extra entry is a copy of the last slice (prefetch past the end)
  std::vector<int> level_slices; // n_levels
  std::vector<int> rows, cols;
  std::vector<T> vals, diag;
};
// lower: diagonal is the LAST entry of each row (unit => value ignored); upper: FIRST entry.
template <class T>
void pack_syncfree_stream(bool lower, bool unit, const Csr<T>& M, SyncFreeStream<T>& out);

// Global lattice stream (BG_TRSV_LEVEL, ILU(0) factors of a 5-point stencil in natural order; trsv.cu
// trsv_lattice_global): the factors of an nx x ny grid (row = y*nx + x) that couple only west/south (L, unit
// lower) and east/north (U) neighbours are solved EXACTLY (global operator, not block-Jacobi) by one warp per
// 32-column STRIP walking the anti-diagonal wavefront: lane x of strip s owns column 32 s + x, its south (north)
// operand is its own previous result, its west (east) operand comes from the neighbour lane by warp shuffle,
// and the first lane's comes from the neighbouring strip through a small boundary buffer in L2.
// No index is stored.  Per strip s the stream holds the coefficients of its 32 columns row by row,
//     lower[((s*ny + y)*32 + x)*2 + {0,1}] = (west, south) coefficient of node (32 s + x, y)
//     upper[((s*ny + y)*32 + x)*2 + {0,1}] = (east, north) coefficient, pre-scaled by dinv = 1/diag(U)
//     udinv[(s*ny + y)*32 + x]             = dinv                      (the device solves a unit upper system)
// so a warp stages whole 32-node row segments with coalesced copies and each lane consumes its own column along
// the wavefront (lane x reaches row y x steps after lane 0).  Columns outside the grid hold zeros.
template <class T> struct LatticeGlobal {
  int nx = 0, ny = 0, strips = 0;
  std::vector<T> lower, upper, udinv;
};
// true (and nx set) when L (unit lower, diagonal last) and U (diagonal first) have the lattice pattern above
template <class T> bool detect_lattice2d(const Csr<T>& L, const Csr<T>& U, int* nx);
template <class T> void pack_lattice_global(const Csr<T>& L, const Csr<T>& U, int nx, LatticeGlobal<T>& out);

// Fused block stream (BG_TRSV_BLOCK): every diagonal block's forward AND backward solve is done
// by one warp out of shared memory in one kernel.  Rows are renumbered block-major
// (new row i = old row perm[i]; a block owns the contiguous new rows [blk[b].row0, +nrows)), so the
// right-hand side and the result of a block are contiguous in HBM.
// A block's work is a sequence of STEPS (<= 32 mutually independent rows = one level, or a 32-row
// slice of a wide level), L steps first, then U steps.  Both factors are rescaled on the host so that
// the device only ever performs UNIT triangular solves:  with D = diag(U), d = 1/D,
//     z = U^-1 L^-1 r  =  Ut^-1 Lh^-1 (d .* r),   Lh = D^-1 L D (unit lower),  Ut = D^-1 U (unit upper)
// and the kernel multiplies by dinv (block-major, n entries) while it loads r.
// Steps are packed, whole, into CHUNKS of <= chunk_target bytes; one chunk = one bulk-async copy:
//     [ u16 ns, K, nsteps, nsA, pad[4] ]                               16 B header
//     [ u8  step_cnt[nstepsA] ]          nstepsA = nsteps+1 rounded up to 16 (one readable byte past the end)
//     [ u16x4 quad[KQ][nsA] ]            quad 0 = (row, c0, c1, c2), quad q>0 = (c[4q-1] .. c[4q+2]);
//                                        KQ = (K+4)/4, local ids; padding columns = the row itself
//     [ T x2  pair[K/2][nsA] ]           values (v0,v1), (v2,v3), ...
//     [ T     last[nsA] ]                only when K is odd
// nsA = ns rounded up to 8.  Chunk c of the whole stream starts at byte 16*chunk_off[c].
struct BlockDesc {      // 32 bytes, read by the kernel with two 16-byte loads
  int row0, nrows;      // new-order row range
  int chunk0, nchunks;  // into chunk_off (L chunks followed by U chunks)
  int levelsL, levelsU; // informational
  int pad0, pad1;
};
template <class T> struct FusedBlockStream {
  int n = 0, nblocks = 0, max_rows = 0, max_levels_L = 0, max_levels_U = 0;
  int max_chunk_bytes = 0, max_chunks = 0;
  int64_t total_steps = 0, total_slots = 0;
  // LATTICE layout (chosen when every block is a tx x ty grid tile, tx <= 32, in row-major local order, whose
  // factors only couple west/south (L) and east/north (U) neighbours -- ILU(0) of a 5-point stencil):
  // no index is stored at all.  blk[b].pad0 = tx, pad1 = ty.  A chunk holds whole anti-diagonals:
  //     [ u16 nlev, u16 first_step, u16 lower(1)/upper(0), pad ]            16 B header
  //     [ T x2 pair ] per row of each anti-diagonal, ascending x             L: (west, south)  U: (east, north)
  // Lower steps walk the anti-diagonals d = x + y upwards, upper steps downwards.  Lane x owns column x of the
  // tile: its south/north operand is its own previous result, its west/east operand arrives by warp shuffle.
  bool lattice = false;
  std::vector<int> perm;             // n: new -> old
  std::vector<T> dinv;               // n, block-major: 1 / diag(U)
  std::vector<BlockDesc> blk;        // nblocks
  std::vector<uint32_t> chunk_off;   // total chunks + 1, in units of 16 bytes
  std::vector<unsigned char> stream;
};
// L: unit lower, diagonal LAST in each row (value ignored); U: diagonal FIRST.  Every off-diagonal
// of a row must lie in the row's own block (block-Jacobi factors).
template <class T>
void pack_fused_block_stream(bool unit_lower, const Csr<T>& L, const Csr<T>& U, int nblocks, const int* block_of_row,
                             int chunk_target, FusedBlockStream<T>& out, bool allow_lattice = true);

// ------------------------------------------------------------------ multi-GPU row partition
// One rank's share of a row-partitioned system: rank p owns the rows with part[i] == p
// (kept in ascending global order), every column it references that another rank owns becomes a
// HALO slot behind its own n_local entries.  Built independently on every rank from the global
// pattern, so both sides of an exchange agree without any communication:
//   local CSR column c < n_local  -> own row l2g[c];  c >= n_local -> halo slot c - n_local = global h2g[...]
//   halo slots are grouped by owner rank (ascending), ascending global id inside a group
//   neighbour k = rank nbr[k]: receives halo slots [recv_ptr[k], recv_ptr[k+1]) and is sent the own
//   rows send_idx[send_ptr[k] .. send_ptr[k+1]) (local ids, ascending global id = the order in which
//   that neighbour stores them in ITS halo)
template <class T> struct DistPlan {
  int nranks = 1, rank = 0;
  int64_t n_global = 0;
  int n_local = 0, n_halo = 0;
  std::vector<int> l2g, h2g;
  Csr<T> A;                       // n_local rows, columns in [0, n_local + n_halo)
  std::vector<int> nbr, recv_ptr, send_ptr, send_idx;
};
template <class T>
void build_dist_plan(int n, const int* rp, const int* ci, const T* v, int nranks, int rank, const int* part,
                     DistPlan<T>& out);

// Row-length histogram -> SpMV format (role of SpMV_inspect.cpp:38-456)
struct SpmvPlan {
  int format = 0, max_row = 0, slice = 32;
  double mean_row = 0;
  int64_t hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t sell_padded = 0;
};
SpmvPlan inspect_rows(int n, const int* rp);
// sliced ELL (slices padded to max(their own longest row, min_width)); slice_ptr in entries.
// slice_ptr of the sliced-ELL layout alone (the fill may then happen on the device: spmv.cu sell_fill_kernel)
void sell_slice_ptr(int n, const int* rp, int slice, int min_width, std::vector<int64_t>& slice_ptr);
template <class T>
void build_sell(int n, const int* rp, const int* ci, const T* v, int slice, std::vector<int64_t>& slice_ptr,
                std::vector<int>& cols, std::vector<T>& vals, int min_width = 0);

}  // namespace bg
