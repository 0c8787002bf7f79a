// host_algos.h -- host-side setup for the B200 GMRES(m)+ILU path: ILU(0)/ILUT factorisation,
// level sets, row/block partitioning, and the packers that turn CSR factors into the streams the
// triangular-solve kernels consume.  Pure C++ (no CUDA): linked into libb200gmres.so and usable
// without a GPU (bg_host_* entry points).
#pragma once
#include <cstdint>
#include <vector>

namespace bg {

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
// ILUT(tau, lfil): dual-threshold incomplete LU (Saad), same output layout, L unit lower.
template <class T>
void ilut(int n, const int* rp, const int* ci, const T* v, double tau, int lfil, const int* block,
          Factors<T>& F);

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
// Level stream (BG_TRSV_LEVEL): rows sorted by (level, row id); per level an ELL slab.
//   slot s in [lev_ptr[l], lev_ptr[l+1]) holds row rows[s];
//   off-diagonal entry k of that slot:  cols/vals[lev_off[l] + k*R_l + (s - lev_ptr[l])], col -1 = padding
//   diag[s] (1 for unit-lower)
template <class T> struct LevelStream {
  int n = 0, n_levels = 0, max_level_rows = 0;
  std::vector<int> level;            // per row (for parity checks)
  std::vector<int> lev_ptr, lev_K, rows;
  std::vector<int64_t> lev_off;
  std::vector<int> cols;
  std::vector<T> vals, diag;
  // launch plan: consecutive levels [seg_begin[i], seg_end[i]) ; seg_single[i]=1 -> one CTA walks
  // them with __syncthreads, 0 -> exactly one fat level launched over many CTAs
  std::vector<int> seg_begin, seg_end, seg_single;
};
// lower: diagonal is the LAST entry of each row (unit => value ignored); upper: FIRST entry.
template <class T>
void pack_level_stream(bool lower, bool unit, const Csr<T>& M, int thin_rows, LevelStream<T>& out);

// Block stream (BG_TRSV_BLOCK): every block is solved by one warp out of shared memory.
// Block b owns rows grow[blk_row_ptr[b] .. blk_row_ptr[b+1]) (ascending); local id = position.
// Its schedule lists local ids level by level; the packed stream is cut into chunks of CS slots,
// each chunk one contiguous, 16-byte aligned record
//     [ sched : CS x u16 ][ cols : K x CS x u16 ][ vals : K x CS x T ][ diag : CS x T ]
// so one bulk-async copy stages a chunk.  Padding slots have sched = 0xFFFF; padding entries have
// val = 0 and col = the slot's own local id.
template <class T> struct BlockStream {
  int n = 0, nblocks = 0, K = 0, CS = 0, max_rows = 0, max_levels = 0, max_chunks = 0;
  int64_t chunk_bytes = 0;
  std::vector<int> grow;            // n, rows grouped by block, ascending inside a block
  std::vector<int> blk_row_ptr;     // nblocks+1
  std::vector<int> blk_lev_ptr;     // nblocks+1 -> into lev_slot
  std::vector<int> lev_slot;        // per block: n_levels+1 slot boundaries (0-based inside block)
  std::vector<int64_t> blk_chunk_ptr;  // nblocks+1 -> chunk index of the block's first chunk
  std::vector<unsigned char> stream;   // all chunks
};
template <class T>
void pack_block_stream(bool lower, bool unit, const Csr<T>& M, int nblocks, const int* block_of_row,
                       int chunk_slots, BlockStream<T>& out);

// Row-length histogram -> SpMV format (role of SpMV_inspect.cpp:38-456)
struct SpmvPlan {
  int format = 0, max_row = 0, slice = 32;
  double mean_row = 0;
  int64_t hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t sell_padded = 0;
};
SpmvPlan inspect_rows(int n, const int* rp);
// sliced ELL (slice height 32, slices padded to their own max row length); slice_ptr in entries.
template <class T>
void build_sell(int n, const int* rp, const int* ci, const T* v, int slice, std::vector<int64_t>& slice_ptr,
                std::vector<int>& cols, std::vector<T>& vals);

}  // namespace bg
