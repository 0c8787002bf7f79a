// pack_check.cpp -- host test of the fused block-stream packer (host_algos.h FusedBlockStream).
// Interprets the packed chunks exactly like trsv_block_fused does (step by step, 32 "lanes",
// 16-bit local ids, unit solves on the rescaled factors) and compares with a plain CSR forward/backward solve.
// Test infrastructure only: built and run by tests/test_host_logic.py, never linked into the library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../gpu-gmres_b200/host/host_algos.h"

using namespace bg;

template <class T>
static int run_case(int n, int nblocks, int max_off, int chunk_target, unsigned seed, bool contiguous_blocks) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  std::vector<int> blk(n);
  for (int i = 0; i < n; ++i) blk[i] = contiguous_blocks ? std::min(nblocks - 1, (int)((int64_t)i * nblocks / n)) : (int)(rng() % nblocks);
  // random block-diagonal strictly-lower / strictly-upper patterns
  Csr<T> L, U;
  L.n = U.n = n;
  L.rp.assign(1, 0);
  U.rp.assign(1, 0);
  std::vector<std::vector<int>> members(nblocks);
  for (int i = 0; i < n; ++i) members[blk[i]].push_back(i);
  std::vector<int> posin(n);
  for (auto& m : members)
    for (size_t k = 0; k < m.size(); ++k) posin[m[k]] = (int)k;
  for (int i = 0; i < n; ++i) {
    const auto& m = members[blk[i]];
    const int k = posin[i];
    std::vector<int> lo, up;
    const int nl = (int)(rng() % (max_off + 1)), nu = (int)(rng() % (max_off + 1));
    for (int q = 0; q < nl && k > 0; ++q) lo.push_back(m[rng() % k]);
    for (int q = 0; q < nu && k + 1 < (int)m.size(); ++q) up.push_back(m[k + 1 + rng() % (m.size() - k - 1)]);
    std::sort(lo.begin(), lo.end()); lo.erase(std::unique(lo.begin(), lo.end()), lo.end());
    std::sort(up.begin(), up.end()); up.erase(std::unique(up.begin(), up.end()), up.end());
    for (int c : lo) { L.ci.push_back(c); L.v.push_back((T)(0.3 * u(rng))); }
    L.ci.push_back(i); L.v.push_back((T)1);
    L.rp.push_back((int)L.ci.size());
    U.ci.push_back(i); U.v.push_back((T)(2.0 + u(rng)));
    for (int c : up) { U.ci.push_back(c); U.v.push_back((T)(0.3 * u(rng))); }
    U.rp.push_back((int)U.ci.size());
  }
  FusedBlockStream<T> S;
  pack_fused_block_stream<T>(true, L, U, nblocks, blk.data(), chunk_target, S, /*allow_lattice=*/false);
  // reference: y = L^-1 r, x = U^-1 y
  std::vector<double> r(n), x(n);
  for (int i = 0; i < n; ++i) r[i] = u(rng);
  std::vector<double> y(r);
  for (int i = 0; i < n; ++i) {
    double a = y[i];
    for (int p = L.rp[i]; p < L.rp[i + 1] - 1; ++p) a -= (double)L.v[p] * y[L.ci[p]];
    y[i] = a;
  }
  x = y;
  for (int i = n - 1; i >= 0; --i) {
    double a = x[i];
    for (int p = U.rp[i] + 1; p < U.rp[i + 1]; ++p) a -= (double)U.v[p] * x[U.ci[p]];
    x[i] = a / (double)U.v[U.rp[i]];
  }
  // interpreter
  std::vector<double> got(n, 0.0), xl(S.max_rows);
  int64_t slots = 0;
  std::vector<char> seen(n, 0);
  for (int i = 0; i < n; ++i) {
    if (S.perm[i] < 0 || S.perm[i] >= n || seen[S.perm[i]]) { std::printf("perm is not a permutation\n"); return 1; }
    seen[S.perm[i]] = 1;
  }
  for (int b = 0; b < S.nblocks; ++b) {
    const BlockDesc& d = S.blk[b];
    for (int i = 0; i < d.nrows; ++i) xl[i] = r[S.perm[d.row0 + i]] * (double)S.dinv[d.row0 + i];
    for (int c = 0; c < d.nchunks; ++c) {
      const size_t off = (size_t)S.chunk_off[d.chunk0 + c] * 16, end = (size_t)S.chunk_off[d.chunk0 + c + 1] * 16;
      if ((int)(end - off) > S.max_chunk_bytes || end > S.stream.size()) { std::printf("chunk bounds\n"); return 1; }
      const unsigned char* rec = S.stream.data() + off;
      const uint16_t* hdr = reinterpret_cast<const uint16_t*>(rec);
      const int ns = hdr[0], K = hdr[1], nst = hdr[2], nsA = hdr[3];
      const int stA = (nst + 16) / 16 * 16, KQ = (K + 4) / 4;
      const unsigned char* cnts = rec + 16;
      const uint16_t* quads = reinterpret_cast<const uint16_t*>(rec + 16 + stA);
      const T* vals = reinterpret_cast<const T*>(rec + 16 + stA + 8 * (size_t)KQ * nsA);
      if ((const unsigned char*)(vals + (size_t)K * nsA) > S.stream.data() + end) { std::printf("record overruns its chunk\n"); return 1; }
      int s0 = 0;
      for (int s = 0; s < nst; ++s) {
        const int cnt = cnts[s];
        if (cnt < 1 || cnt > 32) { std::printf("bad step count\n"); return 1; }
        std::vector<double> res(cnt);
        for (int lane = 0; lane < cnt; ++lane) {  // all lanes read before any lane writes
          const int t = s0 + lane, li = quads[(size_t)t * 4];
          double a = xl[li];
          for (int k = 0; k < K; ++k) {
            const int e = k + 1;
            const int col = quads[((size_t)(e / 4) * nsA + t) * 4 + e % 4];
            const double v = (k == K - 1 && (K & 1)) ? (double)vals[(size_t)(K / 2) * 2 * nsA + t]
                                                     : (double)vals[((size_t)(k / 2) * nsA + t) * 2 + k % 2];
            a -= v * xl[col];
          }
          res[lane] = a;
        }
        for (int lane = 0; lane < cnt; ++lane) xl[quads[(size_t)(s0 + lane) * 4]] = res[lane];
        s0 += cnt;
      }
      if (s0 != ns) { std::printf("steps do not cover the chunk\n"); return 1; }
      slots += ns;
    }
    for (int i = 0; i < d.nrows; ++i) got[S.perm[d.row0 + i]] = xl[i];
  }
  if (slots != 2 * (int64_t)n) { std::printf("slots %lld != 2n\n", (long long)slots); return 1; }
  double err = 0, nrm = 0;
  for (int i = 0; i < n; ++i) { err += (got[i] - x[i]) * (got[i] - x[i]); nrm += x[i] * x[i]; }
  err = std::sqrt(err / (nrm > 0 ? nrm : 1));
  const double tol = sizeof(T) == 4 ? 1e-5 : 1e-12;
  std::printf("n=%d blocks=%d K<=%d target=%d chunks=%zu max_chunk=%d err=%.2e\n", n, nblocks, max_off, chunk_target,
              S.chunk_off.size() - 1, S.max_chunk_bytes, err);
  return err < tol ? 0 : 1;
}

// LATTICE layout: ILU(0)-like factors of a 5-point grid cut into tiles; the interpreter below mirrors
// run_lattice_chunk of trsv_block.cu (lane x owns column x of the tile, own previous result + shuffled neighbour).
template <class T>
static int run_lattice_case(int nx, int ny, int tile_x, int tile_y, int chunk_target, unsigned seed, bool drop_some) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  const int n = nx * ny, tbx = (nx + tile_x - 1) / tile_x, tby = (ny + tile_y - 1) / tile_y, nblocks = tbx * tby;
  std::vector<int> blk(n);
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) blk[y * nx + x] = (y / tile_y) * tbx + x / tile_x;
  Csr<T> L, U;
  L.n = U.n = n;
  L.rp.assign(1, 0);
  U.rp.assign(1, 0);
  auto keep = [&]() { return !drop_some || (rng() % 5) != 0; };
  for (int i = 0; i < n; ++i) {
    const int x = i % nx, y = i / nx;
    if (y > 0 && blk[i - nx] == blk[i] && keep()) { L.ci.push_back(i - nx); L.v.push_back((T)(0.3 * u(rng))); }
    if (x > 0 && blk[i - 1] == blk[i] && keep()) { L.ci.push_back(i - 1); L.v.push_back((T)(0.3 * u(rng))); }
    L.ci.push_back(i); L.v.push_back((T)1);
    L.rp.push_back((int)L.ci.size());
    U.ci.push_back(i); U.v.push_back((T)(2.0 + u(rng)));
    if (x + 1 < nx && blk[i + 1] == blk[i] && keep()) { U.ci.push_back(i + 1); U.v.push_back((T)(0.3 * u(rng))); }
    if (y + 1 < ny && blk[i + nx] == blk[i] && keep()) { U.ci.push_back(i + nx); U.v.push_back((T)(0.3 * u(rng))); }
    U.rp.push_back((int)U.ci.size());
  }
  FusedBlockStream<T> S;
  pack_fused_block_stream<T>(true, L, U, nblocks, blk.data(), chunk_target, S);
  if (!S.lattice) { std::printf("lattice layout was not chosen for a %dx%d grid with %dx%d tiles\n", nx, ny, tile_x, tile_y); return 1; }
  std::vector<double> r(n), x(n);
  for (int i = 0; i < n; ++i) r[i] = u(rng);
  std::vector<double> yv(r);
  for (int i = 0; i < n; ++i) {
    double a = yv[i];
    for (int p = L.rp[i]; p < L.rp[i + 1] - 1; ++p) a -= (double)L.v[p] * yv[L.ci[p]];
    yv[i] = a;
  }
  x = yv;
  for (int i = n - 1; i >= 0; --i) {
    double a = x[i];
    for (int p = U.rp[i] + 1; p < U.rp[i + 1]; ++p) a -= (double)U.v[p] * x[U.ci[p]];
    x[i] = a / (double)U.v[U.rp[i]];
  }
  std::vector<double> got(n, 0.0), xl(S.max_rows);
  for (int b = 0; b < S.nblocks; ++b) {
    const BlockDesc& d = S.blk[b];
    const int tx = d.pad0, ty = d.pad1;
    if (tx * ty != d.nrows || tx > 32) { std::printf("bad tile dims\n"); return 1; }
    for (int i = 0; i < d.nrows; ++i) xl[i] = r[S.perm[d.row0 + i]] * (double)S.dinv[d.row0 + i];
    double zprev[32];
    int expect_step = 0, expect_lower = 1;
    for (int c = 0; c < d.nchunks; ++c) {
      const size_t off = (size_t)S.chunk_off[d.chunk0 + c] * 16, end = (size_t)S.chunk_off[d.chunk0 + c + 1] * 16;
      if ((int)(end - off) > S.max_chunk_bytes || end > S.stream.size()) { std::printf("chunk bounds\n"); return 1; }
      const unsigned char* rec = S.stream.data() + off;
      const uint16_t* hdr = reinterpret_cast<const uint16_t*>(rec);
      const int nlev = hdr[0], e0 = hdr[1], lower = hdr[2];
      if (e0 == 0) { for (double& z : zprev) z = 0.0; if (c > 0) expect_lower = 0; expect_step = 0; }
      if (lower != expect_lower || e0 != expect_step) { std::printf("chunk order\n"); return 1; }
      const T* pr = reinterpret_cast<const T*>(rec + 16);
      const int D = tx + ty - 1;
      int t0 = 0;
      for (int s = 0; s < nlev; ++s) {
        const int e = e0 + s, dd = lower ? e : D - 1 - e;
        const int xlo = std::max(0, dd - (ty - 1)), xhi = std::min(dd, tx - 1);
        double zn[32];
        for (int lane = 0; lane < 32; ++lane) zn[lane] = lower ? zprev[lane > 0 ? lane - 1 : 0] : zprev[lane < 31 ? lane + 1 : 31];  // shuffle
        for (int lane = xlo; lane <= xhi; ++lane) {
          const int li = lane + tx * (dd - lane);
          const double a = (double)pr[2 * (t0 + lane - xlo)], bq = (double)pr[2 * (t0 + lane - xlo) + 1];
          const double z = (xl[li] - bq * zprev[lane]) - a * zn[lane];
          xl[li] = z;
          zprev[lane] = z;
        }
        t0 += xhi - xlo + 1;
      }
      if ((const unsigned char*)(pr + 2 * (size_t)t0) > S.stream.data() + end) { std::printf("record overruns its chunk\n"); return 1; }
      expect_step = e0 + nlev;
    }
    for (int i = 0; i < d.nrows; ++i) got[S.perm[d.row0 + i]] = xl[i];
  }
  double err = 0, nrm = 0;
  for (int i = 0; i < n; ++i) { err += (got[i] - x[i]) * (got[i] - x[i]); nrm += x[i] * x[i]; }
  err = std::sqrt(err / (nrm > 0 ? nrm : 1));
  const double tol = sizeof(T) == 4 ? 1e-5 : 1e-12;
  std::printf("lattice %dx%d tiles %dx%d target=%d chunks=%zu max_chunk=%d err=%.2e\n", nx, ny, tile_x, tile_y, chunk_target,
              S.chunk_off.size() - 1, S.max_chunk_bytes, err);
  return err < tol ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run_lattice_case<double>(70, 45, 32, 32, 4096, 11, false);  // ragged edge tiles (6 wide, 13 high)
  bad += run_lattice_case<double>(64, 64, 16, 16, 512, 12, true);    // missing couplings stay zero coefficients
  bad += run_lattice_case<float>(33, 9, 32, 4, 2048, 13, false);     // 1-wide edge tile
  bad += run_lattice_case<double>(20, 1, 20, 1, 4096, 14, false);    // a single lattice row
  bad += run_case<double>(1000, 7, 3, 2048, 1, false);
  bad += run_case<double>(5000, 5, 12, 1024, 2, true);
  bad += run_case<float>(3000, 40, 2, 512, 3, false);
  bad += run_case<double>(64, 64, 2, 2048, 4, true);   // one row per block
  bad += run_case<double>(300, 1, 40, 4096, 5, true);  // one block, wide rows: a step alone exceeds the target
  bad += run_case<float>(1, 1, 0, 2048, 6, true);
  return bad ? 1 : 0;
}
