// stream_check.cpp -- host test of the two BG_TRSV_LEVEL stream packers (host_algos.h SyncFreeStream,
// LatticeGlobal).  Interprets the packed streams exactly like trsv_syncfree / trsv_lattice_global do
// (slice by slice with a "has this row been published" check standing in for the flag-in-data poll; strip by
// strip, 32 lanes, step by step with the boundary column handed from strip to strip) and compares with a plain
// CSR forward / backward solve.  Test infrastructure only: built and run by tests/test_host_logic.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../gpu-gmres_b200/host/host_algos.h"

using namespace bg;

template <class T> static std::vector<double> csr_lower(const Csr<T>& L, bool unit, const std::vector<double>& r) {
  std::vector<double> y(r);
  for (int i = 0; i < L.n; ++i) {
    double a = y[i];
    for (int p = L.rp[i]; p < L.rp[i + 1] - 1; ++p) a -= (double)L.v[p] * y[L.ci[p]];
    y[i] = unit ? a : a / (double)L.v[L.rp[i + 1] - 1];
  }
  return y;
}
template <class T> static std::vector<double> csr_upper(const Csr<T>& U, const std::vector<double>& r) {
  std::vector<double> x(r);
  for (int i = U.n - 1; i >= 0; --i) {
    double a = x[i];
    for (int p = U.rp[i] + 1; p < U.rp[i + 1]; ++p) a -= (double)U.v[p] * x[U.ci[p]];
    x[i] = a / (double)U.v[U.rp[i]];
  }
  return x;
}

// ---- sync-free stream, interpreted like trsv_syncfree
template <class T>
static int run_syncfree(const Csr<T>& M, bool lower, bool unit, const std::vector<double>& r, const std::vector<double>& want,
                        const char* what) {
  SyncFreeStream<T> S;
  pack_syncfree_stream<T>(lower, unit, M, S);
  const int n = M.n;
  std::vector<double> x(n, 0.0);
  std::vector<char> ready(n, 0);
  int prev_level = -1;
  std::vector<char> seen(n, 0);
  if ((int)S.meta.size() != 4 * (S.n_slices + 1)) { std::printf("%s: metadata size\n", what); return 1; }
  std::vector<int> per_level(S.n_levels, 0);
  for (int q = 0; q < S.n_slices; ++q) {
    const int s0 = S.meta[4 * q], R = S.meta[4 * q + 1] & 255, K = S.meta[4 * q + 1] >> 8;
    if (R < 1 || R > 32) { std::printf("%s: slice %d has %d rows\n", what, q, R); return 1; }
    const int64_t off = (int64_t)S.meta[4 * q + 3] * 32;
    const int lev = S.sched_level[S.rows[s0]];
    if (lev != S.meta[4 * q + 2]) { std::printf("%s: slice %d carries level %d, its rows are level %d\n", what, q, S.meta[4 * q + 2], lev); return 1; }
    per_level[lev]++;
    if (lev < prev_level) { std::printf("%s: slices not in level order\n", what); return 1; }
    prev_level = lev;
    std::vector<double> res(R);
    for (int t = 0; t < R; ++t) {  // the 32 lanes of the warp: nobody may read a row of the same slice
      const int row = S.rows[s0 + t];
      if (S.sched_level[row] != lev) { std::printf("%s: slice %d spans two levels\n", what, q); return 1; }
      if (seen[row]++) { std::printf("%s: row %d scheduled twice\n", what, row); return 1; }
      double acc = r[row];
      int plev = -1;
      bool started = false;
      for (int k = 0; k < K; ++k) {
        const int c = S.cols[off + (int64_t)k * 32 + t];
        if (c < 0) {
          if (started) { std::printf("%s: row %d is not right-aligned in its slice\n", what, row); return 1; }
          continue;
        }
        started = true;
        if (!ready[c]) { std::printf("%s: row %d reads row %d before it is published (deadlock on the device)\n", what, row, c); return 1; }
        if (S.sched_level[c] < plev) { std::printf("%s: row %d: operands not in level order\n", what, row); return 1; }
        plev = S.sched_level[c];
        acc -= (double)S.vals[off + (int64_t)k * 32 + t] * x[c];
      }
      if (K > 0 && M.rp[row + 1] - M.rp[row] - 1 > 0 && S.cols[off + (int64_t)(K - 1) * 32 + t] < 0) { std::printf("%s: row %d: last entry is padding\n", what, row); return 1; }
      res[t] = unit ? acc : acc / (double)S.diag[s0 + t];
    }
    for (int t = 0; t < R; ++t) { x[S.rows[s0 + t]] = res[t]; ready[S.rows[s0 + t]] = 1; }
  }
  for (int l = 0; l < S.n_levels; ++l)
    if (per_level[l] != S.level_slices[l]) { std::printf("%s: level %d has %d slices, level_slices says %d\n", what, l, per_level[l], S.level_slices[l]); return 1; }
  double err = 0, nrm = 0;
  for (int i = 0; i < n; ++i) {
    if (!seen[i]) { std::printf("%s: row %d never scheduled\n", what, i); return 1; }
    err = std::max(err, std::fabs(x[i] - want[i]));
    nrm = std::max(nrm, std::fabs(want[i]));
  }
  if (err > 1e-12 * std::max(1.0, nrm) * (sizeof(T) == 4 ? 1e6 : 1.0)) { std::printf("%s: error %.3e\n", what, err); return 1; }
  return 0;
}

template <class T> static int syncfree_case(int n, int max_off, unsigned seed, double tiny_prob) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  Csr<T> L, U;
  L.n = U.n = n;
  L.rp.assign(1, 0);
  U.rp.assign(1, 0);
  for (int i = 0; i < n; ++i) {
    std::vector<int> lo, up;
    const int nl = (int)(rng() % (max_off + 1)), nu = (int)(rng() % (max_off + 1));
    for (int q = 0; q < nl && i > 0; ++q) lo.push_back((int)(rng() % i));
    for (int q = 0; q < nu && i + 1 < n; ++q) up.push_back(i + 1 + (int)(rng() % (n - i - 1)));
    std::sort(lo.begin(), lo.end()); lo.erase(std::unique(lo.begin(), lo.end()), lo.end());
    std::sort(up.begin(), up.end()); up.erase(std::unique(up.begin(), up.end()), up.end());
    // explicit tiny entries (|v| < 1e-9): getLevel stops at / skips them, the device schedule must not
    for (int c : lo) { L.ci.push_back(c); L.v.push_back((T)((u(rng) < 2 * tiny_prob - 1) ? 1e-12 : 0.3 * u(rng))); }
    L.ci.push_back(i); L.v.push_back((T)(1.5 + 0.5 * u(rng)));
    L.rp.push_back((int)L.ci.size());
    U.ci.push_back(i); U.v.push_back((T)(2.0 + u(rng)));
    for (int c : up) { U.ci.push_back(c); U.v.push_back((T)((u(rng) < 2 * tiny_prob - 1) ? 1e-12 : 0.3 * u(rng))); }
    U.rp.push_back((int)U.ci.size());
  }
  std::vector<double> r(n);
  for (auto& e : r) e = u(rng);
  int bad = 0;
  bad += run_syncfree<T>(L, true, false, r, csr_lower(L, false, r), "syncfree L non-unit");
  bad += run_syncfree<T>(L, true, true, r, csr_lower(L, true, r), "syncfree L unit");
  bad += run_syncfree<T>(U, false, false, r, csr_upper(U, r), "syncfree U");
  return bad;
}

// ---- 5-point grid ILU(0)-like factors
template <class T> static void grid_factors(int nx, int ny, unsigned seed, double drop, Csr<T>& L, Csr<T>& U) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> u(0.05, 0.45);
  std::uniform_real_distribution<double> p01(0.0, 1.0);
  const int n = nx * ny;
  L.n = U.n = n;
  L.rp.assign(1, 0);
  U.rp.assign(1, 0);
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) {
      const int r = y * nx + x;
      if (y > 0 && p01(rng) >= drop) { L.ci.push_back(r - nx); L.v.push_back((T)-u(rng)); }
      if (x > 0 && p01(rng) >= drop) { L.ci.push_back(r - 1); L.v.push_back((T)-u(rng)); }
      L.ci.push_back(r); L.v.push_back((T)1);
      L.rp.push_back((int)L.ci.size());
      U.ci.push_back(r); U.v.push_back((T)(2.0 + u(rng)));
      if (x < nx - 1 && p01(rng) >= drop) { U.ci.push_back(r + 1); U.v.push_back((T)-u(rng)); }
      if (y < ny - 1 && p01(rng) >= drop) { U.ci.push_back(r + nx); U.v.push_back((T)-u(rng)); }
      U.rp.push_back((int)U.ci.size());
    }
}

// trsv_lattice_global, one factor: strips in dependency order, 32 lanes, blocks of 8 steps; lane x reaches time
// row t (= y for L, ny-1-y for U) at step t + lag
template <class T>
static void lattice_emulate(const LatticeGlobal<T>& S, bool lower, const std::vector<double>& in, std::vector<double>& out) {
  const int nx = S.nx, ny = S.ny, BS = 8;
  const std::vector<T>& pair = lower ? S.lower : S.upper;
  std::vector<double> bnd((size_t)S.strips * ny, 0.0);
  std::vector<char> bnd_ok((size_t)S.strips * ny, 0);
  const int nblk = (ny + 31 + BS - 1) / BS;
  for (int b = 0; b < S.strips; ++b) {
    const int s = lower ? b : S.strips - 1 - b;
    const bool has_up = lower ? s > 0 : s < S.strips - 1, has_down = lower ? s < S.strips - 1 : s > 0;
    const int up = lower ? s - 1 : s + 1, edge_in = lower ? 0 : 31, edge_out = lower ? 31 : 0;
    double zprev[32] = {0};
    for (int d = 0; d < nblk * BS; ++d) {
      double z[32];
      for (int lane = 0; lane < 32; ++lane) {
        const int X = 32 * s + lane, lag = lower ? lane : 31 - lane, t = d - lag, y = lower ? t : ny - 1 - t;
        const bool act = X < nx && t >= 0 && t < ny;
        double ca = 0, cb = 0, r = 0;
        if (act) {
          const size_t q = ((size_t)s * ny + y) * 32 + lane;
          ca = (double)pair[2 * q]; cb = (double)pair[2 * q + 1];
          r = in[(size_t)y * nx + X];
          if (!lower) r *= (double)S.udinv[q];
        }
        double zn = lower ? (lane > 0 ? zprev[lane - 1] : 0.0) : (lane < 31 ? zprev[lane + 1] : 0.0);
        if (lane == edge_in) {
          const int ye = lower ? d : ny - 1 - d;  // the edge lane has no lag
          zn = 0.0;
          if (has_up && d < ny) {
            if (!bnd_ok[(size_t)up * ny + ye]) { std::printf("lattice: strip %d step %d reads an unpublished boundary row %d\n", s, d, ye); std::exit(1); }
            zn = bnd[(size_t)up * ny + ye];
          }
        }
        const double tt = r - cb * zprev[lane];
        z[lane] = act ? tt - ca * zn : 0.0;
        if (act) {
          out[(size_t)y * nx + X] = z[lane];
          if (lane == edge_out && has_down) { bnd[(size_t)s * ny + y] = z[lane]; bnd_ok[(size_t)s * ny + y] = 1; }
        }
      }
      for (int lane = 0; lane < 32; ++lane) zprev[lane] = z[lane];
    }
  }
}

template <class T> static int lattice_case(int nx, int ny, unsigned seed, double drop) {
  Csr<T> L, U;
  grid_factors<T>(nx, ny, seed, drop, L, U);
  int dnx = 0;
  if (!detect_lattice2d<T>(L, U, &dnx) || dnx != nx) { std::printf("lattice %dx%d: not detected (nx=%d)\n", nx, ny, dnx); return 1; }
  LatticeGlobal<T> S;
  pack_lattice_global<T>(L, U, nx, S);
  const int n = nx * ny;
  std::mt19937 rng(seed + 17);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  std::vector<double> r(n), y(n, 0.0), x(n, 0.0);
  for (auto& e : r) e = u(rng);
  lattice_emulate<T>(S, true, r, y);
  lattice_emulate<T>(S, false, y, x);
  const std::vector<double> want = csr_upper(U, csr_lower(L, true, r));
  double err = 0, nrm = 0;
  for (int i = 0; i < n; ++i) { err = std::max(err, std::fabs(x[i] - want[i])); nrm = std::max(nrm, std::fabs(want[i])); }
  if (!(err <= (sizeof(T) == 4 ? 1e-5 : 1e-13) * std::max(1.0, nrm))) { std::printf("lattice %dx%d: error %.3e\n", nx, ny, err); return 1; }
  // the general stream must agree on the same factors as well
  int bad = run_syncfree<T>(L, true, true, r, csr_lower(L, true, r), "syncfree lattice L");
  bad += run_syncfree<T>(U, false, false, r, csr_upper(U, r), "syncfree lattice U");
  return bad;
}

int main() {
  int bad = 0;
  bad += syncfree_case<double>(1, 2, 1, 0.0);
  bad += syncfree_case<double>(33, 2, 2, 0.0);
  bad += syncfree_case<double>(700, 3, 3, 0.0);
  bad += syncfree_case<double>(700, 9, 4, 0.2);   // rows wider than the register path, explicit tiny entries
  bad += syncfree_case<float>(500, 5, 5, 0.1);
  bad += syncfree_case<double>(4000, 1, 6, 0.0);  // long thin chains
  bad += lattice_case<double>(5, 4, 11, 0.0);
  bad += lattice_case<double>(32, 32, 12, 0.0);
  bad += lattice_case<double>(33, 7, 13, 0.0);     // partial second strip
  bad += lattice_case<double>(100, 37, 14, 0.1);   // missing couplings stay lattice
  bad += lattice_case<double>(64, 3, 15, 0.0);
  bad += lattice_case<float>(70, 45, 16, 0.0);
  // non-lattice factors must be refused
  {
    Csr<double> L, U;
    grid_factors<double>(20, 10, 21, 0.0, L, U);
    // add a fill entry (r, r-nx+1) to one row: the pattern ILUT produces
    const int r = 5 * 20 + 7;
    Csr<double> L2 = L;
    L2.ci.insert(L2.ci.begin() + L2.rp[r] + 1, r - 19);
    L2.v.insert(L2.v.begin() + L2.rp[r] + 1, 0.01);
    for (int i = r + 1; i <= L2.n; ++i) L2.rp[i]++;
    int nx = 0;
    if (detect_lattice2d<double>(L2, U, &nx)) { std::printf("fill entry accepted as lattice\n"); ++bad; }
  }
  if (bad) { std::printf("FAILED (%d)\n", bad); return 1; }
  std::printf("ok\n");
  return 0;
}
