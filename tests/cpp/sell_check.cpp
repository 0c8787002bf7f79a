// sell_check.cpp -- host test of the sliced-ELL packer (host_algos.h build_sell), at a size small enough for one
// thread and at one large enough for the threaded fill: every CSR entry sits at [slice_ptr[s] + k*slice + t] in row
// order, padding carries value 0 and an in-range column, and a SELL product equals the CSR product bit for bit.
// Test infrastructure only: built and run by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../gpu-gmres_b200/host/host_algos.h"

using namespace bg;

static int check(int n, int slice, int min_width, unsigned seed) {
  std::mt19937 g(seed);
  std::vector<int> rp(n + 1, 0), ci;
  std::vector<double> v;
  for (int i = 0; i < n; ++i) {
    const int len = (int)(g() % 7);  // 0..6 entries: empty rows and ragged slices included
    for (int k = 0; k < len; ++k) { ci.push_back((int)(g() % n)); v.push_back((double)(g() % 1000) / 37.0 - 13.0); }
    rp[i + 1] = (int)ci.size();
  }
  std::vector<int64_t> sp;
  std::vector<int> sc;
  std::vector<double> sv;
  build_sell(n, rp.data(), ci.data(), v.data(), slice, sp, sc, sv, min_width);
  const int ns = (n + slice - 1) / slice;
  if ((int)sp.size() != ns + 1 || sp[0] != 0 || (int64_t)sc.size() != sp[ns] || sv.size() != sc.size()) { std::printf("shape\n"); return 1; }
  std::vector<double> x(n), y0(n, 0.0), y1(n, 0.0);
  for (int i = 0; i < n; ++i) x[i] = (double)(g() % 100) / 7.0;
  for (int i = 0; i < n; ++i) for (int p = rp[i]; p < rp[i + 1]; ++p) y0[i] += v[p] * x[ci[p]];
  for (int s = 0; s < ns; ++s) {
    const int64_t wid = (sp[s + 1] - sp[s]) / slice;
    if ((sp[s + 1] - sp[s]) % slice || wid < min_width) { std::printf("slice %d width\n", s); return 1; }
    int mx = min_width;
    for (int t = 0; t < slice; ++t) {
      const int i = s * slice + t;
      const int len = i < n ? rp[i + 1] - rp[i] : 0;
      mx = std::max(mx, len);
      for (int64_t k = 0; k < wid; ++k) {
        const int64_t q = sp[s] + k * slice + t;
        if (k < len) {
          if (sc[q] != ci[rp[i] + k] || sv[q] != v[rp[i] + k]) { std::printf("entry row %d k %lld\n", i, (long long)k); return 1; }
        } else if (sv[q] != 0.0 || sc[q] < 0 || sc[q] >= n) { std::printf("padding row %d k %lld\n", i, (long long)k); return 1; }
        if (i < n) y1[i] += sv[q] * x[sc[q]];
      }
    }
    if (wid != mx) { std::printf("slice %d wider than its longest row\n", s); return 1; }
  }
  for (int i = 0; i < n; ++i) if (y0[i] != y1[i]) { std::printf("product row %d\n", i); return 1; }
  return 0;
}

int main() {
  int bad = 0;
  bad += check(1000, 64, 0, 1);
  bad += check(1037, 64, 6, 2);   // ragged last slice, uniform padding
  bad += check(5, 128, 0, 3);     // fewer rows than one slice
  bad += check((1 << 21) + 77, 64, 0, 4);   // threaded fill
  bad += check((1 << 21) + 77, 128, 6, 5);
  if (bad) { std::printf("FAILED %d\n", bad); return 1; }
  std::printf("ok\n");
  return 0;
}
