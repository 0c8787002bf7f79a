// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's restarted-GMRES(m) hot path
// (sheldonucr/GPU-GMRES).  Nothing here is product code: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (gpu-gmres_b200/) never links it.
//
// Every routine is templated on the vector value type R (float = what the
// reference runs, double = the 1e-8 parity target) and, where the
// reference mixes precisions (MyILUPPfloat: FP64 factors, FP32 vectors), on the
// factor value type F.  Arithmetic order follows the reference's host twins so
// the float instantiation reproduces the reference's roundings; build with
// -ffp-contract=off (see Makefile) so no FMA contraction sneaks in.
//
// Parity pinning: the reference ships no golden vectors for this path
// (SURVEY.md 8c).  This restatement is pinned against the reference's own CPU
// translation units compiled unchanged into oracle/_ref (computeSpMV,
// LUSolve_ignoreZero, coo2csr_in, partition) by tests/test_oracle_golden.py,
// and against committed fixtures in tests/golden/ generated from it.
// The host GMRES/ILU0 loops live in .cu files that do not compile under CUDA
// 12.9, so for those the pin is "restated, cross-checked against scipy and
// against textbook identities" -- see DESIGN.md.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

// src/defs.h:45-47  `const float eps = 1e-9; #define Equal(a,b) (abs((a)-(b)) < eps)`
static const float kEps = 1e-9f;
template <class T> inline bool is_zero(T v) { return std::fabs(v) < kEps; }

// ---------------------------------------------------------------- BLAS-1 twins
// src/gmres.cu:31-74 (vec_initial, sscal, sapxy, norm2, dot)
template <class R> R dot_(const R* x, const R* y, int n) {
  R t = 0;
  for (int i = 0; i < n; ++i) t += x[i] * y[i];
  return t;
}
template <class R> R nrm2_(const R* v, int n) {
  R t = 0;
  for (int i = 0; i < n; ++i) t += v[i] * v[i];
  return std::sqrt(t);
}
template <class R> void scal_(R* v, const R* x, R a, int n) {
  for (int i = 0; i < n; ++i) v[i] = a * x[i];
}
template <class R> void axpy_(R* y, const R* x, R a, int n) {
  for (int i = 0; i < n; ++i) y[i] = a * x[i] + y[i];
}

// src/SpMV_compute.cpp:19-37 computeSpMV (output first, input vector last)
template <class R, class MV>
void spmv_(R* out, const MV* val, const int* rp, const int* ci, const R* in, int n) {
  for (int i = 0; i < n; ++i) {
    R t = 0;
    for (int j = rp[i]; j < rp[i + 1]; ++j) t += val[j] * in[ci[j]];
    out[i] = t;
  }
}

// src/gmres.cu:77-88 sgemv: v = alpha*A*x + beta*y through a temporary
template <class R>
void gemv_(R* v, const R* val, const int* rp, const int* ci, R alpha, const R* x, R beta,
           const R* y, int n, std::vector<R>& tmp) {
  tmp.resize(n);
  spmv_(tmp.data(), val, rp, ci, x, n);
  for (int i = 0; i < n; ++i) v[i] = alpha * tmp[i] + beta * y[i];
}

// src/gmres.cu:192-216
template <class R> void apply_rot(R* dx, R* dy, R cs, R sn) {
  R t = cs * (*dx) + sn * (*dy);
  *dy = -sn * (*dx) + cs * (*dy);
  *dx = t;
}
template <class R> void gen_rot(R dx, R dy, R* cs, R* sn) {
  if (dy == 0.0) {
    *cs = 1.0;
    *sn = 0.0;
  } else if (std::fabs(dy) > std::fabs(dx)) {
    R t = dx / dy;
    *sn = (R)(1.0 / std::sqrt(1.0 + (double)(R)(t * t)));  // double intermediates, as the reference
    *cs = t * (*sn);
  } else {
    R t = dy / dx;
    *cs = (R)(1.0 / std::sqrt(1.0 + (double)(R)(t * t)));
    *sn = t * (*cs);
  }
}

// src/gmres.cu:92-116 Update: back-substitute H y = s, x += V y
template <class R>
void update_(R* x, int k, const R* H, int m, const R* s, const R* V, int n) {
  std::vector<R> y(s, s + k + 1);
  for (int i = k; i >= 0; --i) {
    y[i] /= H[i + i * (m + 1)];
    for (int j = i - 1; j >= 0; --j) y[j] -= H[j + i * (m + 1)] * y[i];
  }
  for (int j = 0; j <= k; ++j)
    for (int i = 0; i < n; ++i) x[i] += V[i + (size_t)j * n] * y[j];
}

// ---------------------------------------------------------------- preconditioners
enum { PC_NONE = 0, PC_DIAG = 1, PC_ILU0 = 2, PC_SPLIT = 3 };

template <class R, class F> struct Pc {
  int kind = PC_NONE, n = 0;
  // PC_DIAG: dinv[i]                          src/preconditioner.cu:480-484
  const R* dinv = nullptr;
  // PC_ILU0 / PC_SPLIT factors (CSR)
  const F *lv = nullptr, *uv = nullptr;
  const int *lp = nullptr, *li = nullptr, *up = nullptr, *ui = nullptr;
  // PC_SPLIT extras                           src/preconditioner.cu:1350-1416
  const R *middle = nullptr, *lscale = nullptr, *rscale = nullptr;
  const int *permRow = nullptr, *permCol = nullptr;
  std::vector<R> tmp;

  // MyILU0::HostPrecond, src/preconditioner.cu:205-251.  L unit-lower (stored diagonal
  // ignored: the division at :227 acts on the output buffer before it is overwritten at
  // :231), U upper with the diagonal located by scanning from the row end.
  void ilu0(const R* in, R* out) {
    tmp.assign(in, in + n);
    R* v = tmp.data();
    for (int i = 0; i < n; ++i)
      for (int j = lp[i]; j < lp[i + 1]; ++j) {
        if (li[j] >= i) break;
        v[i] -= lv[j] * v[li[j]];
      }
    std::memcpy(out, v, sizeof(R) * n);
    for (int i = n - 1; i >= 0; --i) {
      int lb = up[i], j = up[i + 1] - 1;
      for (; j >= lb; --j) {
        if (ui[j] <= i) break;
        out[i] -= uv[j] * out[ui[j]];
      }
      if (j >= lb && ui[j] == i && !is_zero(uv[j])) out[i] /= uv[j];
    }
  }
  // MyILUPPfloat::HostPrecond_left, src/preconditioner.cu:1374-1393
  void split_left(const R* in, R* out) {
    tmp.resize(n);
    R* v = tmp.data();
    for (int i = 0; i < n; ++i) v[i] = in[i] / lscale[i];
    for (int i = 0; i < n; ++i) out[i] = v[permRow[i]];
    for (int i = 0; i < n; ++i) {
      int lb = lp[i], ub = lp[i + 1];
      for (int j = lb; j < ub - 1; ++j) out[i] -= lv[j] * out[li[j]];
      out[i] /= lv[ub - 1];
    }
  }
  // MyILUPPfloat::HostPrecond_right, src/preconditioner.cu:1396-1416
  void split_right(const R* in, R* out) {
    tmp.resize(n);
    R* v = tmp.data();
    for (int i = 0; i < n; ++i) v[i] = in[i] * middle[i];
    for (int i = n - 1; i >= 0; --i) {
      int lb = up[i], ub = up[i + 1];
      for (int j = lb + 1; j < ub; ++j) v[i] -= uv[j] * v[ui[j]];
      v[i] /= uv[lb];
    }
    for (int i = 0; i < n; ++i) out[i] = v[permCol[i]] / rscale[i];
  }
  // MyILUPPfloat::HostPrecond_starting_value, src/preconditioner.cu:1354-1371
  void split_start(const R* in, R* out) {
    tmp.resize(n);
    R* v = tmp.data();
    for (int i = 0; i < n; ++i) v[i] = in[i] * rscale[i];
    for (int i = 0; i < n; ++i) out[permCol[i]] = v[i];
    for (int i = 0; i < n; ++i) {
      v[i] = 0;
      for (int j = up[i]; j < up[i + 1]; ++j) v[i] += uv[j] * out[ui[j]];
    }
    for (int i = 0; i < n; ++i) out[i] = v[i] / middle[i];
  }

  // Preconditioner::HostPrecond (left-only kinds)
  void full(const R* in, R* out) {
    switch (kind) {
      case PC_NONE: std::memcpy(out, in, sizeof(R) * n); break;  // preconditioner.h MyNONE
      case PC_DIAG:
        for (int i = 0; i < n; ++i) out[i] = in[i] * dinv[i];
        break;
      case PC_ILU0: ilu0(in, out); break;
      default: std::abort();
    }
  }
};

// ---------------------------------------------------------------- GMRES(m)
// One body for both host twins:
//   split == 0 : GMRES(..., Preconditioner&)  src/gmres.cu:1904-2066  (left preconditioning)
//   split == 1 : GMRESilu                     src/gmres.cu:2069-2252  (split preconditioning)
// max_iter / tol are in/out exactly as in the reference.
// hist (optional) receives the residual estimate of every inner iteration.
template <class R, class F>
int gmres_(const R* val, const int* rp, const int* ci, R* x, const R* b, int n, int m,
           int* max_iter, R* tol, Pc<R, F>& M, int split, R* hist, int hist_cap, int* hist_len) {
  R resid = 0;
  int i, j = 1, k, nh = 0;
  std::vector<R> s(m + 1), cs(m + 1), sn(m + 1), w(n), ww(n), r(n), rr(n), bb(n), y(n), tmp;
  std::vector<R> H((size_t)(m + 1) * m), V((size_t)(m + 1) * n);
  R* v = V.data();
  auto left = [&](const R* in, R* out) { split ? M.split_left(in, out) : M.full(in, out); };

  left(b, bb.data());
  R normb = nrm2_(bb.data(), n);
  if (normb == 0.0) normb = 1;
  if (split) M.split_start(x, y.data());
  gemv_(rr.data(), val, rp, ci, (R)-1.0, x, (R)1, b, n, tmp);
  left(rr.data(), r.data());
  R beta = nrm2_(r.data(), n);
  if (hist_len) *hist_len = 0;
  if ((resid = beta / normb) <= *tol) {
    *tol = resid;
    *max_iter = 0;
    return 0;
  }
  while (j <= *max_iter) {
    scal_(v, r.data(), (R)(1.0 / beta), n);
    std::fill(s.begin(), s.end(), (R)0);
    s[0] = beta;
    for (i = 0; i < m && j <= *max_iter; i++, j++) {
      if (split) {
        M.split_right(v + (size_t)i * n, w.data());
        spmv_(ww.data(), val, rp, ci, w.data(), n);
        M.split_left(ww.data(), w.data());
      } else {
        spmv_(ww.data(), val, rp, ci, v + (size_t)i * n, n);
        M.full(ww.data(), w.data());
      }
      for (k = 0; k <= i; k++) {
        H[k + i * (m + 1)] = dot_(w.data(), v + (size_t)k * n, n);
        axpy_(w.data(), v + (size_t)k * n, (R)(-H[k + i * (m + 1)]), n);
      }
      H[(i + 1) + i * (m + 1)] = nrm2_(w.data(), n);
      scal_(v + (size_t)(i + 1) * n, w.data(), (R)(1.0 / H[(i + 1) + i * (m + 1)]), n);
      for (k = 0; k < i; k++) apply_rot(&H[k + i * (m + 1)], &H[(k + 1) + i * (m + 1)], cs[k], sn[k]);
      gen_rot(H[i + i * (m + 1)], H[(i + 1) + i * (m + 1)], &cs[i], &sn[i]);
      apply_rot(&H[i + i * (m + 1)], &H[(i + 1) + i * (m + 1)], cs[i], sn[i]);
      apply_rot(&s[i], &s[i + 1], cs[i], sn[i]);
      resid = std::fabs(s[i + 1]) / normb;
      if (hist && nh < hist_cap) hist[nh++] = resid;
      if (hist_len) *hist_len = nh;
      if (resid < *tol) {
        if (split) {
          update_(y.data(), i, H.data(), m, s.data(), v, n);
          M.split_right(y.data(), x);
        } else {
          update_(x, i, H.data(), m, s.data(), v, n);
        }
        *tol = resid;
        *max_iter = j;
        return 0;
      }
    }
    if (split) {
      update_(y.data(), m - 1, H.data(), m, s.data(), v, n);
      M.split_right(y.data(), x);
    } else {
      update_(x, m - 1, H.data(), m, s.data(), v, n);
    }
    gemv_(rr.data(), val, rp, ci, (R)-1.0, x, (R)1, b, n, tmp);
    left(rr.data(), r.data());
    beta = nrm2_(r.data(), n);
    if ((resid = beta / normb) < *tol) {
      *tol = resid;
      *max_iter = j;
      return 0;
    }
  }
  *tol = resid;
  return 1;
}

// ---------------------------------------------------------------- ILU(0), leftILU CPU semantics
// generateLevel, src/leftILU.cu:339-368 (uses the INPUT values)
template <class R>
void gen_level_(int n, const R* val, const int* rp, const int* ci, int* level) {
  for (int i = 0; i < n; ++i) level[i] = 0;
  for (int r = 0; r < n; ++r)
    for (int p = rp[r]; p != rp[r + 1]; ++p)
      if (!is_zero(val[p]) && ci[p] > r) {
        int c = ci[p];
        if (level[c] < level[r] + 1) level[c] = level[r] + 1;
      }
}
// bucketing, src/leftILU.cu:41-49,188-192: ascending node id inside a level
static int bucket_levels(int n, const int* level, int* levelPtr, int* nodes) {
  int nl = 0;
  for (int i = 0; i < n; ++i) nl = std::max(nl, level[i] + 1);
  if (n == 0) return 0;
  std::vector<int> cnt(nl + 1, 0);
  for (int i = 0; i < n; ++i) cnt[level[i] + 1]++;
  for (int l = 0; l < nl; ++l) cnt[l + 1] += cnt[l];
  if (levelPtr) std::copy(cnt.begin(), cnt.end(), levelPtr);
  if (nodes) {
    std::vector<int> cur(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < n; ++i) nodes[cur[level[i]]++] = i;
  }
  return nl;
}

// csr2csc, src/leftILU.cu:371-417 (stable: rows ascending inside each column)
template <class R>
void transpose_(int n, const R* val, const int* rp, const int* ci, std::vector<R>& tv,
                std::vector<int>& tp, std::vector<int>& ti) {
  int nnz = rp[n];
  tp.assign(n + 1, 0);
  tv.resize(nnz);
  ti.resize(nnz);
  for (int p = 0; p < nnz; ++p) tp[ci[p] + 1]++;
  for (int c = 0; c < n; ++c) tp[c + 1] += tp[c];
  std::vector<int> cur(tp.begin(), tp.end() - 1);
  for (int r = 0; r < n; ++r)
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
      int q = cur[ci[p]]++;
      tv[q] = val[p];
      ti[q] = r;
    }
}

// first index in [lo,hi) with a[i] >= t, linear -- liSearchLowerBound, src/leftILU.cu:831-840
static inline int first_ge(int t, int lo, int hi, const int* a) {
  for (int i = lo; i < hi; ++i)
    if (a[i] >= t) return i;
  return hi;
}

// cpuSequentialTriSolve for one column, src/leftILU.cu:769-825
template <class R>
void factor_column_(int c, R* cv, const int* cp, const int* cr) {
  int lb = cp[c], ub = cp[c + 1];
  R udiag = 0;
  for (int p = lb; p < ub - 1; ++p) {
    int r = cr[p];
    if (r > c) break;
    if (r == c) {
      udiag = cv[p];
      break;
    }
    int llb = cp[r], lub = cp[r + 1];
    for (int q = p + 1; q < ub; ++q) {
      int rr = cr[q];
      int t = first_ge(rr, llb, lub, cr);
      if (t < lub && cr[t] == rr) cv[q] -= cv[t] * cv[p];
    }
  }
  for (int p = lb; p < ub; ++p) {
    if (cr[p] <= c) continue;
    if (!is_zero(udiag)) cv[p] /= udiag;
    else cv[p] = 0;
  }
}

// splitLU_csr, src/leftILU.cu:481-541 (two-pass: count, then fill)
template <class R>
void split_count_(int n, const R* v, const int* rp, const int* ci, int* lp, int* up) {
  lp[0] = up[0] = 0;
  for (int r = 0; r < n; ++r) {
    int nl = 0, nu = 0;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
      if (is_zero(v[p])) continue;
      if (ci[p] < r) nl++;
      else nu++;
    }
    lp[r + 1] = lp[r] + nl + 1;
    up[r + 1] = up[r] + nu;
  }
}
template <class R>
void split_fill_(int n, const R* v, const int* rp, const int* ci, const int* lp, R* lv, int* li,
                 const int* up, R* uv, int* ui) {
  for (int r = 0; r < n; ++r) {
    int a = lp[r], b = up[r];
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
      if (is_zero(v[p])) continue;
      if (ci[p] < r) { lv[a] = v[p]; li[a++] = ci[p]; }
      else { uv[b] = v[p]; ui[b++] = ci[p]; }
    }
    lv[a] = 1.0f;
    li[a] = r;
  }
}

// leftILU, src/leftILU.cu:27-336 with the CPU tail for every level.
template <class R> struct Ilu0Result {
  std::vector<R> fv;  // factored matrix, CSR, columns ascending in each row
  std::vector<int> fp, fi;
  std::vector<int> level, levelPtr, nodes;
};
template <class R>
void ilu0_(int n, const R* val, const int* rp, const int* ci, Ilu0Result<R>& out) {
  out.level.resize(n);
  gen_level_(n, val, rp, ci, out.level.data());
  int nl = bucket_levels(n, out.level.data(), nullptr, nullptr);
  out.levelPtr.resize(nl + 1);
  out.nodes.resize(n);
  bucket_levels(n, out.level.data(), out.levelPtr.data(), out.nodes.data());
  std::vector<R> cv;
  std::vector<int> cp, cr;
  transpose_(n, val, rp, ci, cv, cp, cr);
  for (int t = 0; t < n; ++t) factor_column_(out.nodes[t], cv.data(), cp.data(), cr.data());
  transpose_(n, cv.data(), cp.data(), cr.data(), out.fv, out.fp, out.fi);
}

// getLevel, src/gmres.cu:1033-1077
template <class R>
void get_level_(int isL, const R* val, const int* rp, const int* ci, int* level, int n) {
  for (int i = 0; i < n; ++i) level[i] = 0;
  if (isL) {
    for (int i = 1; i < n; ++i)
      for (int j = rp[i]; j < rp[i + 1]; ++j) {
        if (is_zero(val[j])) break;
        if (ci[j] >= i) break;
        level[i] = std::max(level[i], level[ci[j]] + 1);
      }
  } else {
    for (int i = n - 2; i >= 0; --i)
      for (int j = rp[i + 1] - 1; j >= rp[i]; --j) {
        if (is_zero(val[j])) continue;
        if (ci[j] <= i) break;
        level[i] = std::max(level[i], level[ci[j]] + 1);
      }
  }
}

// partition, src/partition.cpp:22-99 (CSC pattern, 64-bit indices)
void partition_(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, int64_t nNodes,
                int64_t* part_size, int64_t* node_part, int64_t* pinv, int64_t* q) {
  int64_t avg = nNodes / npart, j = 0;
  for (int64_t i = 0; i < m; ++i) node_part[i] = -1;
  for (int64_t i = 0; i < nNodes; ++i) {
    if (i >= j * avg && i < (j + 1) * avg) node_part[i] = j;
    else if (i >= (j + 1) * avg) {
      if (j + 1 < npart) node_part[i] = ++j;
      else if (j + 1 == npart) node_part[i] = j;
      else node_part[i] = 0;
    }
  }
  for (int64_t c = 0; c < m; ++c)
    for (int64_t p = Ap[c]; p < Ap[c + 1]; ++p) {
      int64_t r = Ai[p];
      if (node_part[r] == -1) { node_part[r] = node_part[c]; continue; }
      if (node_part[r] != npart && node_part[c] != npart && node_part[r] != node_part[c]) {
        if (c > r) node_part[c] = npart;
        else node_part[r] = npart;
      }
    }
  for (int k = 0; k <= npart; ++k) part_size[k] = 0;
  for (int64_t c = 0; c < m; ++c) part_size[node_part[c]]++;
  std::vector<int64_t> begin(npart + 1, 0), cur(npart + 1, 0);
  for (int k = 1; k <= npart; ++k) begin[k] = begin[k - 1] + part_size[k - 1];
  for (int64_t c = 0; c < m; ++c) {
    int64_t pos = begin[node_part[c]] + cur[node_part[c]]++;
    pinv[c] = pos;
    q[pos] = c;
  }
}

}  // namespace

// ---------------------------------------------------------------- C entry points (ctypes)
#define ORC_API extern "C" __attribute__((visibility("default")))

#define DEF_BASIC(SFX, R)                                                                          \
  ORC_API void orc_spmv_##SFX(R* out, const R* val, const int* rp, const int* ci, const R* in,     \
                              int n) {                                                             \
    spmv_(out, val, rp, ci, in, n);                                                                \
  }                                                                                                \
  ORC_API void orc_generate_level_##SFX(int n, const R* val, const int* rp, const int* ci,         \
                                        int* level) {                                              \
    gen_level_(n, val, rp, ci, level);                                                             \
  }                                                                                                \
  ORC_API void orc_get_level_##SFX(int isL, const R* val, const int* rp, const int* ci,            \
                                   int* level, int n) {                                            \
    get_level_(isL, val, rp, ci, level, n);                                                        \
  }                                                                                                \
  /* two-call protocol: first with lv==NULL to obtain lp/up (sizes), then to fill */               \
  ORC_API void orc_ilu0_##SFX(int n, const R* val, const int* rp, const int* ci, int* lp, R* lv,   \
                              int* li, int* up, R* uv, int* ui, int* level, int* n_levels,         \
                              int* levelPtr, int* nodes, R* fval, int* fcol) {                     \
    Ilu0Result<R> res;                                                                             \
    ilu0_(n, val, rp, ci, res);                                                                    \
    split_count_(n, res.fv.data(), res.fp.data(), res.fi.data(), lp, up);                          \
    if (lv) split_fill_(n, res.fv.data(), res.fp.data(), res.fi.data(), lp, lv, li, up, uv, ui);   \
    if (level) std::copy(res.level.begin(), res.level.end(), level);                               \
    if (n_levels) *n_levels = (int)res.levelPtr.size() - 1;                                        \
    if (levelPtr) std::copy(res.levelPtr.begin(), res.levelPtr.end(), levelPtr);                   \
    if (nodes) std::copy(res.nodes.begin(), res.nodes.end(), nodes);                               \
    if (fval) std::copy(res.fv.begin(), res.fv.end(), fval);                                       \
    if (fcol) std::copy(res.fi.begin(), res.fi.end(), fcol);                                       \
  }

DEF_BASIC(f32, float)
DEF_BASIC(f64, double)

ORC_API int orc_bucket_levels(int n, const int* level, int* levelPtr, int* nodes) {
  return bucket_levels(n, level, levelPtr, nodes);
}

ORC_API void orc_partition(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart,
                           int64_t nNodes, int64_t* part_size, int64_t* node_part, int64_t* pinv,
                           int64_t* q) {
  partition_(m, Ap, Ai, npart, nNodes, part_size, node_part, pinv, q);
}

// Preconditioner description passed from Python as a flat struct of pointers.
struct orc_pc {
  int kind, n;
  const void* dinv;
  const void *lv, *uv;
  const int *lp, *li, *up, *ui;
  const void *middle, *lscale, *rscale;
  const int *permRow, *permCol;
};

template <class R, class F> static Pc<R, F> make_pc(const orc_pc* d) {
  Pc<R, F> M;
  M.kind = d->kind; M.n = d->n;
  M.dinv = (const R*)d->dinv;
  M.lv = (const F*)d->lv; M.uv = (const F*)d->uv;
  M.lp = d->lp; M.li = d->li; M.up = d->up; M.ui = d->ui;
  M.middle = (const R*)d->middle; M.lscale = (const R*)d->lscale; M.rscale = (const R*)d->rscale;
  M.permRow = d->permRow; M.permCol = d->permCol;
  return M;
}

// which: 0 full (HostPrecond), 1 left, 2 right, 3 starting_value, 4 rhs (= left)
#define DEF_PC(SFX, R, F)                                                                          \
  ORC_API void orc_precond_apply_##SFX(const orc_pc* d, int which, const R* in, R* out) {          \
    Pc<R, F> M = make_pc<R, F>(d);                                                                 \
    switch (which) {                                                                               \
      case 0: M.full(in, out); break;                                                              \
      case 1: case 4: M.split_left(in, out); break;                                                \
      case 2: M.split_right(in, out); break;                                                       \
      case 3: M.split_start(in, out); break;                                                       \
    }                                                                                              \
  }                                                                                                \
  ORC_API int orc_gmres_##SFX(const R* val, const int* rp, const int* ci, R* x, const R* b, int n, \
                              int m, int* max_iter, R* tol, const orc_pc* d, R* hist,              \
                              int hist_cap, int* hist_len) {                                       \
    Pc<R, F> M = make_pc<R, F>(d);                                                                 \
    return gmres_(val, rp, ci, x, b, n, m, max_iter, tol, M, d->kind == PC_SPLIT, hist, hist_cap,  \
                  hist_len);                                                                       \
  }

DEF_PC(f32, float, float)     // reference ILU0 / DIAG / NONE path
DEF_PC(f32f64, float, double) // reference MyILUPPfloat path: FP32 vectors, FP64 factors
DEF_PC(f64, double, double)   // parity target
