// host_algos.cpp -- see host_algos.h.  All routines are written from the behaviour documented in
// SURVEY.md (appendix A) and cite the reference location whose integer artefacts they reproduce.
#include "host_algos.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <set>
#include <stdexcept>

// METIS 5 as shipped inside the CUDA toolkit (libmetis_static.a, 64-bit idx_t, no header).
extern "C" int METIS_SetDefaultOptions(int64_t* options);
extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                                   int64_t* vsize, int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec,
                                   int64_t* options, int64_t* edgecut, int64_t* part);

namespace bg {
template <class T> static void parallel_copy(std::vector<T>& dst, const T* src, size_t n) {
  dst.resize(n);  // (value-initialises: the faults of the zero fill are what is parallelised below for the big arrays)
  T* d = dst.data();
  parallel_ranges(n, [=](size_t b, size_t e) { std::copy(src + b, src + e, d + b); });
}


// ------------------------------------------------------------------ level sets
void factor_levels(int n, const int* rp, const int* ci, const float* v, int* level) {
  std::fill(level, level + n, 0);
  for (int r = 0; r < n; ++r) {
    const int lr = level[r] + 1;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
      const int c = ci[p];
      if (c > r && !is_zero(v[p]) && level[c] < lr) level[c] = lr;
    }
  }
}

template <class T>
void trsv_levels(bool lower, int n, const int* rp, const int* ci, const T* v, int* level) {
  std::fill(level, level + n, 0);
  if (lower) {
    for (int i = 1; i < n; ++i) {
      int best = 0;
      for (int p = rp[i]; p < rp[i + 1]; ++p) {
        if (is_zero(v[p]) || ci[p] >= i) break;  // padding sits at the row end; row sorted by column
        best = std::max(best, level[ci[p]] + 1);
      }
      level[i] = best;
    }
  } else {
    for (int i = n - 2; i >= 0; --i) {
      int best = 0;
      for (int p = rp[i + 1] - 1; p >= rp[i]; --p) {
        if (is_zero(v[p])) continue;
        if (ci[p] <= i) break;
        best = std::max(best, level[ci[p]] + 1);
      }
      level[i] = best;
    }
  }
}
template void trsv_levels<float>(bool, int, const int*, const int*, const float*, int*);
template void trsv_levels<double>(bool, int, const int*, const int*, const double*, int*);

int bucket_levels(int n, const int* level, std::vector<int>& levelPtr, std::vector<int>& nodes) {
  int nl = 0;
  for (int i = 0; i < n; ++i) nl = std::max(nl, level[i] + 1);
  levelPtr.assign(nl + 1, 0);
  nodes.resize(n);
  for (int i = 0; i < n; ++i) levelPtr[level[i] + 1]++;
  for (int l = 0; l < nl; ++l) levelPtr[l + 1] += levelPtr[l];
  std::vector<int> cur(levelPtr.begin(), levelPtr.end() - (nl ? 1 : 0));
  for (int i = 0; i < n; ++i) nodes[cur[level[i]]++] = i;
  return nl;
}

// ------------------------------------------------------------------ ILU helpers
namespace {
// sorted (by column) working copy of the rows, optionally restricted to the block diagonal
template <class T>
void sorted_copy(int n, const int* rp, const int* ci, const T* v, const int* block, Csr<T>& W) {
  W.n = n;
  if (!block) {  // the common case -- every row already sorted, nothing to drop -- is a plain copy
    bool sorted = true;
    for (int i = 0; i < n && sorted; ++i)
      for (int p = rp[i] + 1; p < rp[i + 1]; ++p)
        if (ci[p] < ci[p - 1]) { sorted = false; break; }
    if (sorted) {
      W.rp.assign(rp, rp + n + 1);
      parallel_copy(W.ci, ci, (size_t)rp[n]);
      parallel_copy(W.v, v, (size_t)rp[n]);
      return;
    }
  }
  W.rp.assign(n + 1, 0);
  W.ci.clear();
  W.v.clear();
  W.ci.reserve(rp[n]);
  W.v.reserve(rp[n]);
  std::vector<std::pair<int, T>> row;
  for (int i = 0; i < n; ++i) {
    row.clear();
    bool sorted = true;
    int prev = -1;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      if (block && block[ci[p]] != block[i]) continue;
      if (ci[p] < prev) sorted = false;
      prev = ci[p];
      row.emplace_back(ci[p], v[p]);
    }
    if (!sorted)
      std::stable_sort(row.begin(), row.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
    for (auto& e : row) { W.ci.push_back(e.first); W.v.push_back(e.second); }
    W.rp[i + 1] = (int)W.ci.size();
  }
}

// split a factored CSR (columns ascending) into the reference's L / U layout,
// dropping |v| < eps (splitLU_csr, src/leftILU.cu:481-541)
template <class T> void split_lu(const Csr<T>& W, Factors<T>& F) {
  const int n = W.n;
  F.L.n = F.U.n = n;
  F.L.rp.assign(n + 1, 0);
  F.U.rp.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    int nl = 0, nu = 0;
    for (int p = W.rp[i]; p < W.rp[i + 1]; ++p) {
      if (is_zero(W.v[p])) continue;
      (W.ci[p] < i ? nl : nu)++;
    }
    F.L.rp[i + 1] = F.L.rp[i] + nl + 1;
    F.U.rp[i + 1] = F.U.rp[i] + nu;
  }
  F.L.ci.resize(F.L.rp[n]); F.L.v.resize(F.L.rp[n]);
  F.U.ci.resize(F.U.rp[n]); F.U.v.resize(F.U.rp[n]);
  parallel_ranges((size_t)n, [&](size_t r0, size_t r1) {  // every row writes its own segments of L and U
    for (int i = (int)r0; i < (int)r1; ++i) {
      int a = F.L.rp[i], b = F.U.rp[i];
      for (int p = W.rp[i]; p < W.rp[i + 1]; ++p) {
        if (is_zero(W.v[p])) continue;
        if (W.ci[p] < i) { F.L.ci[a] = W.ci[p]; F.L.v[a++] = W.v[p]; }
        else { F.U.ci[b] = W.ci[p]; F.U.v[b++] = W.v[p]; }
      }
      F.L.ci[a] = i;
      F.L.v[a] = (T)1;
    }
  });
  F.unit_lower = true;
}
}  // namespace

template <class T>
void ilu0(int n, const int* rp, const int* ci, const T* v, const int* block, Factors<T>& F) {
  Csr<T> W;
  sorted_copy(n, rp, ci, v, block, W);
  std::vector<int> diag(n, -1), pos(n, -1);
  F.zero_pivots = 0;
  for (int i = 0; i < n; ++i) {
    const int lb = W.rp[i], ub = W.rp[i + 1];
    for (int p = lb; p < ub; ++p) {
      pos[W.ci[p]] = p;
      if (W.ci[p] == i) diag[i] = p;
    }
    for (int p = lb; p < ub; ++p) {
      const int k = W.ci[p];
      if (k >= i) break;
      const T ukk = diag[k] >= 0 ? W.v[diag[k]] : (T)0;
      // reference: L column of a vanishing pivot is zeroed (src/leftILU.cu:812-818)
      const T lik = is_zero(ukk) ? (T)0 : W.v[p] / ukk;
      W.v[p] = lik;
      // a_ij -= l_ik * u_kj over the part of row k right of its diagonal that row i also stores
      for (int q = (diag[k] >= 0 ? diag[k] + 1 : std::lower_bound(W.ci.begin() + W.rp[k], W.ci.begin() + W.rp[k + 1], k) - W.ci.begin());
           q < W.rp[k + 1]; ++q) {
        const int t = pos[W.ci[q]];
        if (t >= 0) W.v[t] -= lik * W.v[q];
      }
    }
    if (diag[i] < 0 || is_zero(W.v[diag[i]])) F.zero_pivots++;
    for (int p = lb; p < ub; ++p) pos[W.ci[p]] = -1;
  }
  split_lu(W, F);
}
template void ilu0<float>(int, const int*, const int*, const float*, const int*, Factors<float>&);
template void ilu0<double>(int, const int*, const int*, const double*, const int*, Factors<double>&);

// ILU(k): level-of-fill incomplete factorisation.  The symbolic phase follows the rule of ITSOL's lofC, which the
// reference vendors (src/iluk.cpp:193-335): entries of A have level 0, eliminating row i with pivot row k gives
// position (i, j) the level lev(i,k) + lev(k,j) + 1, the smallest level offered wins, positions above `lfil` are not
// created, pivots are taken in ascending column order (fill left of the diagonal becomes a pivot itself), the
// diagonal is always kept.  The numeric phase is the IKJ sweep of ilu0 above on that pattern (multipliers divided by
// the pivot, where ITSOL multiplies by a stored reciprocal), so lfil = 0 reproduces ilu0 bit for bit.
template <class T>
void iluk(int n, const int* rp, const int* ci, const T* v, int lfil, const int* block, Factors<T>& F) {
  Csr<T> A;
  sorted_copy(n, rp, ci, v, block, A);
  // factored rows: columns ascending, values, level of every position; udiag[i] = position of the diagonal
  Csr<T> W;
  W.n = n;
  W.rp.assign((size_t)n + 1, 0);
  W.ci.reserve(A.ci.size() * (size_t)(lfil > 0 ? 2 : 1));
  W.v.reserve(A.ci.size() * (size_t)(lfil > 0 ? 2 : 1));
  std::vector<int> lev;
  lev.reserve(W.ci.capacity());
  std::vector<int> udiag(n, -1);
  std::vector<int> slot(n, -1);      // column -> index in the row under construction
  std::vector<int> cols, levs;       // the row under construction, in discovery order
  std::vector<T> vals;
  std::vector<int> heap;             // columns left of the diagonal still to be used as pivots (min-heap)
  std::vector<int> order;
  F.zero_pivots = 0;
  auto heap_cmp = [](int a, int b) { return a > b; };
  for (int i = 0; i < n; ++i) {
    cols.clear(); levs.clear(); vals.clear(); heap.clear();
    bool has_diag = false;
    for (int p = A.rp[i]; p < A.rp[i + 1]; ++p) {
      const int c = A.ci[p];
      if (c == i) has_diag = true;
      slot[c] = (int)cols.size();
      cols.push_back(c); levs.push_back(0); vals.push_back(A.v[p]);
      if (c < i) heap.push_back(c);
    }
    if (!has_diag) { slot[i] = (int)cols.size(); cols.push_back(i); levs.push_back(0); vals.push_back((T)0); }
    // symbolic: the complete pattern of row i first -- a position refused at one pivot (level too high) may be
    // created by a later one, and the numeric sweep must then see the earlier pivot's contribution as well
    std::make_heap(heap.begin(), heap.end(), heap_cmp);
    while (!heap.empty()) {
      std::pop_heap(heap.begin(), heap.end(), heap_cmp);
      const int k = heap.back();
      heap.pop_back();
      const int lk = levs[slot[k]];
      for (int q = udiag[k] + 1; q < W.rp[k + 1]; ++q) {
        const int c = W.ci[q];
        const int it = lev[q] + lk + 1;
        if (it > lfil) continue;
        const int sc = slot[c];
        if (sc < 0) {
          slot[c] = (int)cols.size();
          cols.push_back(c); levs.push_back(it); vals.push_back((T)0);
          if (c < i) { heap.push_back(c); std::push_heap(heap.begin(), heap.end(), heap_cmp); }
        } else if (it < levs[sc]) {
          levs[sc] = it;
        }
      }
    }
    order.resize(cols.size());
    for (size_t q = 0; q < cols.size(); ++q) order[q] = (int)q;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cols[a] < cols[b]; });
    // numeric: a_ij -= l_ik u_kj over the pivots k < i of the row in ascending order, on the positions the row stores
    for (int q : order) {
      const int k = cols[q];
      if (k >= i) break;
      const T ukk = W.v[udiag[k]];
      const T lik = is_zero(ukk) ? (T)0 : vals[q] / ukk;  // a vanishing pivot zeroes the column like ilu0
      vals[q] = lik;
      for (int r = udiag[k] + 1; r < W.rp[k + 1]; ++r) {
        const int sc = slot[W.ci[r]];
        if (sc >= 0) vals[sc] -= lik * W.v[r];
      }
    }
    for (int q : order) {
      if (cols[q] == i) udiag[i] = (int)W.ci.size();
      W.ci.push_back(cols[q]); W.v.push_back(vals[q]); lev.push_back(levs[q]);
      slot[cols[q]] = -1;
    }
    W.rp[(size_t)i + 1] = (int)W.ci.size();
    if (is_zero(W.v[udiag[i]])) F.zero_pivots++;
  }
  split_lu(W, F);
}
template void iluk<float>(int, const int*, const int*, const float*, int, const int*, Factors<float>&);
template void iluk<double>(int, const int*, const int*, const double*, int, const int*, Factors<double>&);

template <class T>
void ilut(int n, const int* rp, const int* ci, const T* v, double tau, int lfil, const int* block,
          Factors<T>& F) {
  Csr<T> A;
  sorted_copy(n, rp, ci, v, block, A);
  // U rows are kept diagonal-first; L rows strict-lower ascending.
  std::vector<std::vector<std::pair<int, T>>> Lr(n), Ur(n);
  std::vector<T> w(n, (T)0);
  std::vector<char> in_w(n, 0);
  std::vector<int> idx;
  std::set<int> lower;  // pending lower columns, ascending
  F.zero_pivots = 0;
  for (int i = 0; i < n; ++i) {
    idx.clear();
    lower.clear();
    double nrm = 0;
    for (int p = A.rp[i]; p < A.rp[i + 1]; ++p) {
      const int c = A.ci[p];
      w[c] = A.v[p];
      in_w[c] = 1;
      idx.push_back(c);
      if (c < i) lower.insert(c);
      nrm += (double)A.v[p] * (double)A.v[p];
    }
    const int row_nnz = A.rp[i + 1] - A.rp[i];
    nrm = row_nnz ? std::sqrt(nrm / row_nnz) : 0.0;
    const double drop = tau * nrm;
    if (!in_w[i]) { in_w[i] = 1; w[i] = 0; idx.push_back(i); }
    while (!lower.empty()) {
      const int k = *lower.begin();
      lower.erase(lower.begin());
      const T ukk = Ur[k].empty() ? (T)0 : Ur[k][0].second;
      T lik = is_zero(ukk) ? (T)0 : w[k] / ukk;
      if (std::fabs((double)lik) < drop) lik = 0;
      w[k] = lik;
      if (lik == (T)0) continue;
      for (size_t q = 1; q < Ur[k].size(); ++q) {
        const int c = Ur[k][q].first;
        if (!in_w[c]) {
          in_w[c] = 1;
          w[c] = 0;
          idx.push_back(c);
          if (c < i) lower.insert(c);
        }
        w[c] -= lik * Ur[k][q].second;
      }
    }
    std::vector<std::pair<int, T>> lo, up;
    T dii = w[i];
    for (int c : idx) {
      const T val = w[c];
      if (c < i) { if (val != (T)0 && std::fabs((double)val) >= drop) lo.emplace_back(c, val); }
      else if (c > i) { if (std::fabs((double)val) >= drop) up.emplace_back(c, val); }
      w[c] = 0;
      in_w[c] = 0;
    }
    auto keep_largest = [&](std::vector<std::pair<int, T>>& e) {
      if ((int)e.size() > lfil) {
        std::nth_element(e.begin(), e.begin() + lfil, e.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) {
          return std::fabs((double)a.second) > std::fabs((double)b.second);
        });
        e.resize(lfil);
      }
      std::sort(e.begin(), e.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
    };
    keep_largest(lo);
    keep_largest(up);
    if (is_zero(dii)) {  // keep the factor usable: replace a vanishing pivot by the drop scale
      F.zero_pivots++;
      dii = (T)(nrm > 0 ? (tau > 0 ? tau * nrm : nrm) : 1.0);
      if (is_zero(dii)) dii = (T)1;
    }
    Lr[i] = std::move(lo);
    Ur[i].emplace_back(i, dii);
    for (auto& e : up) Ur[i].push_back(e);
  }
  F.L.n = F.U.n = n;
  F.L.rp.assign(n + 1, 0);
  F.U.rp.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    F.L.rp[i + 1] = F.L.rp[i] + (int)Lr[i].size() + 1;
    F.U.rp[i + 1] = F.U.rp[i] + (int)Ur[i].size();
  }
  F.L.ci.resize(F.L.rp[n]); F.L.v.resize(F.L.rp[n]);
  F.U.ci.resize(F.U.rp[n]); F.U.v.resize(F.U.rp[n]);
  for (int i = 0; i < n; ++i) {
    int a = F.L.rp[i];
    for (auto& e : Lr[i]) { F.L.ci[a] = e.first; F.L.v[a++] = e.second; }
    F.L.ci[a] = i; F.L.v[a] = (T)1;
    int b = F.U.rp[i];
    for (auto& e : Ur[i]) { F.U.ci[b] = e.first; F.U.v[b++] = e.second; }
  }
  F.unit_lower = true;
}
template void ilut<float>(int, const int*, const int*, const float*, double, int, const int*, Factors<float>&);
template void ilut<double>(int, const int*, const int*, const double*, double, int, const int*, Factors<double>&);

// ------------------------------------------------------------------ AINV
namespace {
template <class T> struct SpVec { std::vector<int> idx; std::vector<T> val; };  // ascending indices
// y = x - alpha * z on sparse vectors, dropping created entries below tol (never the entry `keep`)
template <class T> void sp_axpy_drop(SpVec<T>& x, double alpha, const SpVec<T>& z, double tol, int keep, SpVec<T>& tmp) {
  tmp.idx.clear(); tmp.val.clear();
  size_t a = 0, b = 0;
  while (a < x.idx.size() || b < z.idx.size()) {
    int k; double val;
    if (b >= z.idx.size() || (a < x.idx.size() && x.idx[a] < z.idx[b])) { k = x.idx[a]; val = (double)x.val[a]; ++a; }
    else if (a >= x.idx.size() || z.idx[b] < x.idx[a]) { k = z.idx[b]; val = -alpha * (double)z.val[b]; ++b; if (std::fabs(val) < tol && k != keep) continue; }
    else { k = x.idx[a]; val = (double)x.val[a] - alpha * (double)z.val[b]; ++a; ++b; if (std::fabs(val) < tol && k != keep) continue; }
    tmp.idx.push_back(k); tmp.val.push_back((T)val);
  }
  x.idx.swap(tmp.idx); x.val.swap(tmp.val);
}
// one biconjugation sweep: columns z_i with  z_i = e_i - sum_{j<i} (a_j^T z_i / p_j) z_j,  a_j^T = row j of M;
// candidates j are found through the columns of M (MT = M^T in CSR) touched by the pattern of z_i
template <class T>
void biconj(int n, const Csr<T>& M, const Csr<T>& MT, double tol, std::vector<SpVec<T>>& Z, std::vector<double>& p) {
  Z.assign(n, SpVec<T>());
  p.assign(n, 0.0);
  std::set<int> cand;
  SpVec<T> tmp;
  std::vector<T> dense(n, (T)0);
  for (int i = 0; i < n; ++i) {
    SpVec<T>& z = Z[i];
    z.idx.assign(1, i); z.val.assign(1, (T)1);
    cand.clear();
    for (int q = MT.rp[i]; q < MT.rp[i + 1]; ++q) if (MT.ci[q] < i) cand.insert(MT.ci[q]);
    while (!cand.empty()) {
      const int j = *cand.begin();
      cand.erase(cand.begin());
      for (size_t t = 0; t < z.idx.size(); ++t) dense[z.idx[t]] = z.val[t];
      double dot = 0;
      for (int q = M.rp[j]; q < M.rp[j + 1]; ++q) dot += (double)M.v[q] * (double)dense[M.ci[q]];
      for (size_t t = 0; t < z.idx.size(); ++t) dense[z.idx[t]] = (T)0;
      if (dot == 0.0 || p[j] == 0.0) continue;
      const size_t before = z.idx.size();
      std::vector<int> old(z.idx);
      sp_axpy_drop(z, dot / p[j], Z[j], tol, i, tmp);
      if (z.idx.size() != before || z.idx != old)  // new pattern entries k (< j) open later rows j' in (j, i)
        for (int k : z.idx)
          if (!std::binary_search(old.begin(), old.end(), k))
            for (int q = MT.rp[k]; q < MT.rp[k + 1]; ++q) if (MT.ci[q] > j && MT.ci[q] < i) cand.insert(MT.ci[q]);
    }
    double dot = 0;
    for (size_t t = 0; t < z.idx.size(); ++t) dense[z.idx[t]] = z.val[t];
    for (int q = M.rp[i]; q < M.rp[i + 1]; ++q) dot += (double)M.v[q] * (double)dense[M.ci[q]];
    for (size_t t = 0; t < z.idx.size(); ++t) dense[z.idx[t]] = (T)0;
    p[i] = dot;
  }
}
template <class T> Csr<T> transpose_csr(const Csr<T>& M) {
  Csr<T> Tm;
  Tm.n = M.n;
  Tm.rp.assign(M.n + 1, 0);
  for (int c : M.ci) Tm.rp[c + 1]++;
  for (int i = 0; i < M.n; ++i) Tm.rp[i + 1] += Tm.rp[i];
  Tm.ci.resize(M.ci.size()); Tm.v.resize(M.v.size());
  std::vector<int> cur(Tm.rp.begin(), Tm.rp.end() - 1);
  for (int i = 0; i < M.n; ++i)
    for (int q = M.rp[i]; q < M.rp[i + 1]; ++q) { const int d = cur[M.ci[q]]++; Tm.ci[d] = i; Tm.v[d] = M.v[q]; }
  return Tm;
}
}  // namespace

template <class T>
void ainv_biconjugation(int n, const int* rp, const int* ci, const T* v, double droptol, Factors<T>& F) {
  Csr<T> A;
  sorted_copy(n, rp, ci, v, (const int*)nullptr, A);
  const Csr<T> AT = transpose_csr(A);
  std::vector<SpVec<T>> Z, W;
  std::vector<double> p, q;
  biconj(n, A, AT, droptol, Z, p);   // z_i: A-conjugation against the rows of A
  biconj(n, AT, A, droptol, W, q);   // w_i: the same against the rows of A^T (= columns of A)
  F.zero_pivots = 0;
  // F.L = D^-1 W^T: row i = w_i / p_i
  F.L.n = F.U.n = n;
  F.L.rp.assign(1, 0); F.L.ci.clear(); F.L.v.clear();
  for (int i = 0; i < n; ++i) {
    double piv = p[i];
    if (std::fabs(piv) < (double)kEps) { F.zero_pivots++; piv = piv < 0 ? -(double)kEps : (double)kEps; }
    for (size_t t = 0; t < W[i].idx.size(); ++t) { F.L.ci.push_back(W[i].idx[t]); F.L.v.push_back((T)((double)W[i].val[t] / piv)); }
    F.L.rp.push_back((int)F.L.ci.size());
  }
  // F.U = Z: the z_i are its columns
  Csr<T> ZT;
  ZT.n = n;
  ZT.rp.assign(1, 0);
  for (int i = 0; i < n; ++i) {
    for (size_t t = 0; t < Z[i].idx.size(); ++t) { ZT.ci.push_back(Z[i].idx[t]); ZT.v.push_back(Z[i].val[t]); }
    ZT.rp.push_back((int)ZT.ci.size());
  }
  F.U = transpose_csr(ZT);
  F.unit_lower = false;
}
template void ainv_biconjugation<float>(int, const int*, const int*, const float*, double, Factors<float>&);
template void ainv_biconjugation<double>(int, const int*, const int*, const double*, double, Factors<double>&);

// ------------------------------------------------------------------ partitioning
namespace {
void finish_partition(int64_t m, int npart, const int64_t* node_part, int64_t* part_size, int64_t* pinv,
                      int64_t* q) {
  for (int k = 0; k <= npart; ++k) part_size[k] = 0;
  for (int64_t j = 0; j < m; ++j) part_size[node_part[j]]++;
  std::vector<int64_t> next(npart + 1, 0);
  for (int k = 1; k <= npart; ++k) next[k] = next[k - 1] + part_size[k - 1];
  for (int64_t j = 0; j < m; ++j) {
    const int64_t slot = next[node_part[j]]++;
    pinv[j] = slot;
    q[slot] = j;
  }
}
}  // namespace

void partition_ref(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, int64_t nNodes,
                   int64_t* part_size, int64_t* node_part, int64_t* pinv, int64_t* q) {
  // contiguous chunks of floor(nNodes/npart) rows, remainder to the last part (partition.cpp:31-45)
  const int64_t avg = nNodes / npart;
  for (int64_t i = 0; i < m; ++i) node_part[i] = -1;
  int64_t cur = 0;
  for (int64_t i = 0; i < nNodes; ++i) {
    if (i >= (cur + 1) * avg) {
      if (cur + 1 < npart) ++cur;
      else if (cur + 1 > npart) { node_part[i] = 0; continue; }
    }
    node_part[i] = cur;
  }
  // single order-dependent sweep over the CSC entries (partition.cpp:54-69): unknowns beyond
  // nNodes inherit the part of the first column touching them; of every cut edge the larger
  // index goes to the separator part `npart`.
  for (int64_t j = 0; j < m; ++j)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; ++p) {
      const int64_t r = Ai[p];
      if (node_part[r] == -1) { node_part[r] = node_part[j]; continue; }
      if (node_part[r] == npart || node_part[j] == npart || node_part[r] == node_part[j]) continue;
      if (j > r) node_part[j] = npart;
      else node_part[r] = npart;
    }
  finish_partition(m, npart, node_part, part_size, pinv, q);
}

void partition3_post(int64_t m, const int64_t* Ap, const int64_t* Ai, int npart, const int64_t* part,
                     int64_t* part_size, int64_t* node_part, int64_t* pinv, int64_t* q) {
  for (int64_t j = 0; j < m; ++j) node_part[j] = part[j];
  std::vector<char> sep(m, 0);  // both endpoints of every cut edge (partition3.cpp:79-90)
  for (int64_t j = 0; j < m; ++j)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; ++p)
      if (part[Ai[p]] != part[j]) sep[Ai[p]] = sep[j] = 1;
  for (int64_t j = 0; j < m; ++j)
    if (sep[j]) node_part[j] = npart;
  finish_partition(m, npart, node_part, part_size, pinv, q);
}

int block_partition(int method, int n, const int* rp, const int* ci, int nblocks, const int* dims,
                    int* block_of_row) {
  if (nblocks < 1) nblocks = 1;
  if (method == 2 || method == 3) {
    const int nx = dims[0], ny = dims[1], nz = method == 3 ? dims[2] : 1;
    const int tx = std::max(1, dims[3]), ty = std::max(1, dims[4]), tz = method == 3 ? std::max(1, dims[5]) : 1;
    if ((int64_t)nx * ny * nz != n) throw std::invalid_argument("grid dims do not match n");
    const int bx = (nx + tx - 1) / tx, by = (ny + ty - 1) / ty, bz = (nz + tz - 1) / tz;
    for (int z = 0; z < nz; ++z)
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x)
          block_of_row[((int64_t)z * ny + y) * nx + x] = ((z / tz) * by + (y / ty)) * bx + (x / tx);
    return bx * by * bz;
  }
  if (method == 1 && nblocks > 1) {
    // symmetrised off-diagonal pattern -> METIS k-way
    std::vector<int64_t> deg(n + 1, 0);
    for (int i = 0; i < n; ++i)
      for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (ci[p] != i) { deg[i + 1]++; deg[ci[p] + 1]++; }
    for (int i = 0; i < n; ++i) deg[i + 1] += deg[i];
    std::vector<int64_t> adj(deg[n]), cur(deg.begin(), deg.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (ci[p] != i) { adj[cur[i]++] = ci[p]; adj[cur[ci[p]]++] = i; }
    std::vector<int64_t> xadj(n + 1, 0);
    int64_t w = 0;
    for (int i = 0; i < n; ++i) {  // sort + unique each list
      int64_t b = deg[i], e = deg[i + 1];
      std::sort(adj.begin() + b, adj.begin() + e);
      int64_t u = std::unique(adj.begin() + b, adj.begin() + e) - (adj.begin() + b);
      std::copy(adj.begin() + b, adj.begin() + b + u, adj.begin() + w);
      xadj[i] = w;
      w += u;
    }
    xadj[n] = w;
    int64_t nv = n, ncon = 1, np = nblocks, cut = 0, options[40];
    METIS_SetDefaultOptions(options);
    std::vector<int64_t> part(n, 0);
    int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &np, nullptr,
                                 nullptr, options, &cut, part.data());
    if (rc != 1) throw std::runtime_error("METIS_PartGraphKway failed");
    for (int i = 0; i < n; ++i) block_of_row[i] = (int)part[i];
    return nblocks;
  }
  // contiguous: floor(n/nblocks) rows per block, remainder into the last one (partition.cpp:31-45)
  const int avg = std::max(1, n / nblocks);
  for (int i = 0; i < n; ++i) block_of_row[i] = std::min(i / avg, nblocks - 1);
  return nblocks;
}

// ------------------------------------------------------------------ packers
void trsv_levels_structural(bool lower, int n, const int* rp, const int* ci, int* level) {
  std::fill(level, level + n, 0);
  if (lower) {
    for (int i = 0; i < n; ++i) {
      int best = 0;
      for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (ci[p] < i) best = std::max(best, level[ci[p]] + 1);
      level[i] = best;
    }
  } else {
    for (int i = n - 1; i >= 0; --i) {
      int best = 0;
      for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (ci[p] > i) best = std::max(best, level[ci[p]] + 1);
      level[i] = best;
    }
  }
}

template <class T>
void pack_syncfree_stream(bool lower, bool unit, const Csr<T>& M, SyncFreeStream<T>& out) {
  const int n = M.n;
  out.n = n;
  out.level.resize(n);
  out.sched_level.resize(n);
  trsv_levels(lower, n, M.rp.data(), M.ci.data(), M.v.data(), out.level.data());
  trsv_levels_structural(lower, n, M.rp.data(), M.ci.data(), out.sched_level.data());
  std::vector<int> lev_ptr;
  out.n_levels = bucket_levels(n, out.sched_level.data(), lev_ptr, out.rows);
  // inside a level: rows that need the same number of 4-entry groups share slices (a slice is as long as its longest
  // row); stable, so rows of equal length keep their order (stencil factors are left as they are)
  for (int l = 0; l < out.n_levels; ++l)
    std::stable_sort(out.rows.begin() + lev_ptr[l], out.rows.begin() + lev_ptr[l + 1], [&](int a, int b) {
      return (M.rp[a + 1] - M.rp[a] + 2) / 4 > (M.rp[b + 1] - M.rp[b] + 2) / 4;  // ceil((len - 1) / 4), len counts the diagonal
    });
  out.meta.clear();
  out.level_slices.assign(out.n_levels, 0);
  out.max_K = out.max_level_rows = 0;
  int64_t off32 = 0;
  for (int l = 0; l < out.n_levels; ++l) {
    out.max_level_rows = std::max(out.max_level_rows, lev_ptr[l + 1] - lev_ptr[l]);
    out.level_slices[l] = (lev_ptr[l + 1] - lev_ptr[l] + 31) / 32;
    for (int s0 = lev_ptr[l]; s0 < lev_ptr[l + 1]; s0 += 32) {
      const int R = std::min(32, lev_ptr[l + 1] - s0);
      int K = 0;
      for (int t = 0; t < R; ++t) {
        const int r = out.rows[s0 + t];
        K = std::max(K, M.rp[r + 1] - M.rp[r] - 1);
      }
      if (K >= (1 << 23) || off32 + K > 0x7fffffffLL) throw std::invalid_argument("factor too large for the sync-free stream");
      out.max_K = std::max(out.max_K, K);
      const int m[4] = {s0, R | (K << 8), l, (int)off32};
      out.meta.insert(out.meta.end(), m, m + 4);
      off32 += K;
    }
  }
  out.n_slices = (int)out.meta.size() / 4;
  if (out.n_slices) out.meta.insert(out.meta.end(), out.meta.end() - 4, out.meta.end());
  else out.meta.assign(4, 0);
  out.cols.assign((size_t)off32 * 32, -1);
  out.vals.assign((size_t)off32 * 32, (T)0);
  out.diag.assign(n, (T)1);
  std::vector<std::pair<int64_t, int>> ord;  // (level of the operand's row, column) -> position in the row
  for (int q = 0; q < out.n_slices; ++q) {
    const int s0 = out.meta[4 * q], R = out.meta[4 * q + 1] & 255, K = out.meta[4 * q + 1] >> 8;
    const int64_t off = (int64_t)out.meta[4 * q + 3] * 32;
    for (int t = 0; t < R; ++t) {
      const int r = out.rows[s0 + t];
      const int lb = M.rp[r], ub = M.rp[r + 1];
      if (ub <= lb) throw std::invalid_argument("triangular factor row without a diagonal");
      const int dpos = lower ? ub - 1 : lb;
      if (!unit) out.diag[s0 + t] = M.v[dpos];
      ord.clear();
      for (int p = lb; p < ub; ++p) {
        if (p == dpos) continue;
        if (lower ? M.ci[p] >= r : M.ci[p] <= r) throw std::invalid_argument("entry on the wrong side of the diagonal");
        ord.emplace_back(((int64_t)out.sched_level[M.ci[p]] << 32) | (uint32_t)M.ci[p], p);
      }
      std::sort(ord.begin(), ord.end());
      int k = K - (int)ord.size();  // right-aligned
      for (const auto& e : ord) {
        out.cols[off + (int64_t)k * 32 + t] = M.ci[e.second];
        out.vals[off + (int64_t)k * 32 + t] = M.v[e.second];
        ++k;
      }
    }
  }
}
template void pack_syncfree_stream<float>(bool, bool, const Csr<float>&, SyncFreeStream<float>&);
template void pack_syncfree_stream<double>(bool, bool, const Csr<double>&, SyncFreeStream<double>&);

template <class T> bool detect_lattice2d(const Csr<T>& L, const Csr<T>& U, int* nx_out) {
  const int n = L.n;
  if (n < 4 || U.n != n) return false;
  int nx = 0;
  for (int r = 0; r < n && !nx; ++r)
    for (int p = L.rp[r]; p < L.rp[r + 1] - 1; ++p)
      if (L.ci[p] != r - 1) { nx = r - L.ci[p]; break; }
  if (nx < 2 || n % nx) return false;
  for (int r = 0; r < n; ++r) {
    const int x = r % nx;
    if (L.rp[r + 1] <= L.rp[r] || L.ci[L.rp[r + 1] - 1] != r) return false;
    if (U.rp[r + 1] <= U.rp[r] || U.ci[U.rp[r]] != r || U.v[U.rp[r]] == (T)0) return false;
    for (int p = L.rp[r]; p < L.rp[r + 1] - 1; ++p) {
      const int c = L.ci[p];
      if (!((x > 0 && c == r - 1) || c == r - nx)) return false;
    }
    for (int p = U.rp[r] + 1; p < U.rp[r + 1]; ++p) {
      const int c = U.ci[p];
      if (!((x < nx - 1 && c == r + 1) || c == r + nx)) return false;
    }
  }
  *nx_out = nx;
  return true;
}
template bool detect_lattice2d<float>(const Csr<float>&, const Csr<float>&, int*);
template bool detect_lattice2d<double>(const Csr<double>&, const Csr<double>&, int*);

template <class T> void pack_lattice_global(const Csr<T>& L, const Csr<T>& U, int nx, LatticeGlobal<T>& out) {
  const int n = L.n, ny = n / nx;
  out.nx = nx; out.ny = ny;
  out.strips = (nx + 31) / 32;
  const size_t slots = (size_t)out.strips * ny * 32;
  out.lower.assign(slots * 2, (T)0);
  out.upper.assign(slots * 2, (T)0);
  out.udinv.assign(slots, (T)0);
  parallel_ranges((size_t)n, [&](size_t r0, size_t r1) {  // a row owns its slot q
    for (int r = (int)r0; r < (int)r1; ++r) {
      const int X = r % nx, y = r / nx;
      const size_t q = ((size_t)(X / 32) * ny + y) * 32 + X % 32;
      for (int p = L.rp[r]; p < L.rp[r + 1] - 1; ++p) out.lower[2 * q + (L.ci[p] == r - 1 && X > 0 ? 0 : 1)] = L.v[p];
      const double di = 1.0 / (double)U.v[U.rp[r]];
      out.udinv[q] = (T)di;
      for (int p = U.rp[r] + 1; p < U.rp[r + 1]; ++p)
        out.upper[2 * q + (U.ci[p] == r + 1 && X < nx - 1 ? 0 : 1)] = (T)((double)U.v[p] * di);
    }
  });
}
template void pack_lattice_global<float>(const Csr<float>&, const Csr<float>&, int, LatticeGlobal<float>&);
template void pack_lattice_global<double>(const Csr<double>&, const Csr<double>&, int, LatticeGlobal<double>&);

template <class T>
void pack_fused_block_stream(bool unit_lower, const Csr<T>& L, const Csr<T>& U, int nblocks, const int* block_of_row,
                             int chunk_target, FusedBlockStream<T>& out, bool allow_lattice) {
  if (!unit_lower) throw std::invalid_argument("the block schedule needs a unit lower factor");
  const int n = L.n;
  out.n = n;
  out.nblocks = nblocks;
  // block-major renumbering, ascending old id inside a block
  std::vector<int> bptr(nblocks + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (block_of_row[i] < 0 || block_of_row[i] >= nblocks) throw std::invalid_argument("block id out of range");
    bptr[block_of_row[i] + 1]++;
  }
  for (int b = 0; b < nblocks; ++b) bptr[b + 1] += bptr[b];
  out.perm.resize(n);
  std::vector<int> local(n);
  {
    std::vector<int> cur(bptr.begin(), bptr.end() - 1);
    for (int i = 0; i < n; ++i) {
      const int b = block_of_row[i];
      local[i] = cur[b] - bptr[b];
      out.perm[cur[b]++] = i;
    }
  }
  out.max_rows = 0;
  for (int b = 0; b < nblocks; ++b) out.max_rows = std::max(out.max_rows, bptr[b + 1] - bptr[b]);
  if (out.max_rows > 65535) throw std::invalid_argument("block too large for 16-bit local ids");
  if (chunk_target < 512) chunk_target = 512;
  // reciprocal pivots (old numbering for the rescaling below, block-major for the device)
  std::vector<double> d(n), di(n);
  for (int i = 0; i < n; ++i) {
    d[i] = (double)U.v[U.rp[i]];
    di[i] = 1.0 / d[i];
  }
  out.dinv.resize(n);
  for (int i = 0; i < n; ++i) out.dinv[i] = (T)di[out.perm[i]];
  out.blk.assign(nblocks, BlockDesc{0, 0, 0, 0, 0, 0, 0, 0});
  out.chunk_off.assign(1, 0u);
  out.stream.clear();
  out.max_levels_L = out.max_levels_U = out.max_chunk_bytes = out.max_chunks = 0;
  out.total_steps = out.total_slots = 0;

  std::vector<int> lvl(out.max_rows), order;
  std::vector<int> cnt;
  struct Step { int first, count; };  // into `order`
  std::vector<Step> steps;
  auto chunk_bytes = [](int ns, int K, int nsteps) {
    const int nsA = (ns + 7) / 8 * 8, stA = (nsteps + 16) / 16 * 16, KQ = (K + 4) / 4;
    return 16 + stA + 8 * KQ * nsA + (int)sizeof(T) * K * nsA;
  };
  // emits the chunks of one factor of one block
  auto emit_factor = [&](const Csr<T>& M, bool lower, int base, int nb, int* n_levels) {
    int nl = 0;
    auto row_level = [&](int li) {
      const int r = out.perm[base + li];
      const int dpos = lower ? M.rp[r + 1] - 1 : M.rp[r];
      int best = 0;
      for (int p = M.rp[r]; p < M.rp[r + 1]; ++p) {
        if (p == dpos) continue;
        const int c = M.ci[p];
        if (block_of_row[c] != block_of_row[r]) throw std::invalid_argument("factor couples two blocks");
        best = std::max(best, lvl[local[c]] + 1);
      }
      lvl[li] = best;
      nl = std::max(nl, best + 1);
    };
    if (lower) for (int li = 0; li < nb; ++li) row_level(li);
    else for (int li = nb - 1; li >= 0; --li) row_level(li);
    *n_levels = nl;
    cnt.assign(nl + 1, 0);
    for (int li = 0; li < nb; ++li) cnt[lvl[li] + 1]++;
    for (int l = 0; l < nl; ++l) cnt[l + 1] += cnt[l];
    order.resize(nb);
    {
      std::vector<int> cur(cnt.begin(), cnt.end() - (nl ? 1 : 0));
      for (int li = 0; li < nb; ++li) order[cur[lvl[li]]++] = li;
    }
    steps.clear();
    for (int l = 0; l < nl; ++l)
      for (int s = cnt[l]; s < cnt[l + 1]; s += 32) steps.push_back(Step{s, std::min(32, cnt[l + 1] - s)});
    auto row_K = [&](int li) {
      const int r = out.perm[base + li];
      return M.rp[r + 1] - M.rp[r] - 1;
    };
    size_t si = 0;
    while (si < steps.size()) {
      // greedy: whole steps while the record stays within the target (always at least one step)
      int ns = 0, K = 0, nst = 0;
      size_t sj = si;
      while (sj < steps.size() && nst < 255) {
        int K2 = K;
        for (int q = 0; q < steps[sj].count; ++q) K2 = std::max(K2, row_K(order[steps[sj].first + q]));
        if (nst > 0 && chunk_bytes(ns + steps[sj].count, K2, nst + 1) > chunk_target) break;
        ns += steps[sj].count; K = K2; ++nst; ++sj;
      }
      const int nsA = (ns + 7) / 8 * 8, stA = (nst + 16) / 16 * 16, KQ = (K + 4) / 4;
      const int bytes = (chunk_bytes(ns, K, nst) + 15) / 16 * 16;
      const size_t off = out.stream.size();
      out.stream.resize(off + bytes, 0);
      unsigned char* rec = out.stream.data() + off;
      uint16_t* hdr = reinterpret_cast<uint16_t*>(rec);
      hdr[0] = (uint16_t)ns; hdr[1] = (uint16_t)K; hdr[2] = (uint16_t)nst; hdr[3] = (uint16_t)nsA;
      unsigned char* sc8 = rec + 16;
      uint16_t* quads = reinterpret_cast<uint16_t*>(rec + 16 + stA);          // [KQ][nsA][4]
      T* vals = reinterpret_cast<T*>(rec + 16 + stA + 8 * (size_t)KQ * nsA);  // pairs then the odd one
      auto col_slot = [&](int k, int t) -> uint16_t& {  // k-th off-diagonal column of slot t
        const int e = k + 1;                             // element 0 of quad 0 is the row id
        return quads[((size_t)(e / 4) * nsA + t) * 4 + e % 4];
      };
      auto val_slot = [&](int k, int t) -> T& {
        if (k == K - 1 && (K & 1)) return vals[(size_t)(K / 2) * 2 * nsA + t];
        return vals[((size_t)(k / 2) * nsA + t) * 2 + k % 2];
      };
      int t = 0;
      for (size_t s = si; s < sj; ++s) {
        sc8[s - si] = (unsigned char)steps[s].count;
        for (int q = 0; q < steps[s].count; ++q, ++t) {
          const int li = order[steps[s].first + q], r = out.perm[base + li];
          const int lb = M.rp[r], ub = M.rp[r + 1], dpos = lower ? ub - 1 : lb;
          quads[(size_t)t * 4] = (uint16_t)li;
          int k = 0;
          for (int p = lb; p < ub; ++p) {
            if (p == dpos) continue;
            const int c = M.ci[p];
            col_slot(k, t) = (uint16_t)local[c];
            // Lh = D^-1 L D, Ut = D^-1 U
            val_slot(k, t) = (T)(lower ? (double)M.v[p] * di[r] * d[c] : (double)M.v[p] * di[r]);
            ++k;
          }
          for (; k < 4 * KQ - 1; ++k) {
            col_slot(k, t) = (uint16_t)li;
            if (k < K) val_slot(k, t) = (T)0;
          }
        }
      }
      out.chunk_off.push_back((uint32_t)(out.stream.size() / 16));
      out.max_chunk_bytes = std::max(out.max_chunk_bytes, bytes);
      out.total_steps += nst;
      out.total_slots += ns;
      si = sj;
    }
  };
  // ---- lattice layout: is every block a tx x ty tile with west/south (L) and east/north (U) couplings only?
  std::vector<int> ltx(nblocks, 0), lty(nblocks, 0);
  auto lattice_dims = [&](int base, int nb, int& tx, int& ty) -> bool {
    tx = 0;
    for (int li = 0; li < nb && !tx; ++li) {
      const int r = out.perm[base + li];
      for (int p = L.rp[r]; p < L.rp[r + 1] - 1; ++p) {
        if (block_of_row[L.ci[p]] != block_of_row[r]) throw std::invalid_argument("factor couples two blocks");
        const int lc = local[L.ci[p]];
        if (lc != li - 1) { tx = li - lc; break; }
      }
    }
    if (!tx) tx = nb;  // a single lattice row
    if (tx < 1 || tx > 32 || nb % tx) return false;
    ty = nb / tx;
    for (int li = 0; li < nb; ++li) {
      const int r = out.perm[base + li], x = li % tx, y = li / tx;
      for (int p = L.rp[r]; p < L.rp[r + 1] - 1; ++p) {
        const int lc = local[L.ci[p]];
        if (!((x > 0 && lc == li - 1) || (y > 0 && lc == li - tx))) return false;
      }
      for (int p = U.rp[r] + 1; p < U.rp[r + 1]; ++p) {
        if (block_of_row[U.ci[p]] != block_of_row[r]) throw std::invalid_argument("factor couples two blocks");
        const int lc = local[U.ci[p]];
        if (!((x < tx - 1 && lc == li + 1) || (y < ty - 1 && lc == li + tx))) return false;
      }
    }
    return true;
  };
  bool lattice = allow_lattice && n > 0;
  for (int b = 0; b < nblocks && lattice; ++b) lattice = lattice_dims(bptr[b], bptr[b + 1] - bptr[b], ltx[b], lty[b]);
  out.lattice = lattice;
  // emits the chunks of one factor of one lattice block: whole anti-diagonals, values only
  auto emit_lattice = [&](bool lower, int base, int tx, int ty) {
    const int D = tx + ty - 1;  // number of anti-diagonals = steps
    auto coef = [&](int li, int which) -> T {  // which 0: west/east neighbour, 1: south/north neighbour
      const int r = out.perm[base + li];
      const int want = lower ? (which == 0 ? li - 1 : li - tx) : (which == 0 ? li + 1 : li + tx);
      const int x = li % tx;
      if (which == 0 && (lower ? x == 0 : x == tx - 1)) return (T)0;  // no west / east neighbour inside the tile
      if (want < 0 || want >= tx * ty) return (T)0;
      const Csr<T>& M = lower ? L : U;
      const int lb = lower ? M.rp[r] : M.rp[r] + 1, ub = lower ? M.rp[r + 1] - 1 : M.rp[r + 1];
      for (int p = lb; p < ub; ++p)
        if (local[M.ci[p]] == want) return (T)(lower ? (double)M.v[p] * di[r] * d[M.ci[p]] : (double)M.v[p] * di[r]);
      return (T)0;
    };
    int e = 0;
    while (e < D) {
      int nlev = 0, rows = 0;
      const int e0 = e;
      while (e < D) {
        const int dd = lower ? e : D - 1 - e;
        const int sz = std::min(dd, tx - 1) - std::max(0, dd - (ty - 1)) + 1;
        if (nlev > 0 && 16 + (rows + sz) * 2 * (int)sizeof(T) > chunk_target) break;
        rows += sz; ++nlev; ++e;
      }
      const int bytes = (16 + rows * 2 * (int)sizeof(T) + 15) / 16 * 16;
      const size_t off = out.stream.size();
      out.stream.resize(off + bytes, 0);
      unsigned char* rec = out.stream.data() + off;
      uint16_t* hdr = reinterpret_cast<uint16_t*>(rec);
      hdr[0] = (uint16_t)nlev; hdr[1] = (uint16_t)e0; hdr[2] = lower ? 1 : 0; hdr[3] = 0;
      T* pr = reinterpret_cast<T*>(rec + 16);
      int t = 0;
      for (int q = e0; q < e0 + nlev; ++q) {
        const int dd = lower ? q : D - 1 - q;
        for (int x = std::max(0, dd - (ty - 1)); x <= std::min(dd, tx - 1); ++x, ++t) {
          const int li = x + tx * (dd - x);
          pr[2 * t] = coef(li, 0);
          pr[2 * t + 1] = coef(li, 1);
        }
      }
      out.chunk_off.push_back((uint32_t)(out.stream.size() / 16));
      out.max_chunk_bytes = std::max(out.max_chunk_bytes, bytes);
      out.total_steps += nlev;
      out.total_slots += rows;
    }
  };
  for (int b = 0; b < nblocks; ++b) {
    const int base = bptr[b], nb = bptr[b + 1] - base;
    BlockDesc& dsc = out.blk[b];
    dsc.row0 = base; dsc.nrows = nb;
    dsc.chunk0 = (int)out.chunk_off.size() - 1;
    if (lattice) {
      dsc.pad0 = ltx[b]; dsc.pad1 = lty[b];
      dsc.levelsL = dsc.levelsU = nb ? ltx[b] + lty[b] - 1 : 0;
      if (nb) { emit_lattice(true, base, ltx[b], lty[b]); emit_lattice(false, base, ltx[b], lty[b]); }
      dsc.nchunks = (int)out.chunk_off.size() - 1 - dsc.chunk0;
      out.max_levels_L = std::max(out.max_levels_L, dsc.levelsL);
      out.max_levels_U = std::max(out.max_levels_U, dsc.levelsU);
      out.max_chunks = std::max(out.max_chunks, dsc.nchunks);
      continue;
    }
    emit_factor(L, true, base, nb, &dsc.levelsL);
    emit_factor(U, false, base, nb, &dsc.levelsU);
    dsc.nchunks = (int)out.chunk_off.size() - 1 - dsc.chunk0;
    out.max_levels_L = std::max(out.max_levels_L, dsc.levelsL);
    out.max_levels_U = std::max(out.max_levels_U, dsc.levelsU);
    out.max_chunks = std::max(out.max_chunks, dsc.nchunks);
    if (out.stream.size() / 16 > 0xFFFFFFF0ull) throw std::invalid_argument("factor stream exceeds 64 GB");
  }
  out.stream.resize(out.stream.size() + 16, 0);  // a readable tail for 16-byte granularity
}
template void pack_fused_block_stream<float>(bool, const Csr<float>&, const Csr<float>&, int, const int*, int, FusedBlockStream<float>&, bool);
template void pack_fused_block_stream<double>(bool, const Csr<double>&, const Csr<double>&, int, const int*, int, FusedBlockStream<double>&, bool);

// ------------------------------------------------------------------ multi-GPU row partition
template <class T>
void build_dist_plan(int n, const int* rp, const int* ci, const T* v, int nranks, int rank, const int* part,
                     DistPlan<T>& out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("bad rank");
  out.nranks = nranks; out.rank = rank; out.n_global = n;
  out.l2g.clear();
  std::vector<int> g2l(n, -1);
  for (int i = 0; i < n; ++i) {
    if (part[i] < 0 || part[i] >= nranks) throw std::invalid_argument("part id out of range");
    if (part[i] == rank) { g2l[i] = (int)out.l2g.size(); out.l2g.push_back(i); }
  }
  out.n_local = (int)out.l2g.size();
  // one pass over the global pattern: what I need from others (halo) and what others need from me
  std::vector<std::vector<int>> need(nranks), give(nranks);
  for (int i = 0; i < n; ++i) {
    const int pi = part[i];
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int c = ci[p], pc = part[c];
      if (pc == pi) continue;
      if (pi == rank) need[pc].push_back(c);
      else if (pc == rank) give[pi].push_back(c);
    }
  }
  auto uniq = [](std::vector<int>& a) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); };
  out.nbr.clear(); out.recv_ptr.assign(1, 0); out.send_ptr.assign(1, 0); out.send_idx.clear(); out.h2g.clear();
  std::vector<int> halo_slot(n, -1);
  for (int q = 0; q < nranks; ++q) {
    uniq(need[q]); uniq(give[q]);
    if (need[q].empty() && give[q].empty()) continue;
    out.nbr.push_back(q);
    for (int g : need[q]) { halo_slot[g] = (int)out.h2g.size(); out.h2g.push_back(g); }
    out.recv_ptr.push_back((int)out.h2g.size());
    for (int g : give[q]) out.send_idx.push_back(g2l[g]);
    out.send_ptr.push_back((int)out.send_idx.size());
  }
  out.n_halo = (int)out.h2g.size();
  Csr<T>& A = out.A;
  A.n = out.n_local;
  A.rp.assign(out.n_local + 1, 0);
  A.ci.clear(); A.v.clear();
  for (int li = 0; li < out.n_local; ++li) {
    const int i = out.l2g[li];
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int c = ci[p];
      A.ci.push_back(part[c] == rank ? g2l[c] : out.n_local + halo_slot[c]);
      A.v.push_back(v[p]);
    }
    A.rp[li + 1] = (int)A.ci.size();
  }
}
template void build_dist_plan<float>(int, const int*, const int*, const float*, int, int, const int*, DistPlan<float>&);
template void build_dist_plan<double>(int, const int*, const int*, const double*, int, int, const int*, DistPlan<double>&);

// ------------------------------------------------------------------ SpMV inspector
SpmvPlan inspect_rows(int n, const int* rp) {
  SpmvPlan p;
  int64_t padded = 0;
  for (int s = 0; s < n; s += p.slice) {
    int mx = 0;
    for (int i = s; i < std::min(n, s + p.slice); ++i) mx = std::max(mx, rp[i + 1] - rp[i]);
    padded += (int64_t)mx * p.slice;
  }
  for (int i = 0; i < n; ++i) {
    const int len = rp[i + 1] - rp[i];
    p.max_row = std::max(p.max_row, len);
    int b = len == 0 ? 0 : len <= 2 ? 1 : len <= 4 ? 2 : len <= 8 ? 3 : len <= 16 ? 4 : len <= 32 ? 5 : len <= 64 ? 6 : 7;
    p.hist[b]++;
  }
  const int64_t nnz = n ? rp[n] : 0;
  p.mean_row = n ? (double)nnz / n : 0.0;
  p.sell_padded = padded;
  // Short, regular rows (stencils): thread-per-row over sliced ELL wastes no lanes and reads
  // values/columns fully coalesced.  Long or ragged rows: CSR with a sub-warp vector per row.
  const bool regular = nnz > 0 && (double)padded <= 1.25 * (double)nnz && p.max_row <= 32;
  p.format = regular ? 2 /*SELL*/ : (p.mean_row >= 24 ? 3 /*CSR warp*/ : 1 /*CSR vector*/);
  return p;
}

void sell_slice_ptr(int n, const int* rp, int slice, int min_width, std::vector<int64_t>& slice_ptr) {
  const int ns = (n + slice - 1) / slice;
  slice_ptr.assign((size_t)ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    int mx = min_width;
    for (int i = s * slice; i < std::min(n, (s + 1) * slice); ++i) mx = std::max(mx, rp[i + 1] - rp[i]);
    slice_ptr[s + 1] = slice_ptr[s] + (int64_t)mx * slice;
  }
}

template <class T>
void build_sell(int n, const int* rp, const int* ci, const T* v, int slice, std::vector<int64_t>& slice_ptr,
                std::vector<int>& cols, std::vector<T>& vals, int min_width) {
  const int ns = (n + slice - 1) / slice;
  sell_slice_ptr(n, rp, slice, min_width, slice_ptr);
  cols.resize(slice_ptr[ns]);
  vals.resize(slice_ptr[ns]);
  // slices are disjoint ranges of the output: filled by several threads at 10 M-row sizes (parallel_ranges)
  parallel_ranges((size_t)ns, [&](size_t s0, size_t s1) {
    for (size_t s = s0; s < s1; ++s) {
      const int64_t base = slice_ptr[s], wid = (slice_ptr[s + 1] - base) / slice;
      for (int t = 0; t < slice; ++t) {
        const int64_t i = (int64_t)s * slice + t;
        int64_t k = 0;
        if (i < n)
          for (int p = rp[i]; p < rp[i + 1]; ++p, ++k) {
            cols[base + k * slice + t] = ci[p];
            vals[base + k * slice + t] = v[p];
          }
        // padding keeps column = own row (a valid, cached address; row 0 past the end) and value 0
        for (; k < wid; ++k) {
          cols[base + k * slice + t] = i < n ? (int)i : 0;
          vals[base + k * slice + t] = (T)0;
        }
      }
    }
  }, ((size_t)1 << 20) / (size_t)slice);
}

template void build_sell<float>(int, const int*, const int*, const float*, int, std::vector<int64_t>&, std::vector<int>&, std::vector<float>&, int);
template void build_sell<double>(int, const int*, const int*, const double*, int, std::vector<int64_t>&, std::vector<int>&, std::vector<double>&, int);

}  // namespace bg
