// oracle/shim/cusp/*.h -- TEST INFRASTRUCTURE ONLY.  Minimal stand-ins for the CUSP types the reference
// names in MyAINV (src/preconditioner.cu:110-201, out of scope) so that its translation units compile
// unchanged without CUSP (which does not build against CUDA 12 Thrust).  Never executed.
#pragma once
#include <cstdlib>
#include <iostream>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
#include <thrust/memory.h>
namespace cusp {
struct device_memory {};
struct host_memory {};
template <class T, class M> struct array1d {
  std::vector<T> v;
  void resize(size_t n) { v.resize(n); }
  size_t size() const { return v.size(); }
  T& operator[](size_t i) { return v[i]; }
};
template <class I, class V, class M> struct csr_matrix {
  array1d<I, M> row_offsets, column_indices;
  array1d<V, M> values;
  void resize(size_t r, size_t, size_t nnz) { row_offsets.resize(r + 1); column_indices.resize(nnz); values.resize(nnz); }
};
template <class X> void print(const X&) { std::abort(); }
template <class X, class S> void print(const X&, S&) { std::abort(); }
}  // namespace cusp
