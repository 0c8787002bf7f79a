// oracle/shim/itpp/base/vec.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for IT++'s vec / ivec / mat / sparse_mat (IT++ is not in this image) so that the reference's netlist
// front end (src/parser.cpp, element.cpp, matrix.cpp, ...) compiles unchanged into oracle/_ref/libref_parser.so.
// Only what those translation units touch: set_size(n, keep), operator()(i), size().
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
namespace itpp {
template <class T> class Vec {
  std::vector<T> d;
 public:
  Vec() {}
  explicit Vec(int n) : d((size_t)n) {}
  void set_size(int n, bool copy = false) {
    if (copy) d.resize((size_t)n);
    else d.assign((size_t)n, T());
  }
  int size() const { return (int)d.size(); }
  int length() const { return (int)d.size(); }
  T& operator()(int i) { return d[(size_t)i]; }
  const T& operator()(int i) const { return d[(size_t)i]; }
  T* _data() { return d.data(); }
};
typedef Vec<double> vec;
typedef Vec<int> ivec;
class mat {};
class sparse_vec {
 public:
  int nnz() const { return 0; }
  int get_nz_index(int) const { return 0; }
  double get_nz_data(int) const { return 0.0; }
};
class sparse_mat {
 public:
  int nnz() const { return 0; }
  int cols() const { return 0; }
  sparse_vec get_col(int) const { return sparse_vec(); }
};
inline int floor_i(double x) { return (int)std::floor(x); }
}  // namespace itpp
