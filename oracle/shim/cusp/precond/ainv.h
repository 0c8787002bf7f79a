#pragma once
#include "../csr_matrix.h"
namespace cusp { namespace precond {
template <class V, class M> struct nonsym_bridson_ainv {
  csr_matrix<int, V, M> w_t, z;
  array1d<V, M> diagonals;
  template <class A> nonsym_bridson_ainv(const A&, double) { std::abort(); }
  template <class X, class Y> void operator()(X&, Y&) { std::abort(); }
};
} }
