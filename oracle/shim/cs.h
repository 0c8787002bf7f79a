// oracle/shim/cs.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for CXSparse's cs.h (SuiteSparse is not in this image) so that the
// reference's src/partition.cpp compiles unchanged into oracle/_ref.  Only the members that
// translation unit touches (m, p, i) need to line up; layout follows CXSparse's cs_dl.
#pragma once
typedef long UF_long;
typedef struct cs_dl_sparse {
  UF_long nzmax, m, n;
  UF_long *p, *i;
  double* x;
  UF_long nz;
} cs_dl;
