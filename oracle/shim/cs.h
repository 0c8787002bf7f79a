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

// the int flavour and the two allocators src/matrix.cpp (mat2cs / mat2csdl) uses; defined in oracle/ref_parser_wrap.cpp
typedef struct cs_sparse {
  int nzmax, m, n;
  int *p, *i;
  double* x;
  int nz;
} cs;
cs* cs_spalloc(int m, int n, int nzmax, int values, int triplet);
cs_dl* cs_dl_spalloc(UF_long m, UF_long n, UF_long nzmax, UF_long values, UF_long triplet);
