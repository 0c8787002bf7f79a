// oracle/ref_io_wrap.cpp -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/libref_io.so together with the
// reference's src_thermal/SpMV_gen.cpp, compiled unchanged): the reference's own matrix / vector file readers
// (readSparseMatrix + genCSRFormat, readInputVector, writeOutputVector), used to pin bg_host_read_matrix & co.
#include <cstdlib>
#include "SpMV.h"
void readSparseMatrix(SpMatrix* m, const char* filename, int format);
void genCSRFormat(SpMatrix* m, float* val, int* rowIndices, int* indices);
void readInputVector(float* y, const char* filename, int numCols);
void writeOutputVector(float* x, const char* filename, int numRows);
extern "C" {
// two calls: sizes first (val == NULL), then the arrays
int ref_readSparseMatrix(const char* path, int* n_rows, int* n_cols, int* nnz, float* val, int* rowIndices, int* indices) {
  SpMatrix m;
  readSparseMatrix(&m, path, 0);
  *n_rows = m.numRows; *n_cols = m.numCols; *nnz = m.numNZEntries;
  if (val) genCSRFormat(&m, val, rowIndices, indices);
  free(m.nzentries); free(m.rowPtrs); free(m.colPtrs);
  return 0;
}
void ref_readInputVector(float* y, const char* path, int n) { readInputVector(y, path, n); }
void ref_writeOutputVector(float* x, const char* path, int n) { writeOutputVector(x, path, n); }
}
