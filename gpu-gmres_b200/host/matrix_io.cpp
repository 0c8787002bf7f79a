// matrix_io.cpp -- matrix / vector files in the reference's own formats, so that the solver
// can be fed the files the reference's drivers read and write.  Pure host C++ behind bg_host_* entry points.
//
//  * BG_FILE_TRIPLETS: the MatrixMarket-like text the thermal drivers read (readSparseMatrix,
//    src_thermal/SpMV_gen.cpp:93-187): '%' comment lines, then "numRows numCols nnz", then nnz lines
//    "row col val", 1-based; entries are sorted by (row, col) into CSR (genCSRFormat, :189-201).
//    Values are parsed as float like the reference does.  Two deliberate differences: indices are parsed as
//    integers-in-double (the reference goes through float and corrupts indices above 2^24), and rows without
//    entries get a proper empty range (the reference leaves their rowPtrs uninitialised, :157-165).
//  * BG_FILE_CSR_F32 / _F64: the binary dump of writeCSR / writeCSRmySpMatrix[Double]
//    (src/formatConvert.cpp:425-485): int m, n, nnz; int rowPtr[m+1]; int colIdx[nnz]; float|double val[nnz].
//  * vectors: one value per line (readInputVector, src_thermal/SpMV_gen.cpp:36-53; writeOutputVector "%20.18e").
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/b200gmres.h"

namespace bg { void set_error(const char* fmt, ...); }

struct bg_host_csr_s {
  int n_rows = 0, n_cols = 0, dtype = BG_F32, empty_rows = 0;
  std::vector<int> rp, ci;
  std::vector<float> vf;
  std::vector<double> vd;
};

namespace {
struct FileCloser { FILE* f; ~FileCloser() { if (f) std::fclose(f); } };

int read_triplets(FILE* f, bg_host_csr_s& M) {
  char line[512];
  bool have = false;
  while (std::fgets(line, sizeof line, f)) if (line[0] != '%') { have = true; break; }
  double dr, dc, dn;
  if (!have || std::sscanf(line, "%lf %lf %lf", &dr, &dc, &dn) != 3 || dr < 0 || dc < 0 || dn < 0 || dr > 2147483646.0 || dn > 2147483646.0) {
    bg::set_error("triplet file: bad size line");
    return BG_ERR_ARG;
  }
  const int nr = (int)dr, nc = (int)dc;
  const size_t nnz = (size_t)dn;
  std::vector<int> r(nnz), c(nnz);
  std::vector<float> v(nnz);
  for (size_t k = 0; k < nnz; ++k) {
    double fr, fc;
    char vbuf[128];
    if (std::fscanf(f, "%lf %lf %127s", &fr, &fc, vbuf) != 3) { bg::set_error("triplet file: entry %zu of %zu is malformed", k + 1, nnz); return BG_ERR_ARG; }
    r[k] = (int)fr - 1;  // 1-based on file (src_thermal/SpMV_gen.cpp:140-143)
    c[k] = (int)fc - 1;
    v[k] = std::strtof(vbuf, nullptr);
    if (r[k] < 0 || r[k] >= nr || c[k] < 0 || c[k] >= nc) { bg::set_error("triplet file: entry %zu has index (%d,%d) outside %dx%d", k + 1, r[k] + 1, c[k] + 1, nr, nc); return BG_ERR_ARG; }
  }
  std::vector<size_t> order(nnz);
  std::iota(order.begin(), order.end(), (size_t)0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return r[a] != r[b] ? r[a] < r[b] : c[a] < c[b]; });  // cmpRow
  M.n_rows = nr; M.n_cols = nc; M.dtype = BG_F32;
  M.rp.assign((size_t)nr + 1, 0);
  M.ci.resize(nnz); M.vf.resize(nnz);
  for (size_t k = 0; k < nnz; ++k) { M.rp[r[k] + 1]++; }
  for (int i = 0; i < nr; ++i) { if (M.rp[i + 1] == 0) M.empty_rows++; M.rp[i + 1] += M.rp[i]; }
  for (size_t k = 0; k < nnz; ++k) { M.ci[k] = c[order[k]]; M.vf[k] = v[order[k]]; }
  return BG_OK;
}

int read_csr_binary(FILE* f, bool dbl, bg_host_csr_s& M) {
  int hdr[3];
  if (std::fread(hdr, sizeof(int), 3, f) != 3 || hdr[0] < 0 || hdr[1] < 0 || hdr[2] < 0) { bg::set_error("binary CSR: bad header"); return BG_ERR_ARG; }
  M.n_rows = hdr[0]; M.n_cols = hdr[1]; M.dtype = dbl ? BG_F64 : BG_F32;
  const size_t nnz = (size_t)hdr[2];
  M.rp.resize((size_t)hdr[0] + 1); M.ci.resize(nnz);
  bool ok = std::fread(M.rp.data(), sizeof(int), M.rp.size(), f) == M.rp.size() && std::fread(M.ci.data(), sizeof(int), nnz, f) == nnz;
  if (dbl) { M.vd.resize(nnz); ok = ok && std::fread(M.vd.data(), sizeof(double), nnz, f) == nnz; }
  else { M.vf.resize(nnz); ok = ok && std::fread(M.vf.data(), sizeof(float), nnz, f) == nnz; }
  if (!ok) { bg::set_error("binary CSR: file is shorter than its header says"); return BG_ERR_ARG; }
  if (M.rp[0] != 0 || (size_t)M.rp.back() != nnz) { bg::set_error("binary CSR: rowPtr[0] != 0 or rowPtr[m] != nnz"); return BG_ERR_ARG; }
  for (int i = 0; i < M.n_rows; ++i) {
    if (M.rp[i + 1] < M.rp[i]) { bg::set_error("binary CSR: rowPtr decreases at row %d", i); return BG_ERR_ARG; }
    if (M.rp[i + 1] == M.rp[i]) M.empty_rows++;
  }
  for (size_t k = 0; k < nnz; ++k)
    if (M.ci[k] < 0 || M.ci[k] >= M.n_cols) { bg::set_error("binary CSR: column index out of range at entry %zu", k); return BG_ERR_ARG; }
  return BG_OK;
}
}  // namespace

extern "C" int bg_host_read_matrix(const char* path, int file_format, bg_host_csr_t* out) {
  if (!path || !out || file_format < BG_FILE_TRIPLETS || file_format > BG_FILE_CSR_F64) return BG_ERR_ARG;
  *out = nullptr;
  FileCloser fc{std::fopen(path, file_format == BG_FILE_TRIPLETS ? "r" : "rb")};
  if (!fc.f) { bg::set_error("cannot open file: %s", path); return BG_ERR_ARG; }
  bg_host_csr_t M = new bg_host_csr_s;
  const int rc = file_format == BG_FILE_TRIPLETS ? read_triplets(fc.f, *M) : read_csr_binary(fc.f, file_format == BG_FILE_CSR_F64, *M);
  if (rc < 0) { delete M; return rc; }
  *out = M;
  return BG_OK;
}
extern "C" int bg_host_csr_get_info(bg_host_csr_t M, bg_host_csr_info* o) {
  if (!M || !o) return BG_ERR_ARG;
  o->n_rows = M->n_rows; o->n_cols = M->n_cols; o->nnz = (int64_t)M->ci.size(); o->dtype = M->dtype; o->empty_rows = M->empty_rows;
  return BG_OK;
}
extern "C" int bg_host_csr_get(bg_host_csr_t M, int* rowptr, int* col, void* val) {
  if (!M) return BG_ERR_ARG;
  if (rowptr) std::copy(M->rp.begin(), M->rp.end(), rowptr);
  if (col) std::copy(M->ci.begin(), M->ci.end(), col);
  if (val) {
    if (M->dtype == BG_F64) std::memcpy(val, M->vd.data(), sizeof(double) * M->vd.size());
    else std::memcpy(val, M->vf.data(), sizeof(float) * M->vf.size());
  }
  return BG_OK;
}
extern "C" int bg_host_csr_destroy(bg_host_csr_t M) { delete M; return BG_OK; }

extern "C" int bg_host_write_csr(const char* path, int file_format, int n_rows, int n_cols, const int* rowptr, const int* col,
                                 const void* val, int val_dtype) {
  if (!path || !rowptr || n_rows < 0 || n_cols < 0 || (file_format != BG_FILE_CSR_F32 && file_format != BG_FILE_CSR_F64)) return BG_ERR_ARG;
  const int nnz = rowptr[n_rows];
  if (nnz > 0 && (!col || !val)) return BG_ERR_ARG;
  FileCloser fc{std::fopen(path, "wb")};
  if (!fc.f) { bg::set_error("cannot open file: %s", path); return BG_ERR_ARG; }
  bool ok = std::fwrite(&n_rows, sizeof(int), 1, fc.f) == 1 && std::fwrite(&n_cols, sizeof(int), 1, fc.f) == 1 && std::fwrite(&nnz, sizeof(int), 1, fc.f) == 1 &&
            std::fwrite(rowptr, sizeof(int), (size_t)n_rows + 1, fc.f) == (size_t)n_rows + 1 && std::fwrite(col, sizeof(int), (size_t)nnz, fc.f) == (size_t)nnz;
  const bool dbl = file_format == BG_FILE_CSR_F64;
  if (dbl == (val_dtype == BG_F64)) {
    ok = ok && std::fwrite(val, dbl ? sizeof(double) : sizeof(float), (size_t)nnz, fc.f) == (size_t)nnz;
  } else if (dbl) {
    std::vector<double> t((size_t)nnz);
    for (int k = 0; k < nnz; ++k) t[k] = (double)((const float*)val)[k];
    ok = ok && std::fwrite(t.data(), sizeof(double), (size_t)nnz, fc.f) == (size_t)nnz;
  } else {
    std::vector<float> t((size_t)nnz);
    for (int k = 0; k < nnz; ++k) t[k] = (float)((const double*)val)[k];  // the cast LDcsc2csrMySpMatrix applies (formatConvert.cpp:300-334)
    ok = ok && std::fwrite(t.data(), sizeof(float), (size_t)nnz, fc.f) == (size_t)nnz;
  }
  if (!ok) { bg::set_error("short write to %s", path); return BG_ERR_ARG; }
  return BG_OK;
}

extern "C" int bg_host_read_vector(const char* path, int n, int dtype, void* out) {
  if (!path || n < 0 || (n && !out)) return BG_ERR_ARG;
  FileCloser fc{std::fopen(path, "r")};
  if (!fc.f) { bg::set_error("cannot open file: %s", path); return BG_ERR_ARG; }
  char line[256];
  int i = 0;
  while (i < n && std::fgets(line, sizeof line, fc.f)) {  // one value per line (readInputVector)
    if (dtype == BG_F64) ((double*)out)[i] = std::strtod(line, nullptr);
    else ((float*)out)[i] = std::strtof(line, nullptr);
    ++i;
  }
  if (i < n) { bg::set_error("vector file %s holds %d of %d values", path, i, n); return BG_ERR_ARG; }
  return BG_OK;
}
extern "C" int bg_host_write_vector(const char* path, int n, int dtype, const void* x) {
  if (!path || n < 0 || (n && !x)) return BG_ERR_ARG;
  FileCloser fc{std::fopen(path, "w")};
  if (!fc.f) { bg::set_error("cannot open file: %s", path); return BG_ERR_ARG; }
  // writeOutputVector(float*, const char*, int), src_thermal/SpMV_gen.cpp:55-69
  for (int i = 0; i < n; ++i) std::fprintf(fc.f, "%20.18e\n", dtype == BG_F64 ? ((const double*)x)[i] : (double)((const float*)x)[i]);
  return BG_OK;
}
