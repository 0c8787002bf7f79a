// SpMV.h (B200 shim) -- source-compatible stand-in for the reference's src/SpMV.h:22-141 (and src_thermal/SpMV.h).
// Same struct names, members and function names; the GPU-side functions forward to the C ABI in
// include/b200gmres.h.  Host-only helpers the reference keeps in its own translation units
// (readSparseMatrix, readInputVector, writeOutputVector, computeSpMV, LUSolve ...: SpMV_gen.cpp, SpMV_compute.cpp,
// which stay in the build) are only declared here so unchanged drivers keep compiling against their own definitions.
#ifndef __SPMV_H__
#define __SPMV_H__
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fstream>
using namespace std;  // the reference's header does (src/SpMV.h:15); its drivers rely on it

struct nzInfo { int rowNum; int colNum; float val; };
typedef struct nzInfo NZEntry;
struct SpM { int numRows; int numCols; int numNZEntries; NZEntry* nzentries; int* rowPtrs; int* colPtrs; };
typedef struct SpM SpMatrix;
struct SpMGPU {
  float* d_val; int* d_indices; int* d_rowIndices;
  int* d_ins_indices; int* d_ins_rowIndices; int* d_ins_inputList;
};
typedef struct SpMGPU SpMatrixGPU;

class MySpMatrix {
 public:
  int isCSR, numRows, numCols, numNZEntries;
  float* d_val; int* d_indices; int* d_rowIndices;
  float* val; int* indices; int* rowIndices;
  void Initilize(SpMatrix& M);  // src/mySpMatrix.cu:11-39: triplets -> CSR on host + device copy
};
void mySpMatrixFree(MySpMatrix*);
class MySpMatrixDouble {
 public:
  int isCSR, numRows, numCols, numNZEntries;
  double* d_val; int* d_indices; int* d_rowIndices;
  double* val; int* indices; int* rowIndices;
};
void mySpMatrixDoubleFree(MySpMatrixDouble*);

// ---- defined by the shim (device side, src/SpMV_alloc.cu:16-211)
void allocateSparseMatrixGPU(SpMatrixGPU* spm, SpMatrix* m, float* h_val, int* h_rowIndices, int* h_indices,
                             const int numRows, const int numCols);
void gpuMallocCpyCSR(SpMatrixGPU* spm, float* h_val, int* h_rowIndices, int* h_indices, const int numRows, const int numCols);
void gpuMallocCpyCSRmySpM(SpMatrixGPU* spm, MySpMatrix* mySpM);
void freeSparseMatrixGPU(SpMatrixGPU* spm);
// x = A y on the device (SpMV_cuda / the SpMV<<<>>> launch sites, src_thermal/main1.cu:224)
void SpMV_cuda(float* d_x, SpMatrixGPU* spm, const float* d_y, const int numRows, const int numCols, const int nnz);
void setGPUdevice();  // src/setGPUdevice.cu:4-29
// releases every device object the shim cached on behalf of raw-pointer callers (matrices, workspaces)
void b200gmresShimRelease();

// ---- stay with the reference (SpMV_gen.cpp, SpMV_compute.cpp); genCSRFormat also has a weak definition in the shim
void addTwoVec2(const float* v1, float* v2, const int num);
void loadVectorFromFile(ifstream& fin, int num, float* vec);
void readInputVector(float* y, const char* filename, int numCols);
void writeOutputVector(float* x, const char* filename, int numRows);
void writeOutputVector(float* x, FILE* f, int numRows);
void readSparseMatrix(SpMatrix* m, const char* filename, int format);
void readSparseMatrixBinaryD2S(SpMatrix* m, const char* filename, int format);
void genCSRFormat(SpMatrix* m, float* val, int* rowIndices, int* indices);
void genCSCFormat(SpMatrix* m, float* val, int* colIndices, int* indices);
void genBCSRFormat(SpMatrix* m, float** val, int** rowIndices, int** indices, int* numblocks, int bsx, int bsy);
void genPaddedCSRFormat(SpMatrix* m, float** val, int** rowIndices, int** indices);
void computeSpMV(float* x, const float* val, const int* rowIndices, const int* indices, const float* y, const int numRows);
void LUSolve(float* x, const float* l_val, const int* l_rowIndices, const int* l_indices, const float* u_val,
             const int* u_rowIndices, const int* u_indices, const float* y, const int numRows);
void LUSolve_ignoreZero(float* x, const float* l_val, const int* l_rowIndices, const int* l_indices, const float* u_val,
                        const int* u_rowIndices, const int* u_indices, const float* y, const int numRows);
void computeSpMV_BCSR(float* x, const float* val, const int* rowIndices, const int* indices, const float* y,
                      const int numRows, const int numCols, const int bsx, const int bsy);
void printSparseMatrix(SpMatrix* m, int Data, int Header);
#endif
