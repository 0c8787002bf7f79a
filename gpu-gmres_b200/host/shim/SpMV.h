// SpMV.h (B200 shim) -- source-compatible stand-in for the reference's src/SpMV.h:22-141.
// Same struct names, members and function names; the GPU-side functions forward to the C ABI in
// include/b200gmres.h.  Host-only helpers the reference keeps in its own translation units
// (computeSpMV, LUSolve, readSparseMatrix, genCSRFormat ...: SpMV_compute.cpp, SpMV_gen.cpp) are
// only declared here so unchanged drivers keep compiling against their own definitions.
#ifndef __SPMV_H__
#define __SPMV_H__
#include <stdio.h>

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

// device CSR upload (src/SpMV_alloc.cu:16-162)
void allocateSparseMatrixGPU(SpMatrixGPU* spm, SpMatrix* m, float* h_val, int* h_rowIndices, int* h_indices,
                             const int numRows, const int numCols);
void gpuMallocCpyCSR(SpMatrixGPU* spm, float* h_val, int* h_rowIndices, int* h_indices, const int numRows, const int numCols);
void gpuMallocCpyCSRmySpM(SpMatrixGPU* spm, MySpMatrix* mySpM);
void freeSparseMatrixGPU(SpMatrixGPU* spm);
// x = A y on the device (SpMV_cuda / the SpMV<<<>>> launch sites, src_thermal/main1.cu:224)
void SpMV_cuda(float* d_x, SpMatrixGPU* spm, const float* d_y, const int numRows, const int numCols, const int nnz);
void genCSRFormat(SpMatrix* m, float* val, int* rowIndices, int* indices);
void setGPUdevice();  // src/setGPUdevice.cu:4-29
#endif
