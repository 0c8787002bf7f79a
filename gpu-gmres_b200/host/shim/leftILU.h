// leftILU.h (B200 shim) -- src/leftILU.h:19-44 host entry points over bg_host_ilu0 / bg_host_factor_levels.
#ifndef LEFTILU_H_
#define LEFTILU_H_
#include <assert.h>
#include <sys/time.h>
#include <exception>
#include <iostream>
#include <vector>
#include "SpMV.h"
#include "defs.h"
using namespace std;
// six malloc'ed arrays returned through the references, owned by the caller (SURVEY 8b ownership)
void leftILU(const int numRows, float* h_val, int* h_rowIndices, int* h_indices, float*& l_val, int*& l_rowIndices,
             int*& l_indices, float*& u_val, int*& u_rowIndices, int*& u_indices);
void generateLevel(const int numRows, const float* h_val, const int* h_rowIndices, const int* h_indices,
                   vector<int>& node, vector<int>& level);
// format helpers of the factorisation (src/leftILU.cu:371-541); three malloc'ed arrays each, owned by the caller
void csr2csc(const int numRows, const float* csr_val, const int* csr_rowIndices, const int* csr_indices, float*& csc_val,
             int*& csc_colIndices, int*& csc_indices);
void splitLU_csc(const int numRows, const float* csc_val, const int* csc_colIndices, const int* csc_indices, float*& L_val,
                 int*& L_colIndices, int*& L_indices, float*& U_val, int*& U_colIndices, int*& U_indices);
void splitLU_csr(const int numRows, const float* csr_val, const int* csr_rowIndices, const int* csr_indices, float*& L_val,
                 int*& L_rowIndices, int*& L_indices, float*& U_val, int*& U_rowIndices, int*& U_indices);
#endif
