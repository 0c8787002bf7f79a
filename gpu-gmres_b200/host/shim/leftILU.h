// leftILU.h (B200 shim) -- src/leftILU.h:19-44 host entry points over bg_host_ilu0 / bg_host_factor_levels.
#ifndef LEFTILU_H_
#define LEFTILU_H_
#include <vector>
// six malloc'ed arrays returned through the references, owned by the caller (SURVEY 8b ownership)
void leftILU(const int numRows, float* h_val, int* h_rowIndices, int* h_indices, float*& l_val, int*& l_rowIndices,
             int*& l_indices, float*& u_val, int*& u_rowIndices, int*& u_indices);
void generateLevel(const int numRows, const float* h_val, const int* h_rowIndices, const int* h_indices,
                   std::vector<int>& node, std::vector<int>& level);
#endif
