// SpMV_kernel.h (B200 shim) -- the kernel symbol the reference's drivers launch directly
// (src/SpMV_kernel.h:19-23; src_thermal/main1.cu:224, main2.cu:254,273,496,499; src/gmres.cu:1221,1230) with the
// reference's signature.  Defined in shim_kernels.cu (compiled for sm_100a into libb200gmres_shim.so); it accepts
// the reference's launch shape -- dim3 grid(1, ceil(n/32)), block(1, 512): one half-warp of 16 threads per row --
// and in fact any 1-D/2-D shape that supplies at least 16 threads per row.
#ifndef _SPMV_KERNEL_H_
#define _SPMV_KERNEL_H_
#include "config.h"
#ifdef __CUDACC__
__global__ void SpMV(float* x, const float* val, const int* rowIndices, const int* indices, const float* y, const int numRows,
                     const int numCols, const int numNonZeroElements);
__global__ void SpMV_s(float* x, const float* val, const int* rowIndices, const int* indices, const float* y, const int numRows,
                       const int numCols, const int numNonZeroElements);
#endif
#endif
