// shim_kernels.cu -- device code of the reference-API shim: the SpMV kernel symbol the unchanged drivers launch
// with <<<grid, block>>> themselves (src/SpMV_kernel.h:19-23).  Hand-written for sm_100a; not a port of
// src/SpMV_kernel.cu: the reference's launch gives every row one half-warp, so row r is owned by the 16 consecutive
// threads [16 r, 16 r + 16) of the launch, which stride over the row (coalesced 64-byte segments of values and
// column indices), gather y through L1/L2 and reduce with four shuffles -- no shared memory, no serial add.
#include <cuda_runtime.h>

#include "SpMV_kernel.h"

namespace {
__device__ __forceinline__ void spmv_half_warp(float* x, const float* __restrict__ val, const int* __restrict__ rp,
                                               const int* __restrict__ ci, const float* __restrict__ y, int n) {
  const long long tpb = (long long)blockDim.x * blockDim.y * blockDim.z;
  const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const long long tid = bid * tpb + ((long long)threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const long long nthreads = tpb * gridDim.x * gridDim.y * gridDim.z;
  const int sub = (int)(tid & 15);
  // the reference shape has exactly one half-warp per row; any other shape strides over the rows
  for (long long row = tid >> 4; row < n; row += nthreads >> 4) {
    const int lb = rp[row], ub = rp[row + 1];
    float acc = 0.f;
    for (int p = lb + sub; p < ub; p += 16) acc += val[p] * y[ci[p]];
    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, 16);
    if (sub == 0) x[row] = acc;
  }
}
}  // namespace

__global__ void SpMV(float* x, const float* val, const int* rowIndices, const int* indices, const float* y, const int numRows,
                     const int, const int) {
  spmv_half_warp(x, val, rowIndices, indices, y, numRows);
}
__global__ void SpMV_s(float* x, const float* val, const int* rowIndices, const int* indices, const float* y, const int numRows,
                       const int, const int) {
  spmv_half_warp(x, val, rowIndices, indices, y, numRows);
}
