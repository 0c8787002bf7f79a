// oracle/shim/helper_cuda.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the CUDA-samples header the
// reference includes (src/gmres.cu:5): only checkCudaErrors is used there.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime_api.h>
#define checkCudaErrors(call)                                                                        \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) { std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); std::exit(-1); } \
  } while (0)
