// lat_probe.cu -- development probe: dependent-issue latency of the instructions on the lattice wavefront chain
// (DFMA, SHFL of a double, FSEL) for a lone warp on an SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x * 1e-3, y = 1.0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) { y = __shfl_up_sync(0xffffffffu, y, 1); y = __shfl_up_sync(0xffffffffu, y, 1); y = __shfl_up_sync(0xffffffffu, y, 1); y = __shfl_up_sync(0xffffffffu, y, 1); }
  long long t2 = clock64();
  double z = x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { double zn = __shfl_up_sync(0xffffffffu, z, 1); zn = threadIdx.x == 0 ? b : zn; double t = fma(-a, z, b); z = fma(-a, zn, t); }
  }
  long long t3 = clock64();
  float f = threadIdx.x;
  for (int i = 0; i < iters; ++i) { f = fmaf(f, (float)a, (float)b); f = fmaf(f, (float)a, (float)b); f = fmaf(f, (float)a, (float)b); f = fmaf(f, (float)a, (float)b); }
  long long t4 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
  out[threadIdx.x] = x + y + z + f;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, iters, 0.999, 0.001);
  cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("dependent DFMA: %.1f cycles; dependent SHFL(double): %.1f cycles; wavefront step (SHFL+SEL+2 DFMA): %.1f cycles; dependent FFMA: %.1f cycles\n",
         h[0] / (4.0 * iters), h[1] / (4.0 * iters), h[2] / (4.0 * iters), h[3] / (4.0 * iters));
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
  return 0;
}
