// handoff_probe.cu -- development probe: latency of a flag-in-data hand-off between two warps on different SMs
// through L2, for the load/store flavours the sync-free SpTRSV could use.  nvcc -arch=sm_100a -O3 -o handoff_probe
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(16) Pk { unsigned lo, f0, hi, f1; };
template <int MODE> __device__ __forceinline__ Pk ld(const Pk* p) {
  Pk v;
  if (MODE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  if (MODE == 3) asm volatile("ld.acquire.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  return v;
}
template <int MODE> __device__ __forceinline__ void st(Pk* p, unsigned s) {
  if (MODE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
  if (MODE == 1) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
  if (MODE == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
  if (MODE == 3) asm volatile("st.release.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
}
// CTA 0 <-> CTA 1 ping-pong; CTAs >= 2 poll the same word (contention); `lanes` lanes take part on each side
template <int MODE> __global__ void pingpong(Pk* a, Pk* b, int iters, int lanes, long long* cycles, volatile int* stop) {
  const int lane = threadIdx.x;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      if (lane < lanes) { st<MODE>(a + lane, i); Pk v; do { v = ld<MODE>(b + lane); } while (v.f0 != (unsigned)i || v.f1 != (unsigned)i); }
      __syncwarp();
    }
    if (lane == 0) { *cycles = clock64() - t0; *stop = 1; __threadfence(); }
  } else if (blockIdx.x == 1) {
    for (int i = 1; i <= iters; ++i) {
      if (lane < lanes) { Pk v; do { v = ld<MODE>(a + lane); } while (v.f0 != (unsigned)i || v.f1 != (unsigned)i); st<MODE>(b + lane, i); }
      __syncwarp();
    }
  } else {
    while (!*stop) { Pk v = ld<MODE>(a + (lane % lanes)); if (v.f0 == 0xffffffffu) break; }
  }
}
template <int MODE> void run(const char* name, int ctas, int lanes) {
  Pk *a, *b; long long* cyc; int* stop;
  cudaMalloc(&a, 4096); cudaMalloc(&b, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&stop, 4);
  cudaMemset(a, 0, 4096); cudaMemset(b, 0, 4096); cudaMemset(stop, 0, 4);
  const int iters = 20000;
  pingpong<MODE><<<ctas, 32>>>(a, b, iters, lanes, cyc, stop);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s ctas=%4d lanes=%2d : %8.1f cycles per round trip (2 hand-offs)  [%s]\n", name, ctas, lanes, (double)h / iters, cudaGetErrorString(e));
  cudaFree(a); cudaFree(b); cudaFree(cyc); cudaFree(stop);
}
int main() {
  for (int ctas : {2, 150, 600}) for (int lanes : {1, 32}) {
    run<0>("volatile (strong.sys)", ctas, lanes);
    run<1>("relaxed.gpu", ctas, lanes);
    run<2>("ld.cg / st.cg", ctas, lanes);
    run<3>("acquire/release.gpu", ctas, lanes);
  }
  return 0;
}
