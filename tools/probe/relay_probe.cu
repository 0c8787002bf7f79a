// relay_probe.cu -- development probe: per-hop latency of a flag-in-data relay through many warps on different
// SMs with "cold" packet addresses (stride bytes apart), the access pattern of trsv_syncfree on a chain.
// hop h is done by warp h % W: it polls packet h-1 (32 lanes x `fan` packets each), then publishes packet h.
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(16) Pk { unsigned lo, f0, hi, f1; };
__device__ __forceinline__ Pk ldv(const Pk* p) {
  Pk v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stv(Pk* p, unsigned s) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s), "r"(s), "r"(s), "r"(s) : "memory");
}
// layout: hop h owns 32 packets (one per lane) at base + h*hop_stride + lane*lane_stride (bytes)
__global__ void relay(char* base, long hop_stride, long lane_stride, int hops, int W, unsigned seq, int prepoll, long long* cyc) {
  const int w = blockIdx.x, lane = threadIdx.x;
  long long t0 = clock64();
  for (int h = w; h < hops; h += W) {
    if (h > 0) {
      const Pk* p = (const Pk*)(base + (long)(h - 1) * hop_stride + (long)lane * lane_stride);
      // neighbour lane's packet too (two dependencies per row, like the 5-point factor)
      const Pk* p2 = (const Pk*)(base + (long)(h - 1) * hop_stride + (long)((lane + 1) & 31) * lane_stride);
      Pk v; do { v = ldv(p); } while (v.f0 != seq || v.f1 != seq);
      do { v = ldv(p2); } while (v.f0 != seq || v.f1 != seq);
    }
    stv((Pk*)(base + (long)h * hop_stride + (long)lane * lane_stride), seq);
  }
  if (w == (hops - 1) % W && lane == 0) *cyc = clock64() - t0;
}
int main() {
  const int hops = 20000;
  long long* cyc; cudaMalloc(&cyc, 8);
  for (long lane_stride : {16L, 20848L}) for (long hop_stride : {512L, 16L * 1304}) for (int W : {2, 21, 148, 592}) {
    const size_t bytes = (size_t)hops * hop_stride + 32 * lane_stride + 4096;
    char* base; if (cudaMalloc(&base, bytes) != cudaSuccess) { printf("alloc %zu failed\n", bytes); continue; }
    cudaMemset(base, 0, bytes);
    for (unsigned seq = 1; seq <= 2; ++seq) {
      relay<<<W, 32>>>(base, hop_stride, lane_stride, hops, W, seq, 0, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("lane_stride %6ld hop_stride %6ld W=%4d pass %u : %8.1f cycles per hop (last warp, incl. its start skew) [%s]\n", lane_stride, hop_stride, W, seq, (double)h / hops, cudaGetErrorString(e));
    }
    cudaFree(base);
  }
  return 0;
}
