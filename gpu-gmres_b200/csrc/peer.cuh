// peer.cuh -- device helpers of the peer-memory (NVLink) transport, shared by orth.cu and spmv.cu.
#pragma once
#include "internal.h"

namespace bg {

// Flag-in-data messages over NVLink (internal.h LL16): one 16-byte store carries a double and its
// sequence number twice; a reader that sees both sequence words equal to the expected one has the value.
__device__ __forceinline__ void ll_store(LL16* p, double v, unsigned seq) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(seq),
               "r"((unsigned)__double2hiint(v)), "r"(seq)
               : "memory");
}
// The same packet between the SMs of ONE GPU (triangular-solve hand-offs, the fused MGS step's sums): GPU scope is
// enough and cheaper than the system scope a volatile access implies (which also covers NVLink peers and the host).
__device__ __forceinline__ void ll_store_gpu(LL16* p, double v, unsigned seq) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(seq),
               "r"((unsigned)__double2hiint(v)), "r"(seq)
               : "memory");
}
__device__ __forceinline__ LL16 ll_load_gpu(const LL16* p) {
  LL16 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.lo), "=r"(v.f0), "=r"(v.hi), "=r"(v.f1) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// poll until the element carries `seq`; a peer that never arrives (crashed rank) must trap, not hang the GPU.
// timeout_ns comes from the owning descriptor (B200GMRES_PEER_TIMEOUT_S, default 120 s: ranks may enter a solve
// with host-side skew, e.g. unequal setup work before the first collective call).
__device__ __forceinline__ double ll_wait(const LL16* p, unsigned seq, unsigned long long timeout_ns) {
  unsigned lo, f0, hi, f1;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
    if (f0 == seq && f1 == seq) break;
    if ((spin & 1023u) == 1023u) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout_ns) __trap();
    }
  }
  return __hiloint2double((int)hi, (int)lo);
}

}  // namespace bg
