// peer_sync.cuh -- system-scope flag primitives of the peer exchange (peer.cu, optimizer.cu)
#pragma once
#include "kernels.h"

namespace drl {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// spin until every peer's flag of `phase` has reached epoch e (flags live in LOCAL memory, written by the peers)
__device__ __forceinline__ void wait_peers(const uint32_t* my_flags, int phase, int world, uint32_t e, uint32_t* err) {
  if (threadIdx.x < world) {
    const uint32_t* f = my_flags + phase * kMaxPeers + threadIdx.x;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {      // 20 s: a peer died; do not hang the GPU, report through the error word
        *err = 1u + threadIdx.x;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}


}  // namespace drl
