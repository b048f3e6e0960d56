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

// Barrier error word (PeerTable::err / OptState::wait_err): [0] lives in DEVICE memory and is what the kernels test,
// the host-visible copy (mapped pinned memory) is written once when the error is raised.  The error is STICKY and
// FATAL: once a wait timed out on this rank, every later exchange / update kernel of this rank returns without
// touching any local or remote buffer, so parameters, slots and step counters stay exactly as they were before the
// failed step and the host gets the error from drl_learner_wait (which also refuses get_params / get_opt_state).
struct PeerErr {
  uint32_t* dev;       // device word: 0 = ok, 1 + peer index = that peer never arrived
  uint32_t* host;      // mapped pinned copy for the host
  unsigned long long timeout_ns;
};

__device__ __forceinline__ bool peer_failed(const PeerErr& e) {
  return *reinterpret_cast<volatile const uint32_t*>(e.dev) != 0u;
}

// spin until every peer's flag of `phase` has reached epoch e (flags live in LOCAL memory, written by the peers);
// returns false (CTA-uniform) when this rank is in the failed state, either from before or because a peer did not
// arrive within the time-out (DRL_B200_PEER_TIMEOUT_S, default 600 s: a rank blocked on its actor queue for minutes is
// normal for IMPALA -- NCCL would block there as well -- whereas a dead peer must not hang the GPU for ever).
__device__ __forceinline__ bool wait_peers(const uint32_t* my_flags, int phase, int world, uint32_t e, const PeerErr& err) {
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = peer_failed(err) ? 0 : 1;
  __syncthreads();
  if (ok_s && threadIdx.x < world) {
    const uint32_t* f = my_flags + phase * kMaxPeers + threadIdx.x;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > err.timeout_ns || peer_failed(err)) {
        atomicCAS(err.dev, 0u, 1u + threadIdx.x);
        *reinterpret_cast<volatile uint32_t*>(err.host) = 1u + threadIdx.x;
        __threadfence_system();
        ok_s = 0;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const bool ok = ok_s != 0;
  __syncthreads();   // a kernel may call wait_peers again (one call per exchange part): everyone has read ok_s before
                     // thread 0 of the next call overwrites it (racecheck, 2 GPUs: profiles/r02_sanitizer_peer_racecheck*.log)
  return ok;
}


}  // namespace drl
