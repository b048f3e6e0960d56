// mma_rate.cu -- issue / completion rate of tcgen05.mma (kind::f16 and kind::tf32, M = 128, K = 32 bytes) versus N, from
// ONE thread of one CTA per SM, operands in shared memory (SWIZZLE_128B K-major, contents irrelevant), one accumulator.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int N, int KIND>   // KIND 0 = f16 (bf16 operands), 1 = tf32
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int distinct, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (128 + 256) * 128 * 4 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t fmt = KIND == 0 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 128 * 2;
    // descriptors precomputed (as a production kernel would: base + (byte offset >> 4)); 8 MMAs per loop trip, unrolled
    uint64_t da[8], db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da[j] = make_desc(a0 + (j & 3) * 32 + ((j >> 2) % distinct) * 16384, 16, 1024, 2);
      db[j] = make_desc(b0 + (j & 3) * 32, 16, 1024, 2);
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da[j]), "l"(db[j]), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da[j]), "l"(db[j]), "r"(idesc), "r"(1u) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int N, int KIND>
void run(const char* name, int grid, int distinct) {
  long long* d;
  cudaMalloc(&d, 16);
  auto k = rate_kernel<N, KIND>;
  const int smem = (128 + 256) * 128 * 4 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  k<<<grid, 128, smem>>>(iters, distinct, d);
  k<<<grid, 128, smem>>>(iters, distinct, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-5s N=%3d grid=%3d distinct=%2d: issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA  (floor 128*N/256 = %d)  %s\n", name, N, grid,
         distinct, (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<32, 0>("f16", grid, 4); run<64, 0>("f16", grid, 4); run<128, 0>("f16", grid, 4); run<256, 0>("f16", grid, 4);
    run<64, 0>("f16", grid, 16);
    run<32, 1>("tf32", grid, 4); run<64, 1>("tf32", grid, 4); run<256, 1>("tf32", grid, 4);
  }
  return 0;
}
