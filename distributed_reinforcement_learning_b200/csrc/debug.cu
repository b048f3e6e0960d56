// debug.cu -- drl_debug_gemm: runs ONE plain GEMM  C[M,N] = A * B  on either contraction core
// (FP32-FFMA gather-GEMM or tcgen05 3xTF32 UMMA) with every operand-major combination the layers use,
// so the shared-memory layouts / UMMA descriptors can be validated in isolation against NumPy.
#include <vector>

#include "gemm_simt.cuh"
#include "gemm_umma.cuh"
#include "kernels.h"
#include "loaders.cuh"

namespace drl {

template <class AL, class BL>
static int run_one(int core, int bn, cudaStream_t s, const AL& al, const BL& bl, float* C, int M, int N, int K,
                   int splits) {
  int kchunk = (K + splits - 1) / splits;
  kchunk = (kchunk + 31) / 32 * 32;
  const int nz = (K + kchunk - 1) / kchunk;
  EpRaw<true> ep{C, N, (size_t)(M + 1) * N, 1.0f, M, N};     // slab z: [M+1, N], row M = column sums of B
  if (core == 1) return launch_gemm_simt<CfgMid>(s, al, bl, ep, M, N, K, nz, kchunk, kchunk);
  switch (bn) {
    case 32: return launch_gemm_umma<UmmaCfg<32, 2, 2>>(s, al, bl, ep, M, N, K, nz, kchunk, kchunk);
    case 64: return launch_gemm_umma<UmmaCfg<64, 2, 2>>(s, al, bl, ep, M, N, K, nz, kchunk, kchunk);
    case 128: return launch_gemm_umma<UmmaCfg<128, 3, 1, 8>>(s, al, bl, ep, M, N, K, nz, kchunk, kchunk);
    case 256: return launch_gemm_umma<UmmaCfg<256, 2, 1, 8>>(s, al, bl, ep, M, N, K, nz, kchunk, kchunk);
    default: set_error("debug_gemm: bn must be 32, 64, 128 or 256"); return DRL_ERR_INVALID;
  }
}

}  // namespace drl

using namespace drl;

// core: 1 = FFMA, 2 = tcgen05.  a_kmajor: A given as [M,K] row-major (1) or as [K,M] row-major (0).
// b_kmajor: B given as [N,K] row-major (1) or [K,N] row-major (0).  C out: [splits][(M+1), N] host floats
// (row M of each slab = column sums of B over that split's K range when B is N-major, else untouched).
extern "C" int drl_debug_gemm(int32_t core, int32_t bn, int32_t a_kmajor, int32_t b_kmajor, int32_t M, int32_t N,
                              int32_t K, int32_t splits, const float* A, const float* B, float* C) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1 || splits < 1) { set_error("debug_gemm: bad argument"); return DRL_ERR_INVALID; }
  if (drl_device_count() < 1) { set_error("CUDA device not available (no CPU fallback)"); return DRL_ERR_CUDA; }
  pdl_break(0);
  float *dA = nullptr, *dB = nullptr, *dC = nullptr;
  const size_t nA = (size_t)M * K, nB = (size_t)N * K, nC = (size_t)splits * (M + 1) * N;
  DRL_CUDA_CHECK(cudaMalloc(&dA, nA * 4 + 64));
  DRL_CUDA_CHECK(cudaMalloc(&dB, nB * 4 + 64));
  DRL_CUDA_CHECK(cudaMalloc(&dC, nC * 4 + 64));
  DRL_CUDA_CHECK(cudaMemcpy(dA, A, nA * 4, cudaMemcpyHostToDevice));
  DRL_CUDA_CHECK(cudaMemcpy(dB, B, nB * 4, cudaMemcpyHostToDevice));
  DRL_CUDA_CHECK(cudaMemset(dC, 0, nC * 4));
  int rc;
  if (a_kmajor && b_kmajor) rc = run_one(core, bn, 0, PlainA{dA, K, 0}, PlainBT{dB, K, 0}, dC, M, N, K, splits);
  else if (a_kmajor && !b_kmajor) rc = run_one(core, bn, 0, PlainA{dA, K, 0}, PlainB{dB, N, 0}, dC, M, N, K, splits);
  else if (!a_kmajor && b_kmajor) rc = run_one(core, bn, 0, PlainAT{dA, M, 0}, PlainBT{dB, K, 0}, dC, M, N, K, splits);
  else rc = run_one(core, bn, 0, PlainAT{dA, M, 0}, PlainB{dB, N, 0}, dC, M, N, K, splits);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == DRL_OK && e != cudaSuccess) { set_error("debug_gemm kernel failed: %s", cudaGetErrorString(e)); rc = DRL_ERR_CUDA; }
  if (rc == DRL_OK) {
    e = cudaMemcpy(C, dC, nC * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("debug_gemm copy failed: %s", cudaGetErrorString(e)); rc = DRL_ERR_CUDA; }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return rc;
}
