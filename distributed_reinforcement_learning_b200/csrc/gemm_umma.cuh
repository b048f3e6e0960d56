// gemm_umma.cuh -- gather-GEMM on the 5th-generation tensor cores (tcgen05.mma, kind::tf32) with
// fp32-grade accuracy by operand splitting ("3xTF32"):
//
//     x = hi + lo,  hi = rna_tf32(x),  lo = x - hi      (lo is exact in fp32, |lo| <= 2^-11 |x|)
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi           (dropped term A_lo*B_lo ~ 2^-22 relative)
//
// all three products accumulate in fp32 in ONE TMEM accumulator.  Same loader / epilogue functor
// concepts as gemm_simt.cuh, so every layer of layers.cu can run on either core.
//
// Structure (one 128 x BN output tile per CTA; 160-320 threads; 1-3 CTAs per SM depending on the configuration):
//   warps 0..PW-1 (PW = 4 or 8): PRODUCERS, then EPILOGUE.  Each thread gathers float4 groups of A and B through the
//               loader functors (im2col, concat, transposes, uint8 -> fp32 happen here), splits them into hi/lo and
//               writes both into the stage's shared-memory operand tiles in the canonical swizzled UMMA layout
//               (K-major for kContigK loaders, MN-major otherwise), then fence.proxy.async and mbarrier-arrive on
//               full[stage].  A weight operand may instead arrive as a pre-split, pre-tiled image with one
//               cp.async.bulk per stage (PretiledB; issued by warp PW+1 when Cfg::LW).
//   warp PW   : TMEM allocation + MMA ISSUER: waits full[stage], one lane issues 3 x (BK/8)
//               tcgen05.mma (M=128, N=BN, K=8) from shared-memory descriptors, then tcgen05.commit ->
//               empty[stage]; after the last K tile tcgen05.commit -> acc_full.
//   epilogue  : the producer warps read their 32 TMEM lanes (tcgen05.ld 32x32b.x32), transpose each 32x32 block
//               through a swizzled shared tile (epilogue_store_32x32) and hand row segments to the epilogue
//               functor (bias/ReLU/mask/split-K partial ...).
// gemm_tma.cuh holds the TMA-fed variant (cp.async.bulk.tensor im2col boxes) for the conv2/conv3 forward.
//
// Canonical SWIZZLE_128B layouts (CuTe mma_traits_sm100.hpp::make_umma_desc), BK = 32 floats:
//   K-major : row r is 128 contiguous bytes (32 floats of K) at r*128; the 16-byte chunk c of row r is
//             stored at chunk position c ^ (r & 7)  (Swizzle<3,4,3>: address bits [4,7) ^= bits [7,10)).
//             8-row groups are 1024 bytes apart (SBO = 1024).  A K=8 MMA slice j starts at +32*j bytes.
//   MN-major: SWIZZLE_128B_BASE32B, atoms of 32 (mn) x 4 (k) -- see UmmaTile below.
// All operand tiles are 1024-byte aligned (base_offset = 0).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace drl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns (one 32-bit word each) -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One lane of a converged warp, chosen by the hardware (elect.sync).  Unlike `lane == 0`, the compiler knows the branch
// holds exactly one thread and keeps the tcgen05 operands (descriptors, TMEM address) in uniform registers with a plain
// R2UR; with `lane == 0` it wrapped EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (cuobjdump), which
// together with the per-MMA descriptor arithmetic cost ~120-165 cycles per MMA against 44-128 for the instruction itself
// (tools/microbench/mma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// descriptor = constant part (everything but the start address) + (shared-memory byte address >> 4)
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t saddr) { return base + (uint64_t)(saddr >> 4); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): descriptor version 1 (sm_100), base_offset 0,
// layout_type 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, dense, no negate; M=128
__host__ __device__ constexpr uint32_t make_idesc(int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                        // c_format  = F32
         | (2u << 7)                      // a_format  = TF32
         | (2u << 10)                     // b_format  = TF32
         | ((a_mn_major ? 1u : 0u) << 15) // a_major
         | ((b_mn_major ? 1u : 0u) << 16) // b_major
         | ((uint32_t)(N >> 3) << 17)     // n_dim
         | ((uint32_t)(128 >> 4) << 24);  // m_dim
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void split4(const float4& x, float4& h, float4& l) {
  h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
  l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
}

}  // namespace umma

// One operand tile of a stage: ROWS (128 for A, BN for B) x 32 floats.
//   K-major : SWIZZLE_128B (layout type 2), rows of 128 bytes, 16-byte chunks XOR (row & 7).
//   MN-major: SWIZZLE_128B_BASE32B (layout type 1) -- the ONLY shared-memory layout the hardware accepts for
//             MN-major tf32 operands (cutlass sm100_common.inl:92).  Atoms of 32 (mn) x 4 (k) = 512 bytes:
//             k-row kr is 128 contiguous bytes (32 mn values), its four 32-byte chunks XOR kr
//             (Swizzle<2,5,2>: byte-address bits [5,7) ^= bits [7,9)).  Atom (g = mn/32, kg = k/4) sits at
//             (kg*(ROWS/32) + g)*512: LBO (mn-group stride) = 512, SBO (k-group stride) = (ROWS/32)*512.
//             A K=8 MMA slice j covers k-groups 2j, 2j+1 and starts at j*2*SBO.
template <int ROWS, bool KMAJOR>
struct UmmaTile {
  static constexpr int BK = 32;
  static constexpr int BYTES = ROWS * BK * 4;
  static constexpr int LBO = KMAJOR ? 16 : 512;                       // K-major: unused by hw (canonical value 1 unit)
  static constexpr int SBO = KMAJOR ? 1024 : (ROWS / 32) * 512;
  static constexpr int LAYOUT_TYPE = KMAJOR ? 2 : 1;
  // byte offset of the 16-byte chunk a loader group writes.
  //   K-major : idx = row, k multiple of 4      MN-major : idx = row (multiple of 4), single k
  __device__ static __forceinline__ int chunk_off(int idx, int k) {
    if (KMAJOR) return idx * 128 + ((((k >> 2) ^ idx) & 7) << 4);
    const int kr = k & 3, c32 = (idx & 31) >> 3, half = (idx & 7) >> 2;
    return ((k >> 2) * (ROWS / 32) + (idx >> 5)) * 512 + kr * 128 + ((c32 ^ kr) << 5) + (half << 4);
  }
  // descriptor start offset of the j-th K=8 slice
  __device__ static __forceinline__ int kslice_off(int j) { return KMAJOR ? j * 32 : j * 2 * SBO; }
};

template <int BN_, int STAGES_, int MINB_ = 1, int PW_ = 4, int LW_ = 0>
struct UmmaCfg {
  static constexpr int BM = 128, BN = BN_, BK = 32, STAGES = STAGES_, MINB = MINB_;
  static constexpr int PW = PW_;            // producer / epilogue warps (4 or 8), warps 0..PW-1
  static constexpr int LW = LW_;            // 1: warp PW+1 issues the bulk copies of a pre-tiled B (see the kernel)
  static constexpr int NPROD = PW * 32;
  static constexpr int NT = NPROD + 32 + 32 * LW;   // + the MMA warp (warp PW) [+ the B loader warp]
  static constexpr int EPI_COLS = BN / (PW / 4);   // accumulator columns each producer warp drains
  static_assert(PW == 4 || PW == 8, "4 or 8 producer warps");
  static_assert(EPI_COLS % 32 == 0, "epilogue reads 32 columns at a time");
  static constexpr int TMEM_COLS = BN;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two in [32,256]");
};

// A loader may declare `static constexpr bool kExactTf32 = true` when its values are exactly representable in
// tf32 (the uint8 frame bytes): then lo == 0, the lo tile is not written and the A_lo*B_hi product is skipped
// ("2xTF32").
template <class L, class = void>
struct loader_exact : std::false_type {};
template <class L>
struct loader_exact<L, std::void_t<decltype(L::kExactTf32)>> : std::bool_constant<L::kExactTf32> {};

// A loader may also declare `kVec16`: it then provides load_raw16()/unpack16() that fetch 16 consecutive
// elements of the contiguous dimension with ONE 128-bit load (the uint8 frames: 4x fewer load instructions).
template <class L, class = void>
struct loader_vec16 : std::false_type {};
template <class L>
struct loader_vec16<L, std::void_t<decltype(L::kVec16)>> : std::bool_constant<L::kVec16> {};

// A B loader may declare `kPretiled`: the operand (a weight matrix) has already been split into hi/lo and stored in
// global memory as per-(n-tile, k-tile) IMAGES of the shared-memory stage layout (retile_b_kernel below, run once
// per parameter update).  The producers then do not touch B at all: one thread fetches [B_hi | B_lo] of a stage
// with a single cp.async.bulk (TMA bulk copy) that completes on the stage's full barrier (complete_tx).
template <class L, class = void>
struct loader_pretiled : std::false_type {};
template <class L>
struct loader_pretiled<L, std::void_t<decltype(L::kPretiled)>> : std::bool_constant<L::kPretiled> {};

template <class Base>
struct PretiledB {
  static constexpr bool kContigK = Base::kContigK;
  static constexpr bool kPretiled = true;
  const uint8_t* image;   // [n-tiles][k-tiles][2 * B_BYTES]
  int ktiles;             // K tiles (of 32) in the image
  struct Row {};
  __device__ __forceinline__ Row row(int, int) const { return Row{}; }
  __device__ __forceinline__ float4 load(const Row&, int) const { return zero4(); }
};

namespace umma {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace umma

// Epilogue store of one 32x32 accumulator block.  tcgen05.ld hands every lane one ROW (v[j] = D[row0 + lane][col0 + j]);
// storing that directly makes each 128-bit store instruction touch 32 different 128-byte lines (ncu: the LSU data
// pipe was the busiest unit of the epilogue-heavy kernels).  The block is transposed through a per-warp 4 KB
// shared-memory tile instead (16-byte chunks XOR-swizzled by row & 7: conflict-free both ways), so that 8 lanes
// cover one 128-byte row segment and an instruction touches 4 lines.  The functor sees the same (m, n) pairs.
constexpr int kEpiStageBytes = 32 * 32 * 4;
// An epilogue functor may provide `bias4(z, n)` + `store4b(z, m, n, o, bias)`: the per-column bias of a lane is the same
// for all 8 rows it stores, so it is fetched ONCE per 32x32 block instead of once per element.  (Measured on the conv1
// kernel: the streaming stores of the output evict the bias line from L1, each of the 8 row iterations then paid an L2
// round trip for it, and the epilogue of one 128x32 tile took 11,000 cycles -- tools/conv1_timeline.py.)
template <class EP, class = void>
struct ep_has_bias4 : std::false_type {};
template <class EP>
struct ep_has_bias4<EP, std::void_t<decltype(std::declval<const EP&>().bias4(0, 0))>> : std::true_type {};

template <class EP>
__device__ __forceinline__ void epilogue_store_32x32(const EP& ep, uint8_t* stg, int lane, int z, int row0, int col0,
                                                     int M, int N, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int c = lane & 7, rr = lane >> 3;
  const int n = col0 + c * 4;
  float4 bias = zero4();
  if constexpr (ep_has_bias4<EP>::value) {
    if (n + 3 < N) bias = ep.bias4(z, n);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rr;
    const float4 o4 = *reinterpret_cast<const float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
    const int m = row0 + r;
    if (m < M) {
      const float o[4] = {o4.x, o4.y, o4.z, o4.w};
      if (n + 3 < N) {
        if constexpr (ep_has_bias4<EP>::value) ep.store4b(z, m, n, o, bias);
        else ep.template store<4>(z, m, n, o);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n + q < N) {
            const float o1[1] = {o[q]};
            ep.template store<1>(z, m, n + q, o1);
          }
      }
    }
  }
  __syncwarp();
}

template <class Cfg, class AL, class BL>
struct UmmaSmem {
  using TA = UmmaTile<Cfg::BM, AL::kContigK>;
  using TB = UmmaTile<Cfg::BN, BL::kContigK>;
  static constexpr bool AEX = loader_exact<AL>::value;
  static constexpr int A_BYTES = TA::BYTES, B_BYTES = TB::BYTES;       // multiples of 1024
  static constexpr int STAGE_BYTES = (AEX ? 1 : 2) * A_BYTES + 2 * B_BYTES;   // hi (and lo) copies of the operands
  static constexpr int AUX_BYTES = 1024 + (BL::kContigK ? 0 : Cfg::NPROD * 16);   // barriers, tmem ptr, colsum scratch
  static constexpr int BYTES = Cfg::STAGES * STAGE_BYTES + AUX_BYTES + 1024;      // + alignment slack
};

template <class Cfg, class AL, class BL, class EP>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB)
gemm_umma_kernel(const AL al, const BL bl, const EP ep, int M, int N, int K, int kchunk, int kstep) {
  pdl_prologue();
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, NPROD = Cfg::NPROD;
  constexpr bool AK = AL::kContigK, BKc = BL::kContigK;
  using SM = UmmaSmem<Cfg, AL, BL>;
  using TA = typename SM::TA;
  using TB = typename SM::TB;
  constexpr bool AEX = SM::AEX;
  constexpr int NGA = BM * BK / 4, NGB = BN * BK / 4;       // float4 groups per stage
  constexpr bool A16 = loader_vec16<AL>::value;              // 16-element (one 128-bit load) A groups
  constexpr bool BPT = loader_pretiled<BL>::value;           // B arrives as pre-split stage images via cp.async.bulk
  // A cp.async.bulk blocks its issuing thread for ~0.24 us whatever its size (tools/microbench/tma_box_bw.cu).  Issued
  // by producer thread 0 that delay sits in front of its own gathers and so of every stage's full barrier; with
  // Cfg::LW a dedicated warp (PW+1) issues the copies instead and arrives on the barrier itself.
  constexpr bool LWB = BPT && Cfg::LW;
  constexpr int GA = (A16 ? NGA / 4 : NGA) / NPROD, GB = BPT ? 1 : NGB / NPROD;   // groups per producer thread
  static_assert((A16 ? NGA / 4 : NGA) % NPROD == 0 && NGB % NPROD == 0, "groups must divide among producers");
  static_assert(!A16 || AEX, "16-wide raw loads are only used for exact (uint8) operands");
  using ARaw = typename std::conditional<A16, uint4, float4>::type;
  constexpr bool kColSum = EP::kColSum && !BKc && !BPT;
  constexpr int OFF_ALO = SM::A_BYTES, OFF_BHI = (AEX ? 1 : 2) * SM::A_BYTES, OFF_BLO = OFF_BHI + SM::B_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux = smem + STAGES * SM::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  float4* cs_scratch = reinterpret_cast<float4*>(aux + 1024);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k0 = z * kstep;
  const int k1 = min(K, k0 + kchunk);
  const int ntiles = (k1 - k0 + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(&full[s], NPROD + (LWB ? 1 : 0));
      umma::mbar_init(&empty[s], 1);
    }
    umma::mbar_init(acc_full, 1);
    umma::fence_barrier_init();
  }
  if (warp == Cfg::PW) umma::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < Cfg::PW) {
    // ================= PRODUCERS =================
    // K-major operand : group g -> row g/8, chunk g%8  (8 lanes cover one 128-byte row: coalesced global
    //                   read of 128 contiguous bytes, conflict-free swizzled 16-byte stores)
    // MN-major operand: group g -> k = g/(ROWS/4), quad q = g%(ROWS/4)  (8 lanes cover 32 contiguous mn
    //                   values of one k: coalesced, conflict-free)
    typename AL::Row arow[GA];
    typename BL::Row brow[GB];
    int a_k[GA], a_o[GA], b_k[GB], b_o[GB];
#pragma unroll
    for (int i = 0; i < GA; ++i) {
      const int g = tid + i * NPROD;
      if (A16) {
        if (AK) {   // group = (row, 16-k half): 32 lanes = 32 consecutive rows -> coalesced, conflict-free stores
          const int r = g % BM, half = g / BM;
          arow[i] = al.row(z, (m0 + r < M) ? m0 + r : -1);
          a_k[i] = half * 16;
          a_o[i] = r;
        } else {    // group = (k, 16 consecutive mn): 32 lanes = the 32 k of the tile for the same 16 mn values
          const int kk = g % BK, q16 = g / BK;
          arow[i] = al.row(z, (m0 + q16 * 16 < M) ? m0 + q16 * 16 : -1);
          a_k[i] = kk;
          a_o[i] = q16 * 16;
        }
      } else if (AK) {
        const int r = g >> 3, kq = g & 7;
        arow[i] = al.row(z, (m0 + r < M) ? m0 + r : -1);
        a_k[i] = kq * 4;
        a_o[i] = TA::chunk_off(r, kq * 4);
      } else {
        const int q = g % (BM / 4), kk = g / (BM / 4);
        arow[i] = al.row(z, (m0 + q * 4 < M) ? m0 + q * 4 : -1);
        a_k[i] = kk;
        a_o[i] = TA::chunk_off(q * 4, kk);
      }
    }
#pragma unroll
    for (int i = 0; i < GB; ++i) {
      const int g = tid + i * NPROD;
      if (BPT) {
        b_k[i] = 0; b_o[i] = 0;
      } else if (BKc) {
        const int r = g >> 3, kq = g & 7;
        brow[i] = bl.row(z, (n0 + r < N) ? n0 + r : -1);
        b_k[i] = kq * 4;
        b_o[i] = TB::chunk_off(r, kq * 4);
      } else {
        const int q = g % (BN / 4), kk = g / (BN / 4);
        brow[i] = bl.row(z, (n0 + q * 4 < N) ? n0 + q * 4 : -1);
        b_k[i] = kk;
        b_o[i] = TB::chunk_off(q * 4, kk);
      }
    }
    float4 csum[GB];
#pragma unroll
    for (int i = 0; i < GB; ++i) csum[i] = zero4();

    // gather tile t into registers
    auto gload = [&](int t, ARaw (&ra)[GA], float4 (&rb)[GB]) {
      const int kb = k0 + t * BK;
#pragma unroll
      for (int i = 0; i < GA; ++i) {
        const int k = kb + a_k[i];
        if constexpr (A16) ra[i] = (k < k1) ? al.load_raw16(arow[i], k) : make_uint4(0u, 0u, 0u, 0u);
        else ra[i] = (k < k1) ? al.load(arow[i], k) : zero4();
      }
      if constexpr (!BPT) {
#pragma unroll
        for (int i = 0; i < GB; ++i) {
          const int k = kb + b_k[i];
          rb[i] = (k < k1) ? bl.load(brow[i], k) : zero4();
        }
      }
    };
    // split tile t into hi/lo and publish it in its shared-memory stage
    auto publish = [&](int t, const ARaw (&ra)[GA], const float4 (&rb)[GB]) {
      const int s = t % STAGES;
      const uint32_t ph = (t / STAGES) & 1;
      umma::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* st = smem + s * SM::STAGE_BYTES;
      if constexpr (BPT && !LWB) {
        if (tid == 0) {   // one bulk copy brings [B_hi | B_lo] of this (n-tile, k-tile); bytes complete on full[s]
          const int kt = (k0 + t * BK) / BK;
          umma::mbar_expect_tx(&full[s], 2 * SM::B_BYTES);
          umma::bulk_g2s(st + OFF_BHI, bl.image + ((size_t)blockIdx.y * bl.ktiles + kt) * (size_t)(2 * SM::B_BYTES),
                         2 * SM::B_BYTES, &full[s]);
        }
      }
#pragma unroll
      for (int i = 0; i < GA; ++i) {
        if constexpr (A16) {
          float4 f[4];
          AL::unpack16(ra[i], f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int off = AK ? TA::chunk_off(a_o[i], a_k[i] + 4 * j) : TA::chunk_off(a_o[i] + 4 * j, a_k[i]);
            *reinterpret_cast<float4*>(st + off) = f[j];
          }
        } else if constexpr (AEX) {
          *reinterpret_cast<float4*>(st + a_o[i]) = ra[i];
        } else {
          float4 h, l;
          umma::split4(ra[i], h, l);
          *reinterpret_cast<float4*>(st + a_o[i]) = h;
          *reinterpret_cast<float4*>(st + OFF_ALO + a_o[i]) = l;
        }
      }
      if constexpr (!BPT) {
#pragma unroll
        for (int i = 0; i < GB; ++i) {
          float4 h, l;
          umma::split4(rb[i], h, l);
          *reinterpret_cast<float4*>(st + OFF_BHI + b_o[i]) = h;
          *reinterpret_cast<float4*>(st + OFF_BLO + b_o[i]) = l;
          if (kColSum) { csum[i].x += rb[i].x; csum[i].y += rb[i].y; csum[i].z += rb[i].z; csum[i].w += rb[i].w; }
        }
      }
      umma::fence_proxy_async();       // generic-proxy writes -> visible to the tensor-core (async) proxy
      umma::mbar_arrive(&full[s]);
    };

    // register prefetch ring of depth PF: the gathers of tiles t+1 .. t+PF-1 are in flight while tile t is split
    // and stored (these kernels are bound by per-K-tile latency, not by producer throughput).  The uint8 operand
    // needs 8 registers per tile, and a pre-tiled B needs none, so those cases afford a deeper ring.
    constexpr int PF = A16 ? 4 : (BPT ? 3 : 2);
    ARaw ra[PF][GA];
    float4 rb[PF][GB];
#pragma unroll
    for (int d = 0; d < PF - 1; ++d)
      if (d < ntiles) gload(d, ra[d], rb[d]);
#pragma unroll 1
    for (int t0 = 0; t0 < ntiles; t0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int t = t0 + u;
        if (t < ntiles) {
          if (t + PF - 1 < ntiles) gload(t + PF - 1, ra[(u + PF - 1) % PF], rb[(u + PF - 1) % PF]);
          publish(t, ra[u], rb[u]);
        }
      }
    }

    // bias-gradient row: column sums of B (dY) over this CTA's K range, reduced in fixed order
    if (kColSum && blockIdx.x == 0) {
      // thread's groups: q = (tid + i*NPROD) % (BN/4).  NPROD is a multiple of BN/4, so q = tid % (BN/4) for all i.
      float4 acc = zero4();
#pragma unroll
      for (int i = 0; i < GB; ++i) { acc.x += csum[i].x; acc.y += csum[i].y; acc.z += csum[i].z; acc.w += csum[i].w; }
      cs_scratch[tid] = acc;
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
      if (tid < BN / 4) {
        float4 tot = zero4();
        for (int j = tid; j < NPROD; j += BN / 4) {
          const float4 v = cs_scratch[j];
          tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        const int n = n0 + tid * 4;
        if (n < N) ep.store_colsum(z, n, tot.x);
        if (n + 1 < N) ep.store_colsum(z, n + 1, tot.y);
        if (n + 2 < N) ep.store_colsum(z, n + 2, tot.z);
        if (n + 3 < N) ep.store_colsum(z, n + 3, tot.w);
      }
    }

    // ================= EPILOGUE =================
    umma::mbar_wait(acc_full, 0);
    umma::tc_fence_after();
    // warp w may only touch TMEM lanes 32*(w%4)..+31; with 8 warps the two warps of a lane quarter split the columns.
    // Every MMA has completed, so the pipeline stages are dead: their first bytes become the per-warp staging tiles.
    uint8_t* stg = smem + warp * kEpiStageBytes;
    const int row0 = m0 + (warp & 3) * 32;
    const int cbeg = (warp >> 2) * Cfg::EPI_COLS;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + Cfg::EPI_COLS; c0 += 32) {
      float v[32];
      umma::tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
      epilogue_store_32x32(ep, stg, lane, z, row0, n0 + c0, M, N, v);
    }
    umma::tc_fence_before();
  } else if (warp > Cfg::PW) {
    // ================= B LOADER (warp PW+1, Cfg::LW) =================
    if constexpr (LWB) {
      if (lane == 0) {
        for (int t = 0; t < ntiles; ++t) {
          const int s = t % STAGES;
          umma::mbar_wait(&empty[s], ((t / STAGES) & 1) ^ 1);
          const int kt = (k0 + t * BK) / BK;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&full[s])),
                       "r"(2 * SM::B_BYTES) : "memory");
          umma::bulk_g2s(smem + s * SM::STAGE_BYTES + OFF_BHI,
                         bl.image + ((size_t)blockIdx.y * bl.ktiles + kt) * (size_t)(2 * SM::B_BYTES), 2 * SM::B_BYTES,
                         &full[s]);
        }
      }
    }
  } else {
    // ================= MMA ISSUER (warp PW) =================
    constexpr uint32_t idesc = umma::make_idesc(BN, !AK, !BKc);
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % STAGES;
      const uint32_t ph = (t / STAGES) & 1;
      umma::mbar_wait(&full[s], ph);
      umma::tc_fence_after();
      if (umma::elect_one()) {
        const uint32_t st = umma::smem_u32(smem + s * SM::STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + OFF_ALO, b_hi = st + OFF_BHI, b_lo = st + OFF_BLO;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint32_t ao = TA::kslice_off(j), bo = TB::kslice_off(j);
          const uint64_t dah = umma::make_desc(a_hi + ao, TA::LBO, TA::SBO, TA::LAYOUT_TYPE);
          const uint64_t dbh = umma::make_desc(b_hi + bo, TB::LBO, TB::SBO, TB::LAYOUT_TYPE);
          const uint64_t dbl = umma::make_desc(b_lo + bo, TB::LBO, TB::SBO, TB::LAYOUT_TYPE);
          const uint32_t first = (t > 0 || j > 0) ? 1u : 0u;
          if (!AEX) {                                                    // small terms first
            const uint64_t dal = umma::make_desc(a_lo + ao, TA::LBO, TA::SBO, TA::LAYOUT_TYPE);
            umma::mma_tf32(tmem_base, dal, dbh, idesc, first);
            umma::mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          } else {
            umma::mma_tf32(tmem_base, dah, dbl, idesc, first);
          }
          umma::mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        umma::mma_commit(&empty[s]);            // frees the smem slot once these MMAs have read it
        if (t == ntiles - 1) umma::mma_commit(acc_full);
      }
      __syncwarp();
    }
    umma::tc_fence_before();
  }
  __syncthreads();
  if (warp == Cfg::PW) {
    umma::tc_fence_after();
    umma::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// Writes the pre-split, pre-tiled, pre-swizzled image of a B operand: for every (n-tile, k-tile) the exact bytes the
// producers would have put into [B_hi | B_lo] of a stage.  Runs once per parameter update.
template <int BN, class BL>
__global__ void __launch_bounds__(256) retile_b_kernel(const BL bl, int N, int K, int ktiles, int ntn,
                                                        uint8_t* __restrict__ image) {
  pdl_prologue();
  using TB = UmmaTile<BN, BL::kContigK>;
  constexpr int GPT = BN * 8;                         // float4 groups per tile (BN rows x 32 floats)
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)ntn * ktiles * GPT) return;
  const int tile = (int)(g / GPT), lg = (int)(g % GPT);
  const int nt = tile / ktiles, kt = tile - nt * ktiles;
  int n, k, off;
  if (BL::kContigK) {
    const int r = lg >> 3, kq = lg & 7;
    n = nt * BN + r; k = kt * 32 + kq * 4; off = TB::chunk_off(r, kq * 4);
  } else {
    const int q = lg % (BN / 4), kk = lg / (BN / 4);
    n = nt * BN + q * 4; k = kt * 32 + kk; off = TB::chunk_off(q * 4, kk);
  }
  const typename BL::Row row = bl.row(0, n < N ? n : -1);
  const float4 v = (k < K) ? bl.load(row, k) : zero4();
  float4 h, l;
  umma::split4(v, h, l);
  uint8_t* dst = image + (size_t)tile * (2 * TB::BYTES);
  *reinterpret_cast<float4*>(dst + off) = h;
  *reinterpret_cast<float4*>(dst + TB::BYTES + off) = l;
}

template <int BN>
inline size_t weight_image_bytes(int N, int K) {
  return (size_t)cdiv(N, BN) * cdiv(K, 32) * 2 * BN * 128;
}

template <int BN, class BL>
inline int launch_retile_b(cudaStream_t s, const BL& bl, int N, int K, uint8_t* image) {
  const int ntn = cdiv(N, BN), ktiles = cdiv(K, 32);
  const long long groups = (long long)ntn * ktiles * BN * 8;
  DRL_CUDA_CHECK((launch_k(retile_b_kernel<BN, BL>, (unsigned)cdiv64(groups, 256), 256, 0, s, bl, N, K, ktiles, ntn, image)));
  return DRL_OK;
}

template <class Cfg, class AL, class BL, class EP>
inline int launch_gemm_umma(cudaStream_t s, const AL& al, const BL& bl, const EP& ep, int M, int N, int K,
                            int zcount, int kchunk, int kstep) {
  using SM = UmmaSmem<Cfg, AL, BL>;
  static bool attr_done = false;
  auto kern = gemm_umma_kernel<Cfg, AL, BL, EP>;
  if (!attr_done) {
    DRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES));
    attr_done = true;
  }
  if ((AL::kContigK || BL::kContigK) && (K % 4 != 0 || kchunk % 4 != 0)) {
    set_error("gemm_umma: K (%d) and kchunk (%d) must be multiples of 4 for K-contiguous operands", K, kchunk);
    return DRL_ERR_INVALID;
  }
  if (kchunk < 1 || K < 1) { set_error("gemm_umma: empty K range"); return DRL_ERR_INVALID; }
  dim3 grid(cdiv(M, Cfg::BM), cdiv(N, Cfg::BN), zcount);
  DRL_CUDA_CHECK((launch_k(kern, grid, Cfg::NT, SM::BYTES, s, al, bl, ep, M, N, K, kchunk, kstep)));
  return DRL_OK;
}

}  // namespace drl
