// layer_defs.cuh -- loader / tile-configuration instantiations, launch macros and split-K plans shared by the
// IMPALA actor-critic (layers.cu) and the Ape-X dueling network (apex.cu): both run the same attention_CNN
// (model/impala_actor_critic.py:5-10 == model/apex_value.py:4-9) through the same gather-GEMM kernels.
#pragma once
#include <algorithm>

#include "gemm_simt.cuh"
#include "gemm_umma.cuh"
#include "gemm_umma16.cuh"
#include "conv1_tma.cuh"
#include "gemm_bulk16.cuh"
#include "gemm_tma.cuh"
#include "gemm_umma_persist.cuh"
#include "kernels.h"
#include "loaders.cuh"

namespace drl {

// ------------------------------------------------------------------------------------------
// Loader / epilogue instantiations for the three convolutions
// ------------------------------------------------------------------------------------------
using Conv1A = ConvFwdA<uint8_t, 84, 84, 4, 20, 20, 8, 4, true>;
using Conv2A = ConvFwdA<float, 20, 20, 32, 9, 9, 4, 2, false>;
using Conv3A = ConvFwdA<float, 9, 9, 64, 7, 7, 3, 1, false>;
// math mode 4: TMA-fed forward of conv2 / conv3 (gemm_tma.cuh): tiles of 1 image (81 rows) / 2 images (98 rows), 4 stages.
// Correct and tested, but measured SLOWER than the software producers on this network (conv2 59 vs 46 us, conv3 41 vs
// 33 us at B=32, T=20): with 32/64 input channels a K tile is one filter tap, so every activation byte is fetched
// 4x (conv2) / 9x (conv3) from L2 for each of the two planes plus a 16 KB weight tile per tap and image -- 2.3x the
// L2->SM bytes of the LDG path, which reads raw fp32 once per tap and splits in registers.
using Conv2Tma = ConvTmaCfg<32, 20, 9, 4, 2, 1, 4>;
using Conv3Tma = ConvTmaCfg<64, 9, 7, 3, 1, 2, 4>;
using Conv1WA = ConvWgradA<uint8_t, 84, 84, 4, 20, 20, 8, 4, true>;
using Conv2WA = ConvWgradA<float, 20, 20, 32, 9, 9, 4, 2, false>;
using Conv3WA = ConvWgradA<float, 9, 9, 64, 7, 7, 3, 1, false>;
using Conv3DA = ConvDgradA<7, 7, 64, 1, 3, 3, 9, 9>;
using Conv3DB = ConvDgradB<64, 64, 1, 3, 3>;
using Conv3DE = EpConvDx<9, 9, 64, 1, 9, 9>;
using Conv2DA = ConvDgradA<9, 9, 64, 2, 4, 4, 10, 10>;
using Conv2DB = ConvDgradB<32, 64, 2, 4, 4>;
using Conv2DE = EpConvDx<20, 20, 32, 2, 10, 10>;

using U32 = UmmaCfg<32, 2, 3>;   // 24 KB / stage (A exact: conv1 only), 3 CTAs per SM
using U64 = UmmaCfg<64, 2, 2>;   // 48 KB / stage, 2 CTAs per SM
using U128 = UmmaCfg<128, 3, 1, 8>;   // 64 KB / stage, 1 CTA per SM, 8 producer warps
using U256 = UmmaCfg<256, 2, 1, 8>;   // 96 KB / stage, 1 CTA per SM, 8 producer warps
// the same with a B-loader warp, for GEMMs whose B operand is a pre-tiled weight image
using U64L = UmmaCfg<64, 2, 2, 4, 1>;
using U128L = UmmaCfg<128, 3, 1, 8, 1>;
using U256L = UmmaCfg<256, 2, 1, 8, 1>;

// math mode 5: 16-bit split operands (gemm_umma16.cuh).  One format per GEMM (f16 x bf16 traps).  bf16 keeps fp32's
// exponent range (gradients span many decades); fp16 has three more mantissa bits but underflows gradually below ~1e-4
// and saturates at 65504 -- it is used where one operand is EXACT in it (the uint8 frames of conv1) and the other is
// an O(0.1) weight.
using Fmt16 = umma16::BF16;     // every GEMM whose operands are fp32 activations / gradients / weights
using FmtC1 = umma16::F16;      // conv1: uint8 frames (exact) x weights
using X32L = Umma16Cfg<32, 4, 2, 4, 1, FmtC1, FmtC1>;        // conv1 fwd: whole K = 256 resident
using X64L = Umma16Cfg<64, 2, 2, 4, 1, Fmt16, Fmt16>;        // conv2 / conv3 fwd: 48 KB per K = 64 stage, 2 CTAs per SM
using X64L8x2 = Umma16Cfg<64, 2, 2, 8, 1, Fmt16, Fmt16>;     // conv2 / conv3 fwd (default): two CTAs per SM x 8 producer warps = twice the loads
                                                             // in flight of X64L (conv2_fwd 41 -> 33 us: the gathers are latency-bound); DRL_B200_C2F=0: X64L
using X64W8x2 = Umma16Cfg<64, 2, 2, 8, 0, Fmt16, Fmt16>;     // conv2 / conv3 weight gradients (default; 39 -> 33, 30 -> 27 us); DRL_B200_CW8=0: X64W
using X64W = Umma16Cfg<64, 2, 2, 4, 0, Fmt16, Fmt16>;        // conv2 / conv3 weight gradients (both operands gathered)
using XH = Umma16Cfg<64, 4, 1, 8, 0, Fmt16, Fmt16>;          // head layers (K = 256 = all four stages in flight), DRL_B200_HEADS_TC=1
// bulk-fed LSTM weight gradient (gemm_bulk16.cuh): both operands are pre-tiled images, no producer warps
using BK256 = Bulk16Cfg<256, 2, 8>;    // 96 KB per stage (A 32 KB + B 64 KB), 2 stages
using X256L = Umma16Cfg<256, 2, 1, 8, 1, Fmt16, Fmt16>;      // lstm fwd: 96 KB per stage
using X256W = Umma16Cfg<256, 2, 1, 8, 0, Fmt16, Fmt16>;      // lstm weight gradient
using X128D = Umma16Cfg<128, 3, 1, 8, 1, Fmt16, Fmt16>;      // lstm data gradient: 64 KB per stage
using X256D = Umma16Cfg<256, 1, 2, 8, 1, Fmt16, Fmt16>;      // dCol GEMMs (K = 64 = ONE stage of 96 KB): two CTAs per SM, so that
                                                             // the 128 KB epilogue of one overlaps the load + MMAs of the other
using X64G = Umma16Cfg<64, 2, 2, 4, 0, Fmt16, Fmt16>;        // conv3 data gradient, gather form (K = 9 taps x 64)
using X32G = Umma16Cfg<32, 3, 2, 4, 0, Fmt16, Fmt16>;        // conv2 data gradient, gather form (4 parity classes, K = 4 taps x 64)

// math mode 3 (experimental): the persistent, fully warp-specialised variant of each tensor-core configuration
template <class U> struct PersistOf;
template <> struct PersistOf<U32> { using type = UmmaPCfg<32, 5, 8>; };
template <> struct PersistOf<U64> { using type = UmmaPCfg<64, 4, 8>; };
template <> struct PersistOf<U128> { using type = UmmaPCfg<128, 3, 8>; };
template <> struct PersistOf<U256> { using type = UmmaPCfg<256, 2, 8>; };
template <> struct PersistOf<U64L> { using type = UmmaPCfg<64, 4, 8>; };
template <> struct PersistOf<U128L> { using type = UmmaPCfg<128, 3, 8>; };
template <> struct PersistOf<U256L> { using type = UmmaPCfg<256, 2, 8>; };


// name the launch for the per-kernel profile, run it on the selected core, count it
#define GEMM(name, SCfg, UCfg, ...)                          \
  do {                                                       \
    prof_mark(s, name);                                      \
    if (mode == 3) {                                         \
      DRL_TRY((launch_gemm_umma_persist<typename PersistOf<UCfg>::type>(s, __VA_ARGS__))); \
    } else if (mode == 2 || mode == 5) {                     \
      DRL_TRY((launch_gemm_umma<UCfg>(s, __VA_ARGS__)));     \
    } else {                                                 \
      DRL_TRY((launch_gemm_simt<SCfg>(s, __VA_ARGS__)));     \
    }                                                        \
    ++n;                                                     \
  } while (0)
// a GEMM on the 16-bit split core (math mode 5)
#define GEMM16(name, XCfg, ...)                              \
  do {                                                       \
    prof_mark(s, name);                                      \
    DRL_TRY((launch_gemm_umma16<XCfg>(s, __VA_ARGS__)));     \
    ++n;                                                     \
  } while (0)
// same, for GEMMs whose B operand is a weight matrix: the tensor-core cores read its pre-tiled image (blp)
#define GEMM_W(name, SCfg, UCfg, al, bl, blp, ...)            \
  do {                                                       \
    prof_mark(s, name);                                      \
    if (mode == 3) {                                         \
      DRL_TRY((launch_gemm_umma_persist<typename PersistOf<UCfg>::type>(s, al, blp, __VA_ARGS__))); \
    } else if (mode == 2) {                                  \
      DRL_TRY((launch_gemm_umma<UCfg>(s, al, blp, __VA_ARGS__))); \
    } else {                                                 \
      DRL_TRY((launch_gemm_simt<SCfg>(s, al, bl, __VA_ARGS__))); \
    }                                                        \
    ++n;                                                     \
  } while (0)
#define GEMM_FFMA(name, SCfg, ...)                           \
  do {                                                       \
    prof_mark(s, name);                                      \
    DRL_TRY((launch_gemm_simt<SCfg>(s, __VA_ARGS__)));       \
    ++n;                                                     \
  } while (0)
// the small dense layers of the heads: in-CTA split-K (KS groups work on the K tiles of one output tile concurrently);
// DRL_B200_HEADS_KS=1 restores the one-group kernel
#define GEMM_FFMA_KS(name, SCfg, KS, ...)                      \
  do {                                                       \
    static const int ks_env = getenv("DRL_B200_HEADS_KS") ? atoi(getenv("DRL_B200_HEADS_KS")) : 0; \
    prof_mark(s, name);                                      \
    if (ks_env == 1) DRL_TRY((launch_gemm_simt<SCfg>(s, __VA_ARGS__)));  \
    else DRL_TRY((launch_gemm_simt<SCfg, KS>(s, __VA_ARGS__)));        \
    ++n;                                                     \
  } while (0)
#define KERNEL(name, call, cnt)                              \
  do {                                                       \
    prof_mark(s, name);                                      \
    DRL_TRY(call);                                           \
    n += (cnt);                                              \
  } while (0)

// split-K plan: about `waves` waves of CTAs on 148 SMs
struct SplitPlan { int splits, kchunk; };
static SplitPlan plan_split(int K, int tiles_mn, int bk, int waves) {
  int target = (waves * 148 + tiles_mn - 1) / tiles_mn;
  if (target < 1) target = 1;
  int kchunk = (K + target - 1) / target;
  kchunk = (kchunk + bk - 1) / bk * bk;
  if (kchunk < bk) kchunk = bk;
  SplitPlan p;
  p.kchunk = kchunk;
  p.splits = (K + kchunk - 1) / kchunk;
  return p;
}
// the conv weight-gradient plans (mode 1: FFMA tiles, 2 waves; mode 2: 128-row UMMA tiles, 2 CTAs/SM)
static SplitPlan plan_conv1_wgrad(int Mb, int mode) {
  return mode >= 2 ? plan_split(Mb * 400, 2, 32, 2) : plan_split(Mb * 400, cdiv(256, CfgWg1::BM), 16, 2);
}
static SplitPlan plan_conv2_wgrad(int Mb, int mode) {
  if (mode == 5) return plan_split(Mb * 81, 4, 64, 2);
  return mode >= 2 ? plan_split(Mb * 81, 4, 32, 2) : plan_split(Mb * 81, cdiv(512, CfgBig::BM), 16, 2);
}
static SplitPlan plan_conv3_wgrad(int Mb, int mode) {
  if (mode == 5) return plan_split(Mb * 49, 5, 64, 2);
  return mode >= 2 ? plan_split(Mb * 49, 5, 32, 2) : plan_split(Mb * 49, cdiv(576, CfgBig::BM), 16, 2);
}

// Fork/join helpers: kernels that are off the critical path (the action-embedding table in the forward, every
// weight gradient except conv1's in the backward) run on the side stream so they fill SMs the critical
// dgrad chain leaves idle.  Inside CUDA-graph capture the event record/wait pairs become graph edges.
static inline int fork_to_side(const Streams& st, int i) {
  if (!st.par) return DRL_OK;
  DRL_CUDA_CHECK(cudaEventRecord(st.ev[i], st.main));
  DRL_CUDA_CHECK(cudaStreamWaitEvent(st.side, st.ev[i], 0));
  pdl_break(st.side);   // the next side-stream kernel depends on a kernel of another stream: full dependency
  return DRL_OK;
}
// main -> side2 edge (ev2[i]); side2 -> other stream edge
static inline int fork_to_side2(const Streams& st, int i) {
  if (!st.par || !st.side2) return DRL_OK;
  DRL_CUDA_CHECK(cudaEventRecord(st.ev2[i], st.main));
  DRL_CUDA_CHECK(cudaStreamWaitEvent(st.side2, st.ev2[i], 0));
  pdl_break(st.side2);
  return DRL_OK;
}
static inline int join_side2_into(const Streams& st, int i, cudaStream_t into) {
  if (!st.par || !st.side2) return DRL_OK;
  DRL_CUDA_CHECK(cudaEventRecord(st.ev2[i], st.side2));
  DRL_CUDA_CHECK(cudaStreamWaitEvent(into, st.ev2[i], 0));
  pdl_break(into);
  return DRL_OK;
}
static inline int join_from_side(const Streams& st, int i) {
  if (!st.par) return DRL_OK;
  DRL_CUDA_CHECK(cudaEventRecord(st.ev[i], st.side));
  DRL_CUDA_CHECK(cudaStreamWaitEvent(st.main, st.ev[i], 0));
  pdl_break(st.main);
  return DRL_OK;
}

}  // namespace drl
