// chain.cuh -- parameter block of the chained per-symbol GEMM kernel (chain.cu).
//
// equalizer_ofdm has runs of per-symbol layers whose intermediate never needs to leave the SM
// (dev/py/model.py:370-379  dense -> (1,K) complex conv;  :437-462  (1,K) conv(eq) | (1,K) conv(corr) -> concat -> dense_5).
// The layer-by-layer schedule wrote every intermediate to HBM and read it back (cat alone: 2 x 470 MB per 65 536-frame
// pass).  chain_tc_kernel keeps the layer-by-layer ARITHMETIC -- every layer's output is rounded to fp32 (bias added in
// fp32, partial sums of the k-blocks added in fp32 round-to-nearest in the same order) before it is split into the fp16
// (hi, lo) pair the next layer's MMAs consume -- but hands the split tile to the next layer through tensor memory.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace dccn {

constexpr int kChainMaxStages = 3;
constexpr int kChainSlots = 4;        // TMEM operand slots of one k-block each: [128 rows x 64 K] as 32 hi + 32 lo columns

struct ChainStage {
  int src;          // >= 0: A operand from HBM through tmA[src] (raw fp32, split by the splitter warps); -1: from the slots
  int nkb;          // k-blocks (64 K elements each) this stage contracts
  int slot0;        // slot of k-block 0 (k-block j reads slot0 + j): staging of the HBM operand or the previous stage's output
  int nsub;         // 128-wide n-subtiles (weight rows sub * 128 ...); intermediate stages have exactly one
  int dst_slot0;    // >= 0: the output [128 x 128] goes to slots dst_slot0, dst_slot0 + 1 as fp16 hi/lo; -1: HBM through `epi`
  float w_scale_inv;   // 1 / (power-of-two weight scale of the fp16 planes), GemmLayer::w_scale_inv
  const float* bias;   // [N padded to 128] fp32 (intermediate stages; the last stage's bias is epi.bias)
  const unsigned* amax_in;   // HBM operand: max |A| recorded by its producer (operand scale, see gemm_tc.cuh), or nullptr
};

struct alignas(64) ChainParams {
  CUtensorMap tmA[2];                       // HBM A operands: fp32 [M, K], box 128 x 32
  CUtensorMap tmW[kChainMaxStages][2];      // weights: fp16 hi / lo [N, K], box 128 x 64
  EpiStore epi;                             // last stage: bias / activation / amax / bulk tensor store (+ aux copy)
  ChainStage st[kChainMaxStages];
  int nst;
  int M;
  int small_first;
  long long* trace;   // measurement aid (tools/chain_trace.py): CTA 0 records clock64() at its hand-off points, or nullptr
};

int launch_chain(const ChainParams& p, cudaStream_t s, int num_sms);

}  // namespace dccn
