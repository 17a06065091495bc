// gemm_tc.cuh -- hand-written sm_100a GEMM: TMA -> shared memory -> tcgen05.mma (kind::tf32)
// -> TMEM accumulators -> tcgen05.ld -> fp32 register accumulation -> fused epilogue.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )       A, B row-major with K contiguous ("K-major")
//
// * persistent: grid = min(#tiles, #SMs); CTA loops over 128 x BN output tiles, n fastest so
//   that concurrently running CTAs share the A tile in L2;
// * warp roles: warp 0 = TMA producer (1 lane), warp 1 = TMEM allocator + MMA issuer
//   (1 lane), [SPLIT: warps 2..5 = A-operand tf32 splitters,] then 4*CG epilogue warps, each
//   owning one TMEM lane quarter and one of CG column groups of the tile;
// * three pipelines: smem full/empty ring (TMA <-> MMA), a 2-deep TMEM accumulator ring
//   (MMA <-> epilogue), static tile schedule;
// * operands are 128-byte-swizzled [rows x 32 fp32] boxes written by TMA and consumed through
//   K-major SWIZZLE_128B shared-memory descriptors; out-of-bounds rows / the K tail are
//   zero-filled by TMA, so ragged M, N, K need no special code in the main loop;
// * SPLIT = true is the 3xTF32 scheme of DCCN_PREC_PARITY: every operand is a (hi, lo) pair of
//   tf32-exact values and each k-step issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  into the same
//   TMEM accumulator (the lo*lo term, <= 2^-24 relative, is dropped).  B (weights) is split once
//   on the host into two planes.  A (activations) stays ONE fp32 plane in HBM/L2: TMA lands the
//   fp32 tile in shared memory and four "splitter" warps rewrite it in place as hi and write lo
//   to a second buffer (same byte offsets, so the 128B swizzle is preserved), then
//   fence.proxy.async + mbarrier hand the stage to the MMA warp.  This halves activation
//   traffic and footprint compared with storing hi/lo planes;
// * PAIR = true (needs ATM) runs 2-CTA clusters with cta_group::2 MMAs: one instruction spans
//   both SMs of the pair (M = 256 x N = BN); each CTA keeps ITS 128 rows of A (in its TMEM) and
//   its 128 accumulator rows, but only HALF of the weight tile (BN/2 rows) in its shared memory.
//   The measured limit of the big layers is the per-SM ingress rate (~32-36 B/clk per SM from
//   L2, multicast does not help because the bytes still arrive); the pair form halves the
//   weight bytes each SM receives per flop.  Only the leader CTA (cluster rank 0) issues MMAs;
//   the peer's splitter / epilogue warps signal the leader's `ready` / `tempty` barriers through
//   mapa + mbarrier.arrive.shared::cluster, and the leader's tcgen05.commit multicasts to the
//   `empty` / `tfull` barriers of both CTAs;
// * K-CHUNKED ACCUMULATION: the tensor core adds into its fp32 accumulator with truncation,
//   so a long dependent chain (K = 896..1024 -> hundreds of MMAs) accumulates a systematic
//   bias ~10x above fp32 round-to-nearest (measured on the shipped 16-QAM checkpoint).  The
//   MMA warp therefore closes an accumulator every `kc` k-blocks; the epilogue warps drain it
//   with tcgen05.ld and add the partial sums in registers with IEEE round-to-nearest while
//   the tensor core already fills the other accumulator.  kc = 0 keeps the whole K in TMEM
//   (DCCN_PREC_FAST, where operand rounding dominates anyway).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "epilogue.cuh"

namespace dccn {

// ---- optional in-kernel timeline (compile with -DDCCN_TRACE; tools/trace_gemm.py) ---------------------------
// CTA 0 records clock64() at the hand-off points of every role: buf[role * 4096 + i].
#ifdef DCCN_TRACE
__device__ long long* g_trace_buf = nullptr;
__device__ int g_abl_dev = 0;   // ablation mask (timing experiments only; results are garbage when != 0)
struct TraceCtr { int n[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; };
#define DCCN_ABL(bit) ((abl_ & (bit)) != 0)
#define DCCN_ABL_DECL const int abl_ = g_abl_dev
#define DCCN_TRACE_DECL TraceCtr trc_; long long* const trc_buf_ = blockIdx.x == 0 ? g_trace_buf : nullptr
#define DCCN_TRACE_EV(role)                                                                     \
  do {                                                                                          \
    if (trc_buf_ && trc_.n[role] < 4096) trc_buf_[(role) * 4096 + trc_.n[role]++] = clock64();  \
  } while (0)
#else
#define DCCN_TRACE_DECL
#define DCCN_TRACE_EV(role)
#define DCCN_ABL(bit) false
#define DCCN_ABL_DECL
#endif

constexpr int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

constexpr int imin(int a, int b) { return a < b ? a : b; }

// ATM ("A in tensor memory", SPLIT only): the splitter warps write the hi / lo tf32 images of
// the A tile into TMEM with tcgen05.st instead of back into shared memory, and the MMAs use
// the TS form (A from TMEM, B from shared memory).  Shared memory then carries only the raw
// fp32 A tile (written once by TMA, read once by the splitters) and the B planes: the SS form
// at N = 128 needs 8 KB of operand reads per 64-cycle MMA = all of the 128 B/clk shared-memory
// bandwidth, which left nothing for TMA writes and the split traffic.
//
// DEC ("decoupled A ring", ATM without PAIR): the raw A tiles get their own small shared-memory
// ring with their own producer warp.  An A slot is free again as soon as the splitters have read
// it (long before the MMAs that use it retire), so A tiles are fetched far ahead and the splitters
// already hold the next tile in registers when a TMEM staging slot frees up: the per-stage
// dependency loop no longer contains "TMA latency of A + split", only "commit -> tcgen05.st".
//
// F16 (DEC only; `DCCN_F16X3=1`, STAGED -- written without a GPU at hand, not yet run or measured): the (hi, lo) pairs
// are fp16 instead of tf32 (same 11-bit significands, tools/acc_split_emul.py) and the MMAs are kind::f16, which
// contracts K = 16 per instruction at the instruction rate of kind::tf32's K = 8.  A k-block is then 64 K-elements: the
// weight planes are [BN x 64 fp16] boxes (128-byte rows, same shared-memory descriptors and the same 32-byte advance
// per k-step), the raw fp32 A tile arrives as TWO [128 x 32 fp32] boxes (two A-ring slots), and the splitters pack the
// 64 values of a row into 32 + 32 TMEM columns (two fp16 per column) -- the same staging footprint as the tf32 form.
// The instruction count per k-block is unchanged (12), the K it covers doubles.
template <int BN, bool SPLIT, int CG, bool ATM = false, bool DEC = false, bool F16 = false>
struct TcCfg {
  static_assert(!ATM || SPLIT, "A-in-TMEM is the parity (3xTF32) configuration");
  static_assert(!DEC || ATM, "the decoupled A ring feeds the TMEM staging");
  static_assert(!F16 || DEC, "the fp16 hi/lo form is built on the decoupled A-in-TMEM pipeline");
  static constexpr int BM = 128;
  static constexpr int BK = 32;                       // 32 fp32 = one 128-byte swizzle row
  static constexpr int KB = F16 ? 64 : 32;            // K elements one k-block contracts
  static constexpr int A_BOXES = F16 ? 2 : 1;         // [128 x 32 fp32] boxes of raw A per k-block
  static constexpr int UMMA_K = 8;                    // 32 bytes of K per instruction: 8 tf32 / 8 TMEM columns of 2 fp16
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int PLANES = SPLIT ? 2 : 1;
  static constexpr int STAGE_BYTES = (DEC ? 0 : (ATM ? A_BYTES : PLANES * A_BYTES)) + PLANES * B_BYTES;
  static constexpr int B_OFF = DEC ? 0 : (ATM ? A_BYTES : PLANES * A_BYTES);   // offset of B_hi inside a stage
#ifndef DCCN_TC_SA
#define DCCN_TC_SA 4             // experiment knob: slots of the raw-A ring (2 = one fp16-form k-block)
#endif
  static constexpr int SA = DEC ? DCCN_TC_SA : 0;                  // slots of the separate raw-A ring
  static constexpr int A_RING_BYTES = SA * A_BYTES;
  static constexpr int A_TMEM_COLS = ATM ? 64 : 0;                 // per stage: 32 hi + 32 lo columns
  static constexpr int TX_BYTES = A_BYTES + PLANES * B_BYTES;   // bytes TMA delivers per stage
  static constexpr int SPLIT_WARPS = SPLIT ? 4 : 0;
  // REGBAL (DEC with 8 epilogue warps): 16 warps in four role-pure warpgroups -- 0..3 = producer, MMA issuer, A-ring producer,
  // (idle); 4..7 = splitters; 8..15 = epilogue -- so that setmaxnreg can move registers to the epilogue warps
#ifndef DCCN_TC_REGBAL
#define DCCN_TC_REGBAL 1      // ON in mode 5 (only the control warpgroup gives registers away), see below
#endif
  static constexpr bool REGBAL = DCCN_TC_REGBAL && DEC && CG == 2;
#ifndef DCCN_TC_REGS_EPI
#define DCCN_TC_REGS_EPI 160
#endif
#ifndef DCCN_TC_REGS_CTRL
#define DCCN_TC_REGS_CTRL 56
#endif
  static constexpr int REGS_CTRL = DCCN_TC_REGS_CTRL, REGS_SPLIT = 104, REGS_EPI = DCCN_TC_REGS_EPI;   // mode 5: 128 * (56 + 128 + 2 * 160) = 64 512
  static constexpr int APROD_WARP = REGBAL ? 2 : 2 + SPLIT_WARPS;  // DEC: producer warp of the A ring
  static constexpr int EPI_WARP0 = REGBAL ? 8 : 2 + SPLIT_WARPS + (DEC ? 1 : 0);
  static constexpr int PATCH_KB = (DCCN_TC_PATCH_KB == 8 && BN <= 128) ? 8 : 4;   // per epilogue warp: one or two 4 KB patches
  static constexpr int PATCH_BYTES = 4 * CG * 1024 * PATCH_KB;     // store-transpose patches of the epilogue warps
  static constexpr int SMEM_BUDGET = 225 * 1024 - PATCH_BYTES;     // operand rings (227 KB minus patches, alignment, barriers)
  static constexpr int STAGES_RAW = (SMEM_BUDGET - A_RING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES_TM = ATM ? (512 - 2 * BN) / 64 : 8;
#ifndef DCCN_TC_STAGE_CAP
#define DCCN_TC_STAGE_CAP 3      // 4 stages of the fp16 form = 226 KB of shared memory = NO L1 data cache left (228 KB unified): bias loads, the
                                 // phase equaliser's f rows and register spills of the epilogue warps then pay L2 latency (measured: 3.31 -> 2.99 ms per pass)
#endif
  static constexpr int STAGES = imin(imin(STAGES_RAW, STAGES_TM), DCCN_TC_STAGE_CAP);
  static constexpr int A_TMEM_COL0 = 2 * BN;                       // A staging columns follow the accumulators
  static constexpr int TMEM_COLS = pow2_at_least(2 * BN + STAGES * A_TMEM_COLS);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + A_RING_BYTES + PATCH_BYTES + 1024 /*align*/ + 512 /*barriers*/;
  static constexpr int THREADS = 32 * EPI_WARP0 + 128 * CG;
  static constexpr int COLS_PER_GROUP = BN / CG;
  static constexpr int NCH = COLS_PER_GROUP / 32;     // 32-column register chunks per epilogue thread
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(BN % (32 * CG) == 0, "each column group is a whole number of 32-column chunks");
  static_assert(NCH >= 1 && NCH <= 4, "at most 128 accumulator registers per epilogue thread");
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(TMEM_COLS <= 512, "TMEM has 512 columns");
};

// Per-tile K range.  Default: the whole K.  ksplit > 1: split-K -- tile index gains a slice coordinate, slice z
// contracts k-blocks [z*per, (z+1)*per) and its rows are stored at row offset z * m_tiles * 128 (the caller sums the
// slices; used by the weight-gradient GEMMs whose K is the batch).  band >= 0: block-banded B operand -- n-tile j
// only contracts k-blocks [j*BN/32 - band, (j+1)*BN/32 + band) (the (S,K) 'same' conv as a Toeplitz matrix has no
// taps between OFDM symbols more than (S-1)/2 apart, so those 128x128 blocks are structurally zero).
struct KSched {
  int ksplit = 1;
  int band = -1;
};

struct TileK {
  int m_blk, n_blk, kslice, kb0, kb1;
};

template <int BN, int KBE = 32>
DCCN_DEVINL TileK tile_decode(int tile, int m_tiles, int n_tiles, int num_kb, int ksplit, int band) {
  TileK t;
  if (ksplit <= 1) {
    t.kslice = 0;
    t.m_blk = tile / n_tiles;
    t.n_blk = tile - t.m_blk * n_tiles;
    t.kb0 = 0;
    t.kb1 = num_kb;
    if (band >= 0) {
      constexpr int KBN = BN / KBE;
      const int lo = t.n_blk * KBN - band, hi = (t.n_blk + 1) * KBN + band;
      t.kb0 = lo > 0 ? lo : 0;
      t.kb1 = hi < num_kb ? hi : num_kb;
    }
    return t;
  }
  const int mn = m_tiles * n_tiles;
  t.kslice = tile / mn;
  const int t2 = tile - t.kslice * mn;
  t.m_blk = t2 / n_tiles;
  t.n_blk = t2 - t.m_blk * n_tiles;
  const int per = (num_kb + ksplit - 1) / ksplit;
  t.kb0 = t.kslice * per;
  t.kb1 = t.kb0 + per < num_kb ? t.kb0 + per : num_kb;
  return t;
}

struct TcOperands {
  CUtensorMap a0;       // A, one fp32 plane
  CUtensorMap b0, b1;   // B hi / lo (b1 unused when !SPLIT)
};

template <int BN, bool SPLIT, int CG, bool ATM, bool PAIR, class Epi, bool F16 = false>
__global__ void __launch_bounds__(TcCfg<BN, SPLIT, CG, ATM, (ATM && !PAIR), F16>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               int M, int N, int K, int kc, int ksplit, int band, float a_scale, float out_scale,
               const unsigned* __restrict__ amax_in, int small_first, const __grid_constant__ Epi epi) {
  constexpr bool DEC = ATM && !PAIR;
  using C = TcCfg<BN, SPLIT, CG, ATM, DEC, F16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem + C::STAGES * C::STAGE_BYTES;                       // DEC: SA raw fp32 A tiles
  uint8_t* patches = a_ring + C::A_RING_BYTES;                               // 4 KB per epilogue warp
  uint64_t* full = reinterpret_cast<uint64_t*>(patches + C::PATCH_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* ready = empty + C::STAGES;   // [STAGES] A tile split into hi/lo (SPLIT only)
  uint64_t* tfull = ready + C::STAGES;   // [2] accumulator (one K chunk) ready for the epilogue
  uint64_t* tempty = tfull + 2;          // [2] accumulator drained
  uint64_t* fullA = tempty + 2;          // [4] DEC: raw A tile landed
  uint64_t* emptyA = fullA + 4;          // [4] DEC: raw A tile consumed by the splitters
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(emptyA + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifndef DCCN_TC_REGBAL_MODE
#define DCCN_TC_REGBAL_MODE 5    // 5 = control warpgroup 128 -> 56, splitters untouched, epilogue warps 128 -> 160 (the deterministic one)
#endif
  if constexpr (C::REGBAL && DCCN_TC_REGBAL_MODE == 4) {   // experiment: rebalance before anything else happens
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_CTRL));
    else if (warp < 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_SPLIT));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_EPI));
  }
  DCCN_TRACE_DECL;
  DCCN_ABL_DECL;
  if constexpr (F16) {
    // Operand scale of the activations: a power of two that puts max |A| (recorded by the producing kernel, see
    // amax_update_warp) into [2^13, 2^14) -- far from fp16's 65 504, and as far from its subnormals as the range allows.
    // Exact (power of two), undone together with the weight scale by `out_scale` in the tile epilogue.
    if (amax_in) {
      const unsigned bits = __ldg(amax_in);
      const int e = (int)(bits >> 23) - 127;                       // amax in [2^e, 2^(e+1))
      if (bits != 0u && e > -100 && e < 100) {                     // 0, denormal-ish, inf / NaN: leave the operands alone
        a_scale = __uint_as_float((uint32_t)(13 - e + 127) << 23);
        out_scale *= __uint_as_float((uint32_t)(e - 13 + 127) << 23);
      }
    }
  }
  const int n_tiles = (N + BN - 1) / BN;
  static_assert(!PAIR || ATM, "the CTA-pair form is built on the A-in-TMEM configuration");
  // PAIR: "tile" below is a pair tile (two consecutive M-tiles x one N-tile); CTA rank r of the
  // cluster owns M-tile 2*m_pair + r.  Both CTAs of a pair run the same tile sequence.
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int m_tiles = (M + C::BM - 1) / C::BM;
  const int num_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles * ksplit) * n_tiles;   // PAIR: full-K tiles only
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int BROWS = PAIR ? BN / 2 : BN;            // weight rows this CTA holds
  constexpr int BH = BROWS * C::BK * 4;                // bytes of one weight plane in a stage
  constexpr int TX = (DEC ? 0 : C::A_BYTES) + C::PLANES * BH;   // bytes TMA lands in THIS CTA per (B) stage
  const int num_kb = (K + C::KB - 1) / C::KB;
  const int kb_per_chunk = (kc <= 0 || kc > num_kb) ? num_kb : kc;
  // (m_blk, n_blk, k-block range) of a tile index
  auto decode = [=](int tile) -> TileK {
    if constexpr (PAIR) return TileK{(tile / n_tiles) * 2 + crank, tile % n_tiles, 0, 0, num_kb};
    else return tile_decode<BN, C::KB>(tile, m_tiles, n_tiles, num_kb, ksplit, F16 && band > 0 ? band / 2 : band);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (SPLIT) tma_prefetch_desc(&tmB1);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < C::STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
        mbar_init(&ready[s], PAIR ? 8 : 4);   // one elected lane of each splitter warp (of both CTAs)
      }
      for (int a = 0; a < 4; ++a) {
        mbar_init(&fullA[a], 1);
        mbar_init(&emptyA[a], 4);   // one elected lane of each splitter warp
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], (PAIR ? 2 : 1) * 4 * CG);   // one elected lane per epilogue warp (of both CTAs)
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_pair(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (PAIR) cluster_sync_all();             // both CTAs' barriers / TMEM exist before anything crosses over
  const uint32_t tmem_base = *tmem_ptr;
  // experiment (ABL 256): only lane 0 polls the mbarrier, the rest of the warp parks at __syncwarp
  auto wait_ = [&](uint64_t* bar, uint32_t par) {
    if (DCCN_ABL(256)) {
      if (lane == 0) mbar_wait(bar, par);
      __syncwarp();
    } else if (DCCN_ABL(512)) {
      while (!mbar_try_wait(bar, par)) __nanosleep(32);
    } else {
      mbar_wait(bar, par);
    }
  };

  auto role_producer = [&]() {
    // =============================== TMA producer ===============================
    // The WHOLE warp runs the loop so that addresses / coordinates stay warp-uniform (uniform
    // registers feed UTMALDG directly); one elected lane issues.  A single-lane branch makes the
    // compiler wrap every TMA / MMA instruction in an R2UR + elect loop (~100 clk per issue).
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const TileK tk = decode(tile);
        const int m_blk = tk.m_blk, n_blk = tk.n_blk;
        for (int kb = tk.kb0; kb < tk.kb1; ++kb) {
          wait_(&empty[stage], phase ^ 1);
          if (DCCN_ABL(8)) {
            if (elect_one()) mbar_arrive(&full[stage]);
          } else if (elect_one()) {
            mbar_expect_tx(&full[stage], TX);
            uint8_t* st = smem + stage * C::STAGE_BYTES;
            if (!DEC) tma_load_2d(st, &tmA0, &full[stage], kb * C::BK, m_blk * C::BM);
            // PAIR: only my half of the weight tile (rows crank*BN/2 ...)
            const int brow = n_blk * BN + crank * BROWS;
            tma_load_2d(st + C::B_OFF, &tmB0, &full[stage], kb * C::KB, brow);
            if (SPLIT) tma_load_2d(st + C::B_OFF + BH, &tmB1, &full[stage], kb * C::KB, brow);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  };
  auto role_mma = [&]() {
    // =============================== MMA issuer =================================
    if (crank == 0) {   // whole warp, warp-uniform control flow; one elected lane issues
      constexpr uint32_t idesc = F16 ? umma_idesc_f16(BN, 128) : umma_idesc_tf32(BN, PAIR ? 256 : 128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const TileK tk = decode(tile);
        for (int kb0 = tk.kb0; kb0 < tk.kb1; kb0 += kb_per_chunk) {
          const int kb1 = kb0 + kb_per_chunk < tk.kb1 ? kb0 + kb_per_chunk : tk.kb1;
          if (lane == 0) DCCN_TRACE_EV(2);
          if (PAIR) {
            mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
          }
          const uint32_t d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = kb0; kb < kb1; ++kb) {
            if (PAIR) {
              mbar_wait_cluster(&ready[stage], phase);
            } else if (DCCN_ABL(1024)) {   // experiment: the old serial waits
              if (kb == kb0) mbar_wait(&tempty[acc], acc_phase ^ 1);
              mbar_wait(SPLIT ? &ready[stage] : &full[stage], phase);
              if (DEC) mbar_wait(&full[stage], phase);
            } else {
              // accumulator drained (first k-block of a chunk) + operands of this stage ready (+ DEC: weight planes
              // landed): probed together -- the three round trips in series left the tensor pipe's short queue empty
              mbar_wait_multi(SPLIT ? &ready[stage] : &full[stage], phase, DEC ? &full[stage] : nullptr, phase,
                              kb == kb0 ? &tempty[acc] : nullptr, acc_phase ^ 1);
            }
            if (lane == 0) DCCN_TRACE_EV(5);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t a_lo = a_hi + C::A_BYTES;
            const uint32_t b_hi = a_hi + C::B_OFF;
            const uint32_t b_lo = b_hi + BH;
            const uint32_t ta_hi = tmem_base + (uint32_t)(C::A_TMEM_COL0 + stage * C::A_TMEM_COLS);
            if (elect_one()) {
#pragma unroll
            for (int k = 0; k < (DCCN_ABL(32) ? 0 : C::BK / C::UMMA_K); ++k) {
              const uint32_t koff = k * C::UMMA_K * 4;   // byte advance inside the 128 B swizzle row
              const uint64_t da_hi = umma_desc_sw128(a_hi + koff);
              const uint64_t db_hi = umma_desc_sw128(b_hi + koff);
              const uint32_t accum = (kb != kb0 || k != 0) ? 1u : 0u;
              if constexpr (PAIR) {
                const uint64_t db_lo = umma_desc_sw128(b_lo + koff);
                const uint32_t ka = ta_hi + (uint32_t)(k * C::UMMA_K);
                umma_tf32_ts_pair(d, ka + 32, db_hi, idesc, accum);         // A_lo * B_hi   (M = 256, both SMs)
                umma_tf32_ts_pair(d, ka, db_lo, idesc, 1u);                 // A_hi * B_lo
                umma_tf32_ts_pair(d, ka, db_hi, idesc, 1u);                 // A_hi * B_hi
              } else if constexpr (ATM) {
                const uint64_t db_lo = umma_desc_sw128(b_lo + koff);
                const uint32_t ka = ta_hi + (uint32_t)(k * C::UMMA_K);      // 8 columns per k-step (8 tf32 / 16 fp16)
                if constexpr (F16) {
                  umma_f16_ts(d, ka + 32, db_hi, idesc, accum);             // A_lo * B_hi
                  umma_f16_ts(d, ka, db_lo, idesc, 1u);                     // A_hi * B_lo
                  if (!small_first) umma_f16_ts(d, ka, db_hi, idesc, 1u);   // A_hi * B_hi
                } else {
                  umma_tf32_ts(d, ka + 32, db_hi, idesc, accum);            // A_lo * B_hi
                  umma_tf32_ts(d, ka, db_lo, idesc, 1u);                    // A_hi * B_lo
                  if (!small_first) umma_tf32_ts(d, ka, db_hi, idesc, 1u);  // A_hi * B_hi
                }
              } else if (SPLIT) {
                const uint64_t da_lo = umma_desc_sw128(a_lo + koff);
                const uint64_t db_lo = umma_desc_sw128(b_lo + koff);
                umma_tf32(d, da_lo, db_hi, idesc, accum);   // small terms first
                umma_tf32(d, da_hi, db_lo, idesc, 1u);
                umma_tf32(d, da_hi, db_hi, idesc, 1u);
              } else {
                umma_tf32(d, da_hi, db_hi, idesc, accum);
              }
            }
            if constexpr (ATM && !PAIR) {
              // small_first: the eight cross terms of the k-block (2^-11 of the result) went into the accumulator while
              // it was still small; the four hi*hi MMAs follow.  The tensor core adds into its fp32 accumulator with
              // truncation at the magnitude of the running sum, so the k-block now pays 4 truncations at full magnitude
              // instead of 12 (measured, profiles/accuracy_r2.txt) -- same instructions, same operands, other order.
              if (small_first) {
#pragma unroll
                for (int k = 0; k < (DCCN_ABL(32) ? 0 : C::BK / C::UMMA_K); ++k) {
                  const uint64_t db_hi = umma_desc_sw128(b_hi + k * C::UMMA_K * 4);
                  const uint32_t ka = ta_hi + (uint32_t)(k * C::UMMA_K);
                  if constexpr (F16) umma_f16_ts(d, ka, db_hi, idesc, 1u);
                  else umma_tf32_ts(d, ka, db_hi, idesc, 1u);
                }
              }
            }
            DCCN_TRACE_EV(9);
            if (PAIR) umma_commit_pair(&empty[stage], 0x3);   // slot released in both CTAs of the pair
            else umma_commit(&empty[stage]);                  // smem slot reusable once these MMAs retire
            // K chunk complete -> epilogue(s) drain it.  The commit above already says so when the chunk is short
            // (2 * kc <= STAGES, so the stage barrier of the chunk's last k-block cannot complete another phase before
            // the epilogue has seen this one): the epilogue warps wait on empty[stage of the last k-block] themselves
            // and the second commit (~90 clk of tensor-pipe time, measured) is dropped.
            if (kb + 1 == kb1 && (PAIR || 2 * kb_per_chunk > C::STAGES)) {
              if (PAIR) umma_commit_pair(&tfull[acc], 0x3);
              else umma_commit(&tfull[acc]);
            }
            }   // elect_one
            __syncwarp();
            if (lane == 0) DCCN_TRACE_EV(8);
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  };
  auto role_aprod = [&]() {
    // =============================== A-ring producer (DEC) ======================
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const TileK tk = decode(tile);
      const int m_blk = tk.m_blk;
      for (int kb = tk.kb0; kb < tk.kb1; ++kb) {
#pragma unroll
        for (int bx = 0; bx < C::A_BOXES; ++bx) {
          wait_(&emptyA[sa], pa ^ 1);
          if (DCCN_ABL(16)) {
            if (elect_one()) mbar_arrive(&fullA[sa]);
          } else if (elect_one()) {
            mbar_expect_tx(&fullA[sa], C::A_BYTES);
            tma_load_2d(a_ring + sa * C::A_BYTES, &tmA0, &fullA[sa], kb * C::KB + bx * C::BK, m_blk * C::BM);
          }
          __syncwarp();
          if (++sa == C::SA) {
            sa = 0;
            pa ^= 1;
          }
        }
      }
    }
  };
  auto role_split_dec = [&]() {
    // =============================== A-operand splitters (DEC) ==================
    // read the raw tile early (frees the A slot at once), keep hi/lo in registers, and write
    // them to the TMEM staging slot the moment the MMAs that used it have retired.
    int stage = 0, sa = 0;
    uint32_t phase = 0, pa = 0;
    const int r = (warp & 3) * 32 + lane;
    const f32x2 a_scale2 = pack2(a_scale, a_scale);
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const TileK tk = decode(tile);
      for (int kb = tk.kb0; kb < tk.kb1; ++kb) {
        float hi[32], lo[32];
        if constexpr (F16) {
          // two raw boxes per k-block; box bx fills packed columns [16 bx, 16 bx + 16) of hi / lo
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            wait_(&fullA[sa], pa);
            const uint32_t rowp = smem_u32(a_ring + sa * C::A_BYTES + r * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              if (DCCN_ABL(1)) {
                hi[16 * bx + 2 * c] = hi[16 * bx + 2 * c + 1] = lo[16 * bx + 2 * c] = lo[16 * bx + 2 * c + 1] = 0.f;
                continue;
              }
              const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));   // undo the 128B swizzle: chunk c of row r
              f16_split_pack_scaled(v.x, v.y, a_scale2, hi[16 * bx + 2 * c], lo[16 * bx + 2 * c]);
              f16_split_pack_scaled(v.z, v.w, a_scale2, hi[16 * bx + 2 * c + 1], lo[16 * bx + 2 * c + 1]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) asm volatile("" : "+f"(hi[16 * bx + c]));   // loads done before the release (below)
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyA[sa]);                // raw box consumed
            if (++sa == C::SA) {
              sa = 0;
              pa ^= 1;
            }
          }
        } else {
        wait_(&fullA[sa], pa);
        const uint32_t rowp = smem_u32(a_ring + sa * C::A_BYTES + r * 128);
        if (DCCN_ABL(1)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) hi[c] = lo[c] = 0.f;
        } else if (DCCN_ABL(64)) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));
            tf32_split_cvt(v.x, hi[4 * c + 0], lo[4 * c + 0]);
            tf32_split_cvt(v.y, hi[4 * c + 1], lo[4 * c + 1]);
            tf32_split_cvt(v.z, hi[4 * c + 2], lo[4 * c + 2]);
            tf32_split_cvt(v.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
        } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));   // undo the 128B swizzle: chunk c of row r
          tf32_split(v.x, hi[4 * c + 0], lo[4 * c + 0]);
          tf32_split(v.y, hi[4 * c + 1], lo[4 * c + 1]);
          tf32_split(v.z, hi[4 * c + 2], lo[4 * c + 2]);
          tf32_split(v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        }
        // The release below must not overtake the loads: an mbarrier arrive does not wait for ld.shared still queued
        // in the LSU (observed: with the epilogue's global stores backing the LSU up at a tile boundary, TMA refilled
        // the slot before the queued loads had read it).  Pinning hi[] -- a function of every loaded word -- in front
        // of the arrive makes the loads complete first.
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("" : "+f"(hi[c]));
        __syncwarp();
        if (lane == 0) mbar_arrive(&emptyA[sa]);                  // raw tile consumed
        if (++sa == C::SA) {
          sa = 0;
          pa ^= 1;
        }
        }
        wait_(&empty[stage], phase ^ 1);                      // TMEM staging slot is free again
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) +
                            (uint32_t)(C::A_TMEM_COL0 + stage * C::A_TMEM_COLS);
        if (!DCCN_ABL(2)) {
          tmem_st_32x32(ta, hi);
          tmem_st_32x32(ta + 32, lo);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[stage]);
        if (warp == 2 && lane == 0) DCCN_TRACE_EV(4);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  };
  auto role_split_smem = [&]() {
    // =============================== A-operand splitters ========================
    // fp32 tile (as landed by TMA, swizzled) -> hi in place, lo at the same offsets of the
    // second buffer.  Elementwise, so the swizzle pattern needs no decoding.
    const int t = threadIdx.x - 64;            // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const TileK tk = decode(tile);
      for (int kb = tk.kb0; kb < tk.kb1; ++kb) {
        wait_(&full[stage], phase);
        if constexpr (ATM) {
          // thread <-> A row (TMEM lane).  Undo the TMA 128B swizzle while reading: the 16-byte
          // chunk c of row r lives at chunk position c ^ (r & 7).
          const int r = (warp & 3) * 32 + lane;
          const uint32_t rowp = smem_u32(smem + stage * C::STAGE_BYTES + r * 128);
          float hi[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));
            tf32_split(v.x, hi[4 * c + 0], lo[4 * c + 0]);
            tf32_split(v.y, hi[4 * c + 1], lo[4 * c + 1]);
            tf32_split(v.z, hi[4 * c + 2], lo[4 * c + 2]);
            tf32_split(v.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) +
                              (uint32_t)(C::A_TMEM_COL0 + stage * C::A_TMEM_COLS);
          tmem_st_32x32(ta, hi);
          tmem_st_32x32(ta + 32, lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&ready[stage]), 0));   // leader's barrier
            else mbar_arrive(&ready[stage]);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        float4* a_hi = reinterpret_cast<float4*>(smem + stage * C::STAGE_BYTES);
        float4* a_lo = reinterpret_cast<float4*>(smem + stage * C::STAGE_BYTES + C::A_BYTES);
#pragma unroll
        for (int i = 0; i < C::A_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;
          const float4 v = a_hi[idx];
          float4 h, l;
          tf32_split(v.x, h.x, l.x);
          tf32_split(v.y, h.y, l.y);
          tf32_split(v.z, h.z, l.z);
          tf32_split(v.w, h.w, l.w);
          a_hi[idx] = h;
          a_lo[idx] = l;
        }
        fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[stage]);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  };
  auto role_epilogue = [&]() {
    // =============================== epilogue warps =============================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int cg = (warp - C::EPI_WARP0) >> 2; // column group
    typename Epi::State st;
    if constexpr (Epi::kWarpStore && C::PATCH_KB == 8) st.tog = 2u;   // two store patches per warp (store_block_tma)
    int acc = 0;
    uint32_t acc_phase = 0;
    int estage = 0;                            // kc = 1: the smem-stage ring position of the chunk being drained
    uint32_t ephase = 0;
    const bool stage_signal = !PAIR && 2 * kb_per_chunk <= C::STAGES;
    // plain-store epilogue: the accumulator columns live as fp32 PAIRS (FFMA2: the per-chunk adds take half the issue slots);
    // the arithmetic-heavy fused epilogues keep the scalar form (the pair form cost them registers: 1.5 KB of spills)
    constexpr bool kPairs = std::is_same<Epi, EpiStore>::value;
    f32x2 r[kPairs ? C::NCH : 1][16];
    float rf[kPairs ? 1 : C::NCH][32];
    const f32x2 one2 = pack2(1.f, 1.f);
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const TileK tk = decode(tile);
      const int n_blk = tk.n_blk;
      const int row_base = (tk.kslice * m_tiles + tk.m_blk) * C::BM + q * 32;   // split-K slices stack along the rows
      const int row = row_base + lane;
      const int num_chunks = (tk.kb1 - tk.kb0 + kb_per_chunk - 1) / kb_per_chunk;
      if constexpr (Epi::kPrefetch) epi.prefetch(row_base, lane, n_blk * BN + cg * C::COLS_PER_GROUP, C::COLS_PER_GROUP);
      for (int ch = 0; ch < num_chunks; ++ch) {
        if (stage_signal) {
          // (estage, ephase) = ring position of the chunk's FIRST k-block; wait for its last one, then step over the chunk
          const int len = (ch + 1 == num_chunks) ? (tk.kb1 - tk.kb0) - ch * kb_per_chunk : kb_per_chunk;
          int ls = estage + len - 1;
          uint32_t lp = ephase;
          if (ls >= C::STAGES) {
            ls -= C::STAGES;
            lp ^= 1;
          }
          wait_(&empty[ls], lp);
          estage += len;
          if (estage >= C::STAGES) {
            estage -= C::STAGES;
            ephase ^= 1;
          }
        } else {
          wait_(&tfull[acc], acc_phase);
        }
        if (warp == C::EPI_WARP0 && lane == 0) DCCN_TRACE_EV(6);
        tc_fence_after();
        const uint32_t t0 =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cg * C::COLS_PER_GROUP);
        if constexpr (kPairs) {
          if (ch == 0) {                           // the tile's sum starts at +0 (one code path for every chunk below)
#pragma unroll
            for (int j = 0; j < C::NCH; ++j)
#pragma unroll
              for (int i = 0; i < 16; ++i) r[j][i] = 0ull;
          }
#pragma unroll
          for (int j = 0; j < C::NCH; ++j) {
            float v[32];
            tmem_ld_32x32(t0 + j * 32, v);
            // fp32 round-to-nearest adds of the K chunks, two per issue slot (v * 1 + r: one rounding, the value of the add)
#pragma unroll
            for (int i = 0; i < 16; ++i) r[j][i] = fma2(pack2(v[2 * i], v[2 * i + 1]), one2, r[j][i]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < C::NCH; ++j) {
            float v[32];
            if (DCCN_ABL(4)) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 1.f;
            } else tmem_ld_32x32(t0 + j * 32, v);
            if (ch == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) rf[j][i] = v[i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) rf[j][i] = __fadd_rn(rf[j][i], v[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // leader's barrier
          else mbar_arrive(&tempty[acc]);
        }
        if (warp == C::EPI_WARP0 && lane == 0) DCCN_TRACE_EV(7);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (warp == C::EPI_WARP0 && lane == 0) DCCN_TRACE_EV(0);
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {
        if (warp == C::EPI_WARP0 && lane == 0 && j > 0) DCCN_TRACE_EV(1);
        const int col = n_blk * BN + cg * C::COLS_PER_GROUP + j * 32;
        if constexpr (kPairs) {
          float y[32];
          if constexpr (F16) {
            const f32x2 os2 = pack2(out_scale, out_scale);
#pragma unroll
            for (int i = 0; i < 16; ++i) unpack2(mul2(r[j][i], os2), y[2 * i], y[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) unpack2(r[j][i], y[2 * i], y[2 * i + 1]);
          }
          epi.run_warp(st, row_base, lane, col, y, smem_u32(patches + (warp - C::EPI_WARP0) * 1024 * C::PATCH_KB));
        } else {
          if constexpr (F16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rf[j][i] *= out_scale;
          }
          if constexpr (Epi::kWarpStore)
            epi.run_warp(st, row_base, lane, col, rf[j], smem_u32(patches + (warp - C::EPI_WARP0) * 1024 * C::PATCH_KB));
          else
            epi.template run<32>(st, row, col, rf[j]);
        }
      }
      if (warp == C::EPI_WARP0 && lane == 0) DCCN_TRACE_EV(1);
    }
    epi.flush(st);
  };
  // No load may still be in flight when a warp gives registers away (setmaxnreg.dec below): the prologue's amax load is issued
  // by every warp, also by those whose role never reads the scale; its late write-back would land in a register that by then
  // belongs to an epilogue warp (seen as run-to-run differences in ~0.03 % of the outputs, 20 x more without an L1).
  asm volatile("" ::"f"(a_scale), "f"(out_scale), "r"(tmem_base) : "memory");
  if constexpr (C::REGBAL) {
    // warpgroup-aligned roles + setmaxnreg: the epilogue warps (64 accumulators + 32 loaded values + epilogue arithmetic) spilled
    // at the 128 registers a 16-warp CTA gets per thread; the producer-side warpgroups hand theirs over (csrc/chain.cu)
    // DCCN_TC_REGBAL_MODE (experiments on the non-determinism): 1 = dec + inc, 2 = role layout only (no setmaxnreg),
    // 3 = dec only (epilogue warps stay at 128)
    if (warp < 4) {
      if (DCCN_TC_REGBAL_MODE != 2 && DCCN_TC_REGBAL_MODE != 4 && DCCN_TC_REGBAL_MODE != 6) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_CTRL));
      if (warp == 0) role_producer();
      else if (warp == 1) role_mma();
      else if (warp == C::APROD_WARP) role_aprod();
    } else if (warp < 8) {
      if (DCCN_TC_REGBAL_MODE != 2 && DCCN_TC_REGBAL_MODE != 4 && DCCN_TC_REGBAL_MODE != 5) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_SPLIT));
      role_split_dec();
    } else {
      if (DCCN_TC_REGBAL_MODE == 1 || DCCN_TC_REGBAL_MODE == 5 || DCCN_TC_REGBAL_MODE == 6) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_EPI));   // 5: only the control warpgroup gives registers away, 6: only the splitters
      role_epilogue();
    }
  } else {
    if (warp == 0) role_producer();
    else if (warp == 1) role_mma();
    else if (DEC && warp == C::APROD_WARP) role_aprod();
    else if (DEC && warp < C::APROD_WARP) role_split_dec();
    else if (SPLIT && warp < C::EPI_WARP0) role_split_smem();
    else role_epilogue();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();             // no CTA leaves while its peer can still signal it
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, bool SPLIT, int CG, bool ATM, bool PAIR, class Epi, bool F16 = false>
inline int launch_gemm_tc(const TcOperands& op, int M, int N, int K, int kc, const Epi& epi, cudaStream_t s,
                          int num_sms, KSched ks = KSched(), float a_scale = 1.f, float out_scale = 1.f,
                          const unsigned* amax_in = nullptr, int small_first = 0) {
  using C = TcCfg<BN, SPLIT, CG, ATM, (ATM && !PAIR), F16>;
  if (M <= 0) return 0;
  auto kern = gemm_tc_kernel<BN, SPLIT, CG, ATM, PAIR, Epi, F16>;
  static bool attr_set = false;
  if (!attr_set) {
    DCCN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (M + C::BM - 1) / C::BM, n_tiles = (N + BN - 1) / BN;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (PAIR) {
    const int pair_tiles = ((m_tiles + 1) / 2) * n_tiles;
    const int pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
    cfg.gridDim = dim3(2 * pairs);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int num_kb = (K + C::KB - 1) / C::KB;
    if (ks.ksplit > 1) {   // every slice must own at least one k-block
      const int per = (num_kb + ks.ksplit - 1) / ks.ksplit;
      ks.ksplit = (num_kb + per - 1) / per;
      ks.band = -1;
    }
    if (ks.ksplit < 1) ks.ksplit = 1;
    const int tiles = m_tiles * n_tiles * ks.ksplit;
    cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
  }
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = s;
  if (PAIR) ks = KSched();
  DCCN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, op.a0, op.b0, SPLIT ? op.b1 : op.b0, M, N, K, kc, ks.ksplit, ks.band, a_scale,
                                  out_scale, amax_in, small_first, epi));
  return 0;
}

}  // namespace dccn
