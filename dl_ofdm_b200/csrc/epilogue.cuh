// epilogue.cuh -- fused GEMM epilogues shared by the tcgen05 and the SIMT main loops.
//
// Contract: the main loop hands each thread NC consecutive accumulator columns of ONE
// output row (NC even, col0 even, so an (I,Q) pair never straddles two calls):
//     epi.run<NC>(row, col0, v);     ... per chunk
//     epi.flush();                   ... once per thread at kernel end
// Activations are stored as one fp32 plane (p1 == nullptr) or as a tf32 hi/lo pair
// of planes (DCCN_PREC_PARITY) that the next layer's TMA loads feed straight to the
// tensor core.
#pragma once
#include "common.cuh"

namespace dccn {

// destination of an activation matrix [rows, ld] (+ column offset) in 1 or 2 planes
struct ActOut {
  float* p0;
  float* p1;   // nullptr => single full-precision plane
  int ld;
  int col_off;
};

template <int NC>
DCCN_DEVINL void store_act(const ActOut& o, int row, int col, const float (&y)[NC]) {
  float* d0 = o.p0 + (size_t)row * o.ld + o.col_off + col;
  if (o.p1 == nullptr) {
    if constexpr (NC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < NC; i += 4) *reinterpret_cast<float4*>(d0 + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < NC; i += 2) *reinterpret_cast<float2*>(d0 + i) = make_float2(y[i], y[i + 1]);
    }
  } else {
    float* d1 = o.p1 + (size_t)row * o.ld + o.col_off + col;
    float hi[NC], lo[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) tf32_split(y[i], hi[i], lo[i]);
    if constexpr (NC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < NC; i += 4) {
        *reinterpret_cast<float4*>(d0 + i) = make_float4(hi[i], hi[i + 1], hi[i + 2], hi[i + 3]);
        *reinterpret_cast<float4*>(d1 + i) = make_float4(lo[i], lo[i + 1], lo[i + 2], lo[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; i += 2) {
        *reinterpret_cast<float2*>(d0 + i) = make_float2(hi[i], hi[i + 1]);
        *reinterpret_cast<float2*>(d1 + i) = make_float2(lo[i], lo[i + 1]);
      }
    }
  }
}

// Warp-cooperative store of a [32 rows x 32 cols] fp32 block whose rows are held one per lane
// (the tcgen05.ld layout): transposed through a 4 KB shared-memory patch (16-byte chunks XOR-
// swizzled by row, conflict-free both ways) so that every global store instruction writes four
// complete 128-byte lines instead of 32 scattered 16-byte pieces.
DCCN_DEVINL void store_block_warp(float* base, int ld, int row0, int M, int lane, const float (&y)[32],
                                  uint32_t patch /* shared address of this warp's 4 KB patch */) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t a = patch + (uint32_t)((lane * 8 + (c ^ (lane & 7))) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(y[4 * c]), "f"(y[4 * c + 1]),
                 "f"(y[4 * c + 2]), "f"(y[4 * c + 3])
                 : "memory");
  }
  __syncwarp();
  const int cc = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    const float4 v = lds128(patch + (uint32_t)((rr * 8 + (cc ^ (rr & 7))) << 4));
    if (row0 + rr < M) *reinterpret_cast<float4*>(base + (size_t)(row0 + rr) * ld + cc * 4) = v;
  }
  __syncwarp();
}

// Same block, handed to the TMA engine instead: the lanes write their rows into the patch in the 128-byte-swizzle
// pattern (16-byte chunk c of row r at chunk position c ^ (r & 7) -- conflict-free, and exactly what a SWIZZLE_128B
// tensor map undoes), one lane issues a bulk tensor store per destination and the warp moves on.  The warp no longer
// waits for global-store credits (measured: ~3.5 k clk per block with st.global while all 148 SMs drain their tiles at
// once, during which the K-chunk drains of the next tile were stalled); rows >= M and columns >= N are clipped by
// the tensor map.  `tm1` is an optional second destination of the same block (API side outputs).
// DCCN_TC_PATCH_KB = 8: a warp owns TWO 4 KB patches and alternates between them (`tog`), so a block only waits for the
// store issued two blocks ago.  Why: a 4 KB store queues in the TMA unit behind whatever loads the producer warps have in
// flight (~160 KB in the ingress-bound layers = ~4 k clk), and with one patch every block of the tile epilogue sat out that
// latency before it could overwrite the patch.
#ifndef DCCN_TC_PATCH_KB
#define DCCN_TC_PATCH_KB 4
#endif
DCCN_DEVINL void store_block_tma(const CUtensorMap* tm0, int col0_0, const CUtensorMap* tm1, int col0_1, int row0,
                                 int lane, const float (&y)[32], uint32_t patch, unsigned& tog) {
  if (tog & 2u) {                              // bit 1: the kernel gave this warp two patches; bit 0: which one is next
    patch += (tog & 1u) * 4096u;
    tog ^= 1u;
    if (lane == 0) tma_store_wait_read1();     // the store before last (same patch) has been read
  } else {
    if (lane == 0) tma_store_wait_read();      // the previous block has left the patch
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t a = patch + (uint32_t)((lane * 8 + (c ^ (lane & 7))) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(y[4 * c]), "f"(y[4 * c + 1]),
                 "f"(y[4 * c + 2]), "f"(y[4 * c + 3])
                 : "memory");
  }
  fence_proxy_async();                         // generic-proxy writes -> visible to the TMA engine
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm0, patch, col0_0, row0);
    if (tm1) tma_store_2d(tm1, patch, col0_1, row0);
    tma_store_commit();
  }
}

// max |y| of a produced activation, kept per buffer as the bit pattern of a non-negative float (unsigned order = float
// order).  The fp16 hi/lo GEMM that consumes the buffer derives its power-of-two operand scale from it (gemm_tc.cuh,
// `amax_in`), so that no trained weight set can push an operand past fp16's 65 504.  One redux + one RED per 32x32 block.
template <int NV>
DCCN_DEVINL void amax_update_warp(unsigned* amax, const float (&y)[NV], bool row_ok, unsigned& seen) {
  unsigned m = 0u;
#pragma unroll
  for (int i = 0; i < NV; ++i) m = max(m, __float_as_uint(y[i]) & 0x7FFFFFFFu);
  if (!row_ok) m = 0u;
  m = __reduce_max_sync(0xffffffffu, m);
  // `seen` = the largest value this warp has already reported (warp-uniform): every block of every CTA firing a RED at
  // the SAME address serialises in one L2 slice (~1e5 same-address atomics per layer); after a warp's first few blocks
  // its running maximum rarely moves, so almost all of them are dropped here.
  if (m > seen) {
    seen = m;
    if ((threadIdx.x & 31) == 0) atomicMax(amax, m);
  }
}

// -------------------------------------------------------------------------------------
// y = act(acc + bias)  ->  activation planes      (tf.layers.dense / packed complex layers)
// -------------------------------------------------------------------------------------
struct EpiStore {
  const float* bias;   // [N] (may be nullptr)
  ActOut out;
  float* aux;          // optional extra full-precision copy [rows, aux_ld] (API outputs), or nullptr
  int aux_ld;
  int act;             // 0 = linear, 1 = tanh
  int M, N;
  unsigned* amax = nullptr;   // optional: running max |y| of the destination buffer (see amax_update_warp)
  CUtensorMap tm_out;  // tensor-core path: [M, N] view of out.p0 + out.col_off, box 32 x 32 (set by run_gemm)
  CUtensorMap tm_aux;  // same for aux
  struct State {
    unsigned amax_seen = 0u;   // see amax_update_warp
    unsigned tog = 0u;         // store patches of this warp: bit 1 = two of them, bit 0 = which is next (store_block_tma)
  };

  template <int NC>
  DCCN_DEVINL void run(State&, int row, int col0, float (&v)[NC]) const {
    if (row >= M || col0 >= N) return;
    float y[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      float t = v[i] + (bias ? __ldg(bias + col0 + i) : 0.f);
      y[i] = (act == 1) ? tanhf(t) : t;
    }
    store_act<NC>(out, row, col0, y);
    if (aux) {
      float* d = aux + (size_t)row * aux_ld + col0;
#pragma unroll
      for (int i = 0; i < NC; i += 2) *reinterpret_cast<float2*>(d + i) = make_float2(y[i], y[i + 1]);
    }
  }
  // tensor-core path: the 32 lanes of a warp hold 32 consecutive rows (row0 + lane)
  static constexpr bool kWarpStore = true;
  static constexpr bool kPrefetch = false;
  DCCN_DEVINL void run_warp(State& st, int row0, int lane, int col0, float (&v)[32], uint32_t patch) const {
    if (col0 >= N || row0 >= M) return;          // warp-uniform
    // All bias loads first (8 independent 16-byte broadcasts; the bias array is padded to a multiple of 128 floats),
    // then the arithmetic, with the activation switch outside the loop: with the per-element `act ? tanhf : id`
    // branch inside the loop the compiler serialised load -> wait -> tanh per column (~8 k clk per 32x32 block).
    if (bias) {
      const float4* bp = reinterpret_cast<const float4*>(bias + col0);
      float4 b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = __ldg(bp + i);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[4 * i + 0] += b[i].x;
        v[4 * i + 1] += b[i].y;
        v[4 * i + 2] += b[i].z;
        v[4 * i + 3] += b[i].w;
      }
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
    }
    if (amax) {
      // columns >= N of a ragged tile hold bias-padding zeros + accumulated zeros (TMA zero-fills the weight rows)
      amax_update_warp<32>(amax, v, row0 + lane < M, st.amax_seen);
    }
    store_block_tma(&tm_out, col0, aux ? &tm_aux : nullptr, col0, row0, lane, v, patch, st.tog);
  }
  DCCN_DEVINL void flush(State&) const { tma_store_wait_read(); }   // shared memory outlives the last bulk store
};

// -------------------------------------------------------------------------------------
// Phase-only equaliser fused behind the (S,K) 'same' complex conv (dev/py/model.py:426-437):
//   chest = acc + bias                       (complex, IQ interleaved)
//   eq    = f * conj(chest) / |chest|        (no epsilon, like the reference)
//   corr  = eq * conj(eq)                    (real part only is non-zero)
// f = inputs_complex (output of the learned-DFT layer), read back as hi+lo.
// -------------------------------------------------------------------------------------
template <bool SYM>
struct EpiPhaseEqT {
  const float* bias;    // [N] packed (b0-b1, b1-b0) pattern
  const float* f0;      // inputs_complex plane 0 [M, ld_f]
  const float* f1;      // plane 1 (lo) or nullptr
  int ld_f;
  ActOut eq;            // [M, N]      IQ interleaved
  ActOut corr;          // [M, N/2]    real part only (imag is exactly 0 and is dropped from the GEMM)
  float* chest_out;     // optional fp32 [M, N] ('chest' fetch), or nullptr
  int M, N;
  // SYM (folded schedule): eq column c of symbol s = c / sym_cols lands at  s * sym_stride + c % sym_cols  and the
  // real corr value of complex point j at  s * sym_stride + j % (sym_cols / 2)  (+ the ActOut column offsets), i.e.
  // both interleave per symbol in one [M*S, 3K] operand.  !SYM: plain [M, N] / [M, N/2] matrices.
  int sym_cols = 0, sym_stride = 0;
  int act = 0;          // 1: chest = tanh(acc + bias) (ablation graphs whose chest is a tanh dense, model.py:775-779)
  unsigned* amax_eq = nullptr;     // optional running max |eq| / max corr of the destination buffers (amax_update_warp)
  unsigned* amax_corr = nullptr;
  CUtensorMap tm_eq;    // tensor-core path: view of eq.p0 + eq.col_off, box 32 x 32 (set by run_gemm)
  CUtensorMap tm_chest; // same for chest_out
  CUtensorMap tm_corr;  // view of corr.p0 + corr.col_off, box 32 x 32
  struct State {
    float c_even[16];   // corr values of the even 32-column chunk, kept until the odd chunk completes a 32 x 32 block
    unsigned seen_eq = 0u, seen_corr = 0u;   // see amax_update_warp
    unsigned tog = 0u;                       // store patches of this warp (store_block_tma)
  };
  static constexpr bool kWarpStore = true;
  static constexpr bool kPrefetch = true;
  int pf = 1;           // pull the tile's f rows into L2 while the tile is still accumulating (DCCN_EPI_PREFETCH=0 disables)
  // The tail below reads f with per-lane row loads that nothing else has touched since the learned-DFT layer wrote it
  // ~1 ms earlier: from DRAM that latency sits between the last K chunk and the first store of every tile, while the
  // MMA warp has only two accumulators of head start on the next tile.
  DCCN_DEVINL void prefetch(int row0, int lane, int col0, int cols) const {
    const int row = row0 + lane;
    if (!pf || row >= M || col0 >= N) return;
    const float* p = f0 + (size_t)row * ld_f + col0;
    for (int c = 0; c < cols && col0 + c < N; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + c));
  }
  DCCN_DEVINL int eq_col(int c) const {
    if constexpr (SYM) return (c / sym_cols) * sym_stride + c % sym_cols;
    else return c;
  }
  DCCN_DEVINL int corr_col(int j) const {
    if constexpr (SYM) return (j / (sym_cols >> 1)) * sym_stride + j % (sym_cols >> 1);
    else return j;
  }

  // tensor-core path: rows row0 + lane; eq goes out through the coalescing transpose
  DCCN_DEVINL void run_warp(State& st, int row0, int lane, int col0, float (&v)[32], uint32_t patch) const {
    if (col0 >= N || row0 >= M) return;          // warp-uniform
    const int row = row0 + lane;
    const bool ok = row < M;
    const float4* fp0 = reinterpret_cast<const float4*>(f0 + (size_t)(ok ? row : 0) * ld_f + col0);   // ld_f % 4 == 0
    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
    float4 b4[8], f4[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b4[i] = __ldg(bp + i);   // all loads in flight before the arithmetic
#ifndef DCCN_PE_ABL
#define DCCN_PE_ABL 0      // timing experiments only (results are garbage when != 0): 1 no f loads, 2 no corr store, 4 no amax, 8 no eq store
#endif
#pragma unroll
    for (int i = 0; i < 8; ++i) f4[i] = (DCCN_PE_ABL & 1) ? make_float4(1.f, 2.f, 3.f, 4.f) : fp0[i];
    float e[32], c[16];
    // the activation switch stays OUTSIDE the arithmetic loop (a per-element runtime branch makes the compiler serialise
    // load -> wait -> tanh per column, the EpiStore finding)
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float4 bq = b4[i >> 2];
      v[i] += (i & 2) ? bq.z : bq.x;
      v[i + 1] += (i & 2) ? bq.w : bq.y;
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float4 fq = f4[i >> 2];
      const float cr = v[i], ci = v[i + 1];
      const float2 f = (i & 2) ? make_float2(fq.z, fq.w) : make_float2(fq.x, fq.y);
      const float inv = rsqrtf(cr * cr + ci * ci);
      const float nr = cr * inv, ni = (-ci) * inv;
      const float er = f.x * nr - f.y * ni, ei = f.x * ni + f.y * nr;
      e[i] = er;
      e[i + 1] = ei;
      c[i / 2] = er * er + ei * ei;
    }
    if (amax_eq && !(DCCN_PE_ABL & 4)) amax_update_warp<32>(amax_eq, e, ok, st.seen_eq);
    if (amax_corr && corr.p0 && !(DCCN_PE_ABL & 4)) amax_update_warp<16>(amax_corr, c, ok, st.seen_corr);
    if (!(DCCN_PE_ABL & 8)) store_block_tma(&tm_eq, eq_col(col0), nullptr, 0, row0, lane, e, patch, st.tog);
    if (chest_out) store_block_tma(&tm_chest, col0, nullptr, 0, row0, lane, v, patch, st.tog);
    if (corr.p0 && !(DCCN_PE_ABL & 2)) {
      // corr is half as wide as eq: the 16 values of an even 32-column chunk wait in registers for the 16 of the odd
      // chunk (a warp owns both: its column group is 64 wide), then leave as ONE 32 x 32 block through the TMA engine --
      // 64-byte per-thread st.global pieces from the epilogue warps backed the LSU up (the round-1 finding for eq itself)
      if (((col0 >> 5) & 1) == 0 && col0 + 32 < N) {
#pragma unroll
        for (int i = 0; i < 16; ++i) st.c_even[i] = c[i];
      } else if ((col0 >> 5) & 1) {
        float blk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          blk[i] = st.c_even[i];
          blk[16 + i] = c[i];
        }
        store_block_tma(&tm_corr, corr_col((col0 - 32) / 2), nullptr, 0, row0, lane, blk, patch, st.tog);
      } else if (ok) {
        store_act<16>(corr, row, corr_col(col0 / 2), c);      // lone even chunk at the ragged edge
      }
    }
  }

  template <int NC>
  DCCN_DEVINL void run(State&, int row, int col0, float (&v)[NC]) const {
    if (row >= M || col0 >= N) return;
    const float* fp0 = f0 + (size_t)row * ld_f + col0;
    const float* fp1 = f1 ? f1 + (size_t)row * ld_f + col0 : nullptr;
    float e[NC], c[NC / 2];
#pragma unroll
    for (int i = 0; i < NC; i += 2) {
      float cr = v[i] + __ldg(bias + col0 + i);
      float ci = v[i + 1] + __ldg(bias + col0 + i + 1);
      if (act == 1) {
        cr = tanhf(cr);
        ci = tanhf(ci);
      }
      float2 f = *reinterpret_cast<const float2*>(fp0 + i);
      if (fp1) {
        float2 fl = *reinterpret_cast<const float2*>(fp1 + i);
        f.x += fl.x;
        f.y += fl.y;
      }
      // conj(c)/|c| (model.py:431-433): one rsqrt instead of hypot + two IEEE divides (differs from
      // the reference's op sequence by ~2e-7 relative; |c| = 0 gives NaN exactly like the reference)
      const float inv = rsqrtf(cr * cr + ci * ci);
      float nr = cr * inv;
      float ni = (-ci) * inv;
      float er = f.x * nr - f.y * ni;            // complex multiply    model.py:434
      float ei = f.x * ni + f.y * nr;
      e[i] = er;
      e[i + 1] = ei;
      c[i / 2] = er * er + ei * ei;              // eq*conj(eq): real = er^2+ei^2, imag = 0
      v[i] = cr;
      v[i + 1] = ci;
    }
    store_act<NC>(eq, row, eq_col(col0), e);
    if (corr.p0) store_act<NC / 2>(corr, row, corr_col(col0 / 2), c);
    if (chest_out) {
      float* d = chest_out + (size_t)row * N + col0;
#pragma unroll
      for (int i = 0; i < NC; i += 2) *reinterpret_cast<float2*>(d + i) = make_float2(v[i], v[i + 1]);
    }
  }
  DCCN_DEVINL void flush(State&) const { tma_store_wait_read(); }
};

typedef EpiPhaseEqT<false> EpiPhaseEq;
typedef EpiPhaseEqT<true> EpiPhaseEqSym;

// -------------------------------------------------------------------------------------
// Demodulation head fused behind the 896->2D dense (dev/py/model.py:1275-1291) + the BER
// head (dev/py/ofdmreceiver_np.py:154-169).  Per data subcarrier d with (I,Q)=out_iq:
//   h   = leaky( [conv2d_1(] conv2d(I,Q) [)] )            2 -> 2^nb (-> 2^nb for the v1 head)
//   lg  = leaky( dense_1( concat(h, I, Q) ) )              -> 2*nb
//   p   = softmax over each bit's pair; hard = argmax (first index wins ties)
//   conf[truth][hard]++ ; ce += logsumexp(p) - p[truth]    (softmax-xent ON the softmax, Q5)
// -------------------------------------------------------------------------------------
struct HeadWeights {
  float Wc[2][16];
  float bc[16];
  float Wc1[16][16];
  float bc1[16];
  float W1[18][8];
  float b1[8];
};

template <int NB, bool V1>
struct EpiHead {
  static constexpr int MO = 1 << NB;
  const float* bias;          // [N = 2*D]
  HeadWeights hw;
  const uint8_t* bits;        // [M, D, NB] or nullptr
  float* soft;                // [M, D, NB, 2] or nullptr
  uint8_t* hard;              // [M, D, NB] or nullptr
  unsigned long long* conf;   // [4] or nullptr
  double* ce_sum;             // [1] or nullptr
  int M, N;
  static constexpr bool kWarpStore = false;
  static constexpr bool kPrefetch = false;
  struct State {   // per-thread accumulators (flushed once per thread); a thread sees far fewer than 2^32 decisions
    unsigned int n = 0;                       // decisions counted
    unsigned int c01 = 0, c10 = 0, c11 = 0;   // truth / decision pairs (c00 = n - the rest)
    float ce = 0.f;
  };

  // The head of ONE data subcarrier: (I, Q) = out_iq (bias already added) -> probabilities p[2*NB] (p0, p1 per bit) and
  // the hard decisions as bit k of the return value.
  DCCN_DEVINL unsigned subcarrier(float I, float Q, float (&p)[2 * NB]) const {
    // The two small layers run on packed fp32 pairs (FFMA2: the kernel was bound by its instruction count): outputs m, m + 1
    // of the 1x1 conv(s) and the two logits of a bit are adjacent words of HeadWeights, i.e. one uniform-register pair.
    // Summation order: bias first, then the inputs in index order, (I, Q) last -- plain fp32 FMAs like before.
    const f32x2 I2 = pack2(I, I), Q2 = pack2(Q, Q), leak = pack2(0.2f, 0.2f);
    float h[MO];
#pragma unroll
    for (int m = 0; m < MO; m += 2) {
      f32x2 a = fma2(I2, pack2(hw.Wc[0][m], hw.Wc[0][m + 1]), pack2(hw.bc[m], hw.bc[m + 1]));
      a = fma2(Q2, pack2(hw.Wc[1][m], hw.Wc[1][m + 1]), a);
      unpack2(a, h[m], h[m + 1]);
    }
    if constexpr (V1) {
      float h2[MO];
#pragma unroll
      for (int n = 0; n < MO; n += 2) {
        f32x2 a = pack2(hw.bc1[n], hw.bc1[n + 1]);
#pragma unroll
        for (int m = 0; m < MO; ++m) a = fma2(pack2(h[m], h[m]), pack2(hw.Wc1[m][n], hw.Wc1[m][n + 1]), a);
        unpack2(a, h2[n], h2[n + 1]);
      }
#pragma unroll
      for (int m = 0; m < MO; ++m) h[m] = h2[m];
    }
#pragma unroll
    for (int m = 0; m < MO; m += 2) {
      float t0, t1;
      unpack2(mul2(pack2(h[m], h[m + 1]), leak), t0, t1);
      h[m] = fmaxf(t0, h[m]);
      h[m + 1] = fmaxf(t1, h[m + 1]);
    }
    unsigned hbits = 0u;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      f32x2 a = pack2(hw.b1[2 * k], hw.b1[2 * k + 1]);
#pragma unroll
      for (int m = 0; m < MO; ++m) a = fma2(pack2(h[m], h[m]), pack2(hw.W1[m][2 * k], hw.W1[m][2 * k + 1]), a);
      a = fma2(I2, pack2(hw.W1[MO][2 * k], hw.W1[MO][2 * k + 1]), a);
      a = fma2(Q2, pack2(hw.W1[MO + 1][2 * k], hw.W1[MO + 1][2 * k + 1]), a);
      float l0, l1, t0, t1;
      unpack2(a, l0, l1);
      unpack2(mul2(a, leak), t0, t1);
      l0 = fmaxf(t0, l0);
      l1 = fmaxf(t1, l1);
      // softmax over the pair: exp(l - max) / sum  ==  {1, t} / (1 + t),  t = exp(-|l1 - l0|) in (0, 1].
      // ex2.approx on -|d| * log2(e) is within 2 ulp of exp() for |d| < 1 and its absolute error only shrinks beyond;
      // 1 / s from rcp.approx (1 ulp on (1, 2]), t / s one more rounding: |dp| <= 2e-7, far inside the 1e-5 budget
      // (the libm expf + two IEEE divides this replaces were a quarter of the kernel's instruction count)
      const float t = __expf(-fabsf(l1 - l0));
      float pb;                                                // probability of the larger / smaller logit
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pb) : "f"(1.0f + t));   // 1 + t in (1, 2]: within 1 ulp, no slow path
      const float ps = t * pb;
      const bool one_big = l1 > l0;
      const float p0 = one_big ? ps : pb, p1 = one_big ? pb : ps;
      p[2 * k] = p0;
      p[2 * k + 1] = p1;
      hbits |= (p1 > p0 ? 1u : 0u) << k;                       // tf.argmax: first index on ties
    }
    return hbits;
  }

  // scalar form (the fused GEMM-epilogue variant, run<NC>, unrolls 16 subcarriers: the packed-pair form made of inline asm
  // sent the compiler front end into a >25 min optimisation there)
  DCCN_DEVINL unsigned subcarrier_scalar(float I, float Q, float (&p)[2 * NB]) const {
    float h[MO];
#pragma unroll
    for (int m = 0; m < MO; ++m) h[m] = I * hw.Wc[0][m] + Q * hw.Wc[1][m] + hw.bc[m];
    if constexpr (V1) {
      float h2[MO];
#pragma unroll
      for (int n = 0; n < MO; ++n) {
        float a = hw.bc1[n];
#pragma unroll
        for (int m = 0; m < MO; ++m) a += h[m] * hw.Wc1[m][n];
        h2[n] = a;
      }
#pragma unroll
      for (int m = 0; m < MO; ++m) h[m] = h2[m];
    }
#pragma unroll
    for (int m = 0; m < MO; ++m) h[m] = fmaxf(0.2f * h[m], h[m]);
    unsigned hbits = 0u;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      float l0 = hw.b1[2 * k], l1 = hw.b1[2 * k + 1];
#pragma unroll
      for (int m = 0; m < MO; ++m) {
        l0 += h[m] * hw.W1[m][2 * k];
        l1 += h[m] * hw.W1[m][2 * k + 1];
      }
      l0 += I * hw.W1[MO][2 * k] + Q * hw.W1[MO + 1][2 * k];
      l1 += I * hw.W1[MO][2 * k + 1] + Q * hw.W1[MO + 1][2 * k + 1];
      l0 = fmaxf(0.2f * l0, l0);
      l1 = fmaxf(0.2f * l1, l1);
      // softmax over the pair: exp(l - max) / sum  ==  {1, t} / (1 + t),  t = exp(-|l1 - l0|) in (0, 1].
      // ex2.approx on -|d| * log2(e) is within 2 ulp of exp() for |d| < 1 and its absolute error only shrinks beyond;
      // 1 / s is the correctly rounded reciprocal, t / s one more rounding: |dp| <= 1.5e-7, far inside the 1e-5 budget
      // (the libm expf + two IEEE divides this replaces were a quarter of the kernel's instruction count)
      const float t = __expf(-fabsf(l1 - l0));
      const float pb = __frcp_rn(1.0f + t), ps = t * pb;       // probability of the larger / smaller logit
      const bool one_big = l1 > l0;
      const float p0 = one_big ? ps : pb, p1 = one_big ? pb : ps;
      p[2 * k] = p0;
      p[2 * k + 1] = p1;
      hbits |= (p1 > p0 ? 1u : 0u) << k;                       // tf.argmax: first index on ties
    }
    return hbits;
  }

  // labels y (bit k = label of bit k), decisions hbits, probabilities p -> confusion counts + double-softmax CE
  DCCN_DEVINL void account(State& st, unsigned y, unsigned hbits, const float (&p)[2 * NB]) const {
    constexpr unsigned MASK = (1u << NB) - 1u;
    st.n += NB;
    st.c11 += __popc(y & hbits);
    st.c10 += __popc(y & ~hbits & MASK);
    st.c01 += __popc(~y & hbits & MASK);
    const unsigned wrong = y ^ hbits;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      // softmax-xent applied ON the softmax outputs (ofdmreceiver_np.py:155-159):  logsumexp(p0, p1) - p_y
      //   = max(p) + log(1 + exp(-|p1 - p0|)) - p_y = log(1 + exp(-|d|)) + (decision == label ? 0 : |d|)     (monitor only)
      const float ad = fabsf(p[2 * k + 1] - p[2 * k]);
      const float lg = __logf(1.0f + __expf(-ad));
      st.ce += ((wrong >> k) & 1u) ? lg + ad : lg;
    }
  }

  template <int NC>
  DCCN_DEVINL void run(State& st, int row, int col0, float (&v)[NC]) const {
    if (row >= M || col0 >= N) return;
    const int D = N >> 1;
    constexpr int ND = NC / 2;                 // data subcarriers in this call
    constexpr int NBYTES = ND * NB;            // label / hard-decision bytes in this call
    // Labels and hard decisions of the ND subcarriers are NBYTES contiguous bytes.  With the
    // 32-column chunks of the tensor-core path they are 16-byte aligned -> vector access.
    constexpr bool VEC = (NBYTES % 16 == 0);
    constexpr int NW = (NBYTES + 3) / 4;
    uint32_t yw[NW], hw_out[NW];
    const size_t o0 = ((size_t)row * D + (col0 >> 1)) * NB;
#pragma unroll
    for (int w = 0; w < NW; ++w) hw_out[w] = 0u;
    if (bits) {
      if constexpr (VEC) {
#pragma unroll
        for (int w = 0; w < NW; w += 4) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(bits + o0) + (w >> 2));
          yw[w] = t.x; yw[w + 1] = t.y; yw[w + 2] = t.z; yw[w + 3] = t.w;
        }
      } else if constexpr (NBYTES == 4) {
        yw[0] = __ldg(reinterpret_cast<const uint32_t*>(bits + o0));
      } else if constexpr (NBYTES == 2) {
        yw[0] = __ldg(reinterpret_cast<const uint16_t*>(bits + o0));
      } else {
#pragma unroll
        for (int w = 0; w < NW; ++w) yw[w] = 0u;
#pragma unroll
        for (int b = 0; b < NBYTES; ++b) yw[b >> 2] |= (uint32_t)bits[o0 + b] << ((b & 3) * 8);
      }
    }
#pragma unroll
    for (int i = 0; i < NC; i += 2) {
      const float I = v[i] + (bias ? __ldg(bias + col0 + i) : 0.f);
      const float Q = v[i + 1] + (bias ? __ldg(bias + col0 + i + 1) : 0.f);
      float p[2 * NB];
      const unsigned hbits = subcarrier_scalar(I, Q, p);
      unsigned y = 0u;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int bidx = (i >> 1) * NB + k;              // byte index inside this call
        hw_out[bidx >> 2] |= ((hbits >> k) & 1u) << ((bidx & 3) * 8);
        if (bits) y |= ((yw[bidx >> 2] >> ((bidx & 3) * 8)) & 1u) << k;
      }
      if (soft) {
        float* sp = soft + (o0 + (size_t)(i >> 1) * NB) * 2;
        if constexpr ((2 * NB) % 4 == 0) {
#pragma unroll
          for (int q = 0; q < 2 * NB; q += 4) *reinterpret_cast<float4*>(sp + q) = make_float4(p[q], p[q + 1], p[q + 2], p[q + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 2 * NB; q += 2) *reinterpret_cast<float2*>(sp + q) = make_float2(p[q], p[q + 1]);
        }
      }
      if (bits) account(st, y, hbits, p);
    }
    if (hard) {
      if constexpr (VEC) {
#pragma unroll
        for (int w = 0; w < NW; w += 4)
          *(reinterpret_cast<uint4*>(hard + o0) + (w >> 2)) = make_uint4(hw_out[w], hw_out[w + 1], hw_out[w + 2], hw_out[w + 3]);
      } else if constexpr (NBYTES == 4) {
        *reinterpret_cast<uint32_t*>(hard + o0) = hw_out[0];
      } else if constexpr (NBYTES == 2) {
        *reinterpret_cast<uint16_t*>(hard + o0) = (uint16_t)hw_out[0];
      } else {
#pragma unroll
        for (int b = 0; b < NBYTES; ++b) hard[o0 + b] = (uint8_t)((hw_out[b >> 2] >> ((b & 3) * 8)) & 0xFFu);
      }
    }
  }
  DCCN_DEVINL void flush(State& st) const {
    if (!bits) return;
    unsigned a = __reduce_add_sync(0xffffffffu, st.n - st.c01 - st.c10 - st.c11);
    unsigned b = __reduce_add_sync(0xffffffffu, st.c01);
    unsigned c = __reduce_add_sync(0xffffffffu, st.c10);
    unsigned d = __reduce_add_sync(0xffffffffu, st.c11);
    double e = (double)st.ce;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if ((threadIdx.x & 31) == 0) {
      if (conf) {
        if (a) atomicAdd(conf + 0, (unsigned long long)a);
        if (b) atomicAdd(conf + 1, (unsigned long long)b);
        if (c) atomicAdd(conf + 2, (unsigned long long)c);
        if (d) atomicAdd(conf + 3, (unsigned long long)d);
      }
      if (ce_sum) atomicAdd(ce_sum, e);
    }
  }
};

}  // namespace dccn
