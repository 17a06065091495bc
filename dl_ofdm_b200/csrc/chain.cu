// chain.cu -- chained per-symbol GEMMs on tcgen05 (sm_100a): up to three layers of equalizer_ofdm's per-symbol runs
// (dev/py/model.py:370-379, :437-462) in ONE persistent kernel; the intermediate [128 x 128] tiles go from the epilogue
// warps' registers straight into the tensor-memory operand slots of the next layer instead of through HBM.
//
// Built on the decoupled A-in-TMEM fp16 hi/lo pipeline of gemm_tc.cuh (same roles, same operand formats, same MMA order):
//   warp 0      weight producer: [128 x 64] fp16 hi / lo boxes of the current (stage, n-subtile, k-block) into a 4-deep ring
//   warp 1      TMEM allocator + MMA issuer: per k-block 12 kind::f16 TS-form MMAs (cross terms first) into one of two
//               128-column accumulators, ONE tcgen05.commit per k-block (kc = 1: the accumulator is drained every k-block and
//               the partial sums are added in fp32 round-to-nearest by the epilogue warps -- the tensor core's own
//               accumulator truncates)
//   warp 2      producer of the raw-A ring (4 x [128 x 32 fp32])
//   warps 4..7  splitters: raw fp32 A boxes of the HBM-fed stages -> fp16 (hi, lo) -> tcgen05.st into an operand slot
//   warps 8..15 epilogue: drain every k-block, add; at the end of a stage either (intermediate) add the bias, pick a
//               power-of-two scale per ROW and k-block (exact; recorded in shared memory and undone when the next stage's
//               partial sums are drained), split to fp16 hi/lo and tcgen05.st the result into the operand slots, or
//               (last stage) EpiStore: bias / amax / swizzled patch / bulk tensor store.
// Tensor memory (512 columns): accumulators at 0 and 128, four operand slots of 64 columns (32 hi + 32 lo: two fp16 per
// column) at 256.  A slot holds one k-block of A, whether it was staged from HBM or produced by the previous stage; the
// staging of a stage's HBM operand may alias that stage's own output slots because the output is only written after the
// stage's last k-block has been drained (= all of its MMAs have retired).
#include "chain.cuh"

namespace dccn {

namespace {

constexpr int CH_BM = 128, CH_BN = 128, CH_KB = 64;
// weight ring: 3 stages, not 4 -- with 4 the kernel held 226 KB of shared memory, which leaves NO L1 data cache (228 KB
// unified): every bias load and every register spill of the epilogue warps then pays the L2 latency
constexpr int CH_STAGES = 3;
constexpr int CH_W_PLANE = CH_BN * CH_KB * 2;      // 16 KB: one fp16 plane of a weight stage
constexpr int CH_W_STAGE = 2 * CH_W_PLANE;         // 32 KB
constexpr int CH_SA = 3;                           // raw-A ring slots (a k-block is two boxes; 3 leaves room for 8 KB patches)
constexpr int CH_A_BYTES = CH_BM * 32 * 4;         // 16 KB: [128 x 32 fp32]
// 16 warps = 4 warpgroups, so that setmaxnreg can move registers between the roles: warps 0..3 = weight producer, MMA issuer,
// raw-A producer, (idle); warps 4..7 = splitters; warps 8..15 = epilogue.  An SM sub-partition holds 16 K registers, i.e.
// 128 per thread at 4 warps each; the epilogue warps (64 accumulators + 32 freshly loaded values + the stage state) spilled
// their ACCUMULATORS at 128 -- the control warpgroup (three busy warps that need < 56 registers) hands them 32 per thread.
constexpr int CH_EPI_WARP0 = 8;
constexpr int CH_THREADS = 512;
constexpr int CH_REGS_CTRL = 56, CH_REGS_SPLIT = 128, CH_REGS_EPI = 160;   // 128 * (56 + 128 + 2 * 160) = 64 512 of 65 536
constexpr int CH_PATCH_BYTES = 8 * 8192;           // per epilogue warp: both [32 x 32] fp32 blocks of its 64 columns
constexpr int CH_BIAS_FLOATS = 128 * (kChainMaxStages - 1) + 256;   // intermediate stages: 128 each; last stage: up to 256
constexpr int CH_SMALL_BYTES = CH_BIAS_FLOATS * 4 + 1024;   // biases, row-scale exponents (4 x 128 int8), barriers, TMEM pointer
constexpr int CH_SMEM_BYTES = CH_STAGES * CH_W_STAGE + CH_SA * CH_A_BYTES + CH_PATCH_BYTES + CH_SMALL_BYTES + 1024;
constexpr int CH_ACT_COL0 = 2 * CH_BN;             // first operand-slot column
constexpr int CH_TMEM_COLS = 512;
static_assert(CH_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(CH_ACT_COL0 + kChainSlots * 64 <= CH_TMEM_COLS, "tensor memory budget");

// power-of-two operand scale from max |v| (bit pattern of a non-negative float): puts the maximum into [2^13, 2^14).
// Returns the exponent k of the INVERSE scale 2^k (0 when the operand is left alone: zero / denormal-ish / inf / NaN).
DCCN_DEVINL int scale_exp_from_amax(unsigned bits) {
  const int e = (int)(bits >> 23) - 127;
  return (bits != 0u && e > -100 && e < 100) ? e - 13 : 0;
}
DCCN_DEVINL float pow2f(int k) { return __uint_as_float((uint32_t)(k + 127) << 23); }

// (the epilogue warps are issue-bound -- two of them share a scheduler with a splitter warp -- so the per-k-block partial-sum
// adds and the stage finalisation run on packed fp32 pairs: pack2 / fma2 / mul2 / sub2, common.cuh)

// The (stage, n-subtile, k-block) nest of one tile, flattened on the host (launch_chain): every role walks the same list.
struct ChainStep {
  uint8_t stage, kb, sub, slot;
  uint8_t flags;       // see CS_*
  uint8_t pad[3];
};
enum { CS_FIRST = 1,   // first k-block of its (stage, n-subtile): the partial sum starts here
       CS_LAST = 2,    // last k-block of its (stage, n-subtile): the stage output is complete after this drain
       CS_READY = 4,   // the MMA warp has to wait for the operand slot (first pass of the stage over its slots)
       CS_TILE_END = 8,   // last step of the tile: operand slots are free for the next tile's staging afterwards
       CS_HBM = 16 };     // the slot is staged from HBM by the splitter warps (first pass of an HBM-fed stage)
constexpr int kChainMaxSteps = 24;
struct ChainSched {
  int nsteps;
  ChainStep steps[kChainMaxSteps];
};

__global__ void __launch_bounds__(CH_THREADS, 1) chain_tc_kernel(const __grid_constant__ ChainParams p,
                                                                  const __grid_constant__ ChainSched sc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem + CH_STAGES * CH_W_STAGE;
  uint8_t* patches = a_ring + CH_SA * CH_A_BYTES;
  float* sbias = reinterpret_cast<float*>(patches + CH_PATCH_BYTES);         // stage s < last: [128] at s * 128; last: [256]
  int8_t* rsexp = reinterpret_cast<int8_t*>(sbias + CH_BIAS_FLOATS);         // [kChainSlots][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(rsexp + kChainSlots * 128);   // [STAGES] weight planes landed
  uint64_t* empty = full + CH_STAGES;     // [STAGES] MMAs of the k-block retired: weight stage free + accumulator ready
  uint64_t* aready = empty + CH_STAGES;   // [kChainSlots] operand slot written (4 warps: splitters or one column group)
  uint64_t* tempty = aready + kChainSlots;   // [2] accumulator drained (8 epilogue warps)
  uint64_t* fullA = tempty + 2;           // [SA] raw A box landed
  uint64_t* emptyA = fullA + CH_SA;       // [SA] raw A box consumed (4 splitter warps)
  uint64_t* sfree = emptyA + CH_SA;       // [1] every MMA of the tile that reads the operand slots has retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // timeline of CTA 0 (role r: trace[r * 1024 + i], i = k-block step counted from the kernel's start): 0 MMAs issued,
  // 1 drain starts (7: the epilogue warp reached the drain's wait), 2 drain done, 3 stage output finished,
  // 4 splitter k-block staged, 5 / 6 intermediate stage: bias + amax done / fp16 pairs packed, 8 / 9 / 10 last stage: patch free /
  // blocks written to the patch / proxy fence done
#ifdef DCCN_CHAIN_TRACE   // tools/build_chain_trace.sh -> libdccn_chtrace.so; the product build carries no trace code
  long long* const trc = (blockIdx.x == 0 && lane == 0) ? p.trace : nullptr;
  int trn = 0;
#define CH_TRACE_AT(role, idx)                                       \
  do {                                                               \
    if (trc && (idx) < 1024) trc[(role) * 1024 + (idx)] = clock64(); \
  } while (0)
#define CH_TRACE_NEXT() ++trn
#else
#define CH_TRACE_AT(role, idx)
#define CH_TRACE_NEXT()
#endif
  const int m_tiles = (p.M + CH_BM - 1) / CH_BM;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;
  const int nsteps = sc.nsteps;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.tmA[i]);
    for (int i = 0; i < p.nst; ++i) {
      tma_prefetch_desc(&p.tmW[i][0]);
      tma_prefetch_desc(&p.tmW[i][1]);
    }
  }
  // the biases of every stage, once per CTA (the epilogue warps read them every tile: broadcast ld.shared instead of 16
  // global loads per thread and stage)
  for (int i = threadIdx.x; i < CH_BIAS_FLOATS; i += CH_THREADS) {
    const int st = i < 128 * (kChainMaxStages - 1) ? i >> 7 : p.nst - 1;
    const int c = i < 128 * (kChainMaxStages - 1) ? i & 127 : i - 128 * (kChainMaxStages - 1);
    float b = 0.f;
    if (st < p.nst - 1) {
      if (p.st[st].bias) b = __ldg(p.st[st].bias + c);
    } else if (i >= 128 * (kChainMaxStages - 1) && p.epi.bias && c < ((p.epi.N + 127) & ~127)) {
      b = __ldg(p.epi.bias + c);
    }
    sbias[i] = b;
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < CH_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < kChainSlots; ++s) mbar_init(&aready[s], 4);
      for (int a = 0; a < CH_SA; ++a) {
        mbar_init(&fullA[a], 1);
        mbar_init(&emptyA[a], 4);
      }
      for (int a = 0; a < 2; ++a) mbar_init(&tempty[a], 8);
      mbar_init(sfree, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, CH_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CH_REGS_CTRL));
  if (warp == 0) {
    // =============================== weight producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int i = 0; i < nsteps; ++i) {
        const ChainStep stp = sc.steps[i];
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], CH_W_STAGE);
          uint8_t* st = smem + stage * CH_W_STAGE;
          tma_load_2d(st, &p.tmW[stp.stage][0], &full[stage], stp.kb * CH_KB, stp.sub * CH_BN);
          tma_load_2d(st + CH_W_PLANE, &p.tmW[stp.stage][1], &full[stage], stp.kb * CH_KB, stp.sub * CH_BN);
        }
        __syncwarp();
        if (++stage == CH_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =====================================
    constexpr uint32_t idesc = umma_idesc_f16(CH_BN, 128);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t arph = 0;                          // phase bit of every operand-slot barrier
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int i = 0; i < nsteps; ++i) {
        const ChainStep stp = sc.steps[i];
        const int slot = stp.slot;
        const bool wait_ready = (stp.flags & CS_READY) != 0;
        // weight planes landed + accumulator drained + (first pass over the slots of this stage) operand slot written
        mbar_wait_multi(&full[stage], phase, &tempty[acc], acc_phase ^ 1, wait_ready ? &aready[slot] : nullptr,
                        (arph >> slot) & 1u);
        if (wait_ready) arph ^= 1u << slot;
        CH_TRACE_AT(0, trn);
        CH_TRACE_NEXT();
        tc_fence_after();
        const uint32_t b_hi = smem_u32(smem + stage * CH_W_STAGE);
        const uint32_t b_lo = b_hi + CH_W_PLANE;
        const uint32_t d = tmem_base + (uint32_t)(acc * CH_BN);
        const uint32_t ta_hi = tmem_base + (uint32_t)(CH_ACT_COL0 + slot * 64);
        if (elect_one()) {
          if (p.small_first) {
            // the 8 cross-term MMAs of the k-block while the accumulator is small, then the 4 hi*hi MMAs (gemm_tc.cuh)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t db_hi = umma_desc_sw128(b_hi + k * 32);
              const uint64_t db_lo = umma_desc_sw128(b_lo + k * 32);
              const uint32_t ka = ta_hi + (uint32_t)(k * 8);
              umma_f16_ts(d, ka + 32, db_hi, idesc, k != 0 ? 1u : 0u);   // A_lo * B_hi
              umma_f16_ts(d, ka, db_lo, idesc, 1u);                      // A_hi * B_lo
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(d, ta_hi + (uint32_t)(k * 8), umma_desc_sw128(b_hi + k * 32), idesc, 1u);   // A_hi * B_hi
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t db_hi = umma_desc_sw128(b_hi + k * 32);
              const uint64_t db_lo = umma_desc_sw128(b_lo + k * 32);
              const uint32_t ka = ta_hi + (uint32_t)(k * 8);
              umma_f16_ts(d, ka + 32, db_hi, idesc, k != 0 ? 1u : 0u);
              umma_f16_ts(d, ka, db_lo, idesc, 1u);
              umma_f16_ts(d, ka, db_hi, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);          // weight stage reusable + accumulator ready (the epilogue waits on it too)
          if (stp.flags & CS_TILE_END) umma_commit(sfree);   // operand slots reusable by the next tile's staging
        }
        __syncwarp();
        if (++stage == CH_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp == 2) {
    // =============================== raw-A producer =================================
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int i = 0; i < nsteps; ++i) {
        const ChainStep stp = sc.steps[i];
        if (!(stp.flags & CS_HBM)) continue;
        const int src = p.st[stp.stage].src;
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) {
          mbar_wait(&emptyA[sa], pa ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&fullA[sa], CH_A_BYTES);
            tma_load_2d(a_ring + sa * CH_A_BYTES, &p.tmA[src], &fullA[sa], stp.kb * CH_KB + bx * 32, tile * CH_BM);
          }
          __syncwarp();
          if (++sa == CH_SA) {
            sa = 0;
            pa ^= 1;
          }
        }
      }
    }
  }   // (warp 3 idles: it only fills the control warpgroup)
  } else if (warp < 8) {
    // (the splitters keep their 128 registers: in gemm_tc_kernel a setmaxnreg.dec of the SPLITTER warpgroup -- and only that one --
    // made full-size passes differ from run to run; here it never showed, but nothing is gained by it either)
    static_assert(CH_REGS_SPLIT == 128, "splitter warpgroup keeps its launch-time registers");
    // =============================== splitters ======================================
    int sa = 0;
    uint32_t pa = 0;
    const int r = (warp & 3) * 32 + lane;
    int iter = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step, ++iter) {
      bool gated = (iter == 0);                 // the previous tile's MMAs no longer read the slots
      for (int i = 0; i < nsteps; ++i) {
        const ChainStep stp = sc.steps[i];
        if (!(stp.flags & CS_HBM)) continue;
        const unsigned* amax_in = p.st[stp.stage].amax_in;
        float a_scale = 1.f;
        if (amax_in) a_scale = pow2f(-scale_exp_from_amax(__ldg(amax_in)));
        const f32x2 a_scale2 = pack2(a_scale, a_scale);
        float hi[32], lo[32];
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) {
          mbar_wait(&fullA[sa], pa);
          const uint32_t rowp = smem_u32(a_ring + sa * CH_A_BYTES + r * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));   // undo the 128B swizzle: chunk c of row r
            f16_split_pack_scaled(v.x, v.y, a_scale2, hi[16 * bx + 2 * c], lo[16 * bx + 2 * c]);
            f16_split_pack_scaled(v.z, v.w, a_scale2, hi[16 * bx + 2 * c + 1], lo[16 * bx + 2 * c + 1]);
          }
          // the release must not overtake the loads still queued in the LSU (gemm_tc.cuh)
#pragma unroll
          for (int c = 0; c < 16; ++c) asm volatile("" : "+f"(hi[16 * bx + c]));
          __syncwarp();
          if (lane == 0) mbar_arrive(&emptyA[sa]);
          if (++sa == CH_SA) {
            sa = 0;
            pa ^= 1;
          }
        }
        if (!gated) {
          mbar_wait(sfree, (uint32_t)((iter - 1) & 1));
          gated = true;
        }
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(CH_ACT_COL0 + stp.slot * 64);
        tmem_st_32x32(ta, hi);
        tmem_st_32x32(ta + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aready[stp.slot]);
        if (warp == 4) {
          CH_TRACE_AT(4, trn);
          CH_TRACE_NEXT();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CH_REGS_EPI));
    // =============================== epilogue warps =================================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int cg = (warp - CH_EPI_WARP0) >> 2;       // column group: accumulator columns cg * 64 .. + 63
    const int rrow = q * 32 + lane;                  // row of the tile this thread owns
#ifdef DCCN_CHAIN_TRACE
    const bool tr_w = warp == CH_EPI_WARP0;
#else
    constexpr bool tr_w = false;
#endif
    const uint32_t patch = smem_u32(patches + (warp - CH_EPI_WARP0) * 8192);
    EpiStore::State est;
    int acc = 0;
    int estage = 0;
    uint32_t ephase = 0;
    f32x2 r[2][16];                                  // this row's 64 output columns of the current stage, as fp32 pairs
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int i = 0; i < nsteps; ++i) {
        const ChainStep stp = sc.steps[i];
        const ChainStage& cs = p.st[stp.stage];
        if (tr_w) CH_TRACE_AT(7, trn);
        mbar_wait(&empty[estage], ephase);
        if (++estage == CH_STAGES) {
          estage = 0;
          ephase ^= 1;
        }
        if (tr_w) CH_TRACE_AT(1, trn);
        tc_fence_after();
        // slot-fed stage: the operand of this k-block carried a per-row power-of-two scale; undo it here (exact)
        float rs = 1.f;
        if (cs.src < 0) rs = pow2f((int)rsexp[stp.slot * 128 + rrow]);
        const f32x2 rs2 = pack2(rs, rs);
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CH_BN + cg * 64);
        if (stp.flags & CS_FIRST) {                    // the partial sum of a (stage, n-subtile) starts at +0
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 16; ++c) r[j][c] = 0ull;
        }
        // partial sums of the k-blocks are added in fp32 round-to-nearest, in k order (v * rs is exact); one code path for
        // the first and the later k-blocks keeps the 64 accumulators in place (no register shuffling between the paths)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[32];
          tmem_ld_32x32(t0 + j * 32, v);
#pragma unroll
          for (int c = 0; c < 16; ++c) r[j][c] = fma2(pack2(v[2 * c], v[2 * c + 1]), rs2, r[j][c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (tr_w) CH_TRACE_AT(2, trn);
        acc ^= 1;
        if (stp.flags & CS_LAST) {
          // ---- r = the layer's pre-bias output columns sub*128 + cg*64 .. +63 of row rrow, up to the operand scales:
          // weights always; HBM operand: the uniform activation scale (slot operands were unscaled per k-block above)
          float o_sc = cs.w_scale_inv;
          if (cs.src >= 0 && cs.amax_in) o_sc *= pow2f(scale_exp_from_amax(__ldg(cs.amax_in)));
          const f32x2 o2 = pack2(o_sc, o_sc);
          if (cs.dst_slot0 < 0) {
            const int row_base = tile * CH_BM + q * 32;
            const int col = stp.sub * CH_BN + cg * 64;
            if (col < p.epi.N && row_base < p.epi.M) {          // warp-uniform
              const float4* bp = reinterpret_cast<const float4*>(sbias + 128 * (kChainMaxStages - 1) + col);
              // both [32 x 32] blocks go into the warp's 8 KB patch (128B-swizzle pattern, conflict-free), ONE proxy fence, then
              // the bulk tensor stores (plain coalesced st.global from the patch measured slower: 0.26 vs 0.22 ms front chain)
              if (lane == 0) tma_store_wait_read();              // the previous tile's blocks have left the patch
              __syncwarp();
              if (tr_w) CH_TRACE_AT(8, trn);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                float v[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 b = bp[j * 8 + c];
                  unpack2(fma2(r[j][2 * c], o2, pack2(b.x, b.y)), v[4 * c], v[4 * c + 1]);
                  unpack2(fma2(r[j][2 * c + 1], o2, pack2(b.z, b.w)), v[4 * c + 2], v[4 * c + 3]);
                }
                // (N = 160: the second block of the last column group is past N -- not recorded, not stored)
                if (p.epi.amax && col + 32 * j < p.epi.N)
                  amax_update_warp<32>(p.epi.amax, v, row_base + lane < p.epi.M, est.amax_seen);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const uint32_t a = patch + (uint32_t)(j * 4096) + (uint32_t)((lane * 8 + (c ^ (lane & 7))) << 4);
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                               "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                               : "memory");
                }
              }
              if (tr_w) CH_TRACE_AT(9, trn);
              fence_proxy_async();                               // generic-proxy writes -> visible to the TMA engine
              __syncwarp();
              if (tr_w) CH_TRACE_AT(10, trn);
              if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  if (col + 32 * j < p.epi.N) {                  // (columns past N inside a block are clipped by the tensor map)
                    tma_store_2d(&p.epi.tm_out, patch + j * 4096, col + 32 * j, row_base);
                    if (p.epi.aux) tma_store_2d(&p.epi.tm_aux, patch + j * 4096, col + 32 * j, row_base);
                  }
                }
                tma_store_commit();
              }
            }
          } else {
            // intermediate layer: y = r * o_sc + bias (one fp32 rounding: the product with the power-of-two scale is exact --
            // the value the layer-by-layer schedule stored), then the fp16 (hi, lo) pair of y * 2^-k with k chosen from this
            // row's 64 values -- the next stage's k-block `cg`
            const float4* bp = reinterpret_cast<const float4*>(sbias + stp.stage * 128 + cg * 64);
            float am = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float4 b[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = bp[j * 8 + h * 4 + c];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  r[j][8 * h + 2 * c] = fma2(r[j][8 * h + 2 * c], o2, pack2(b[c].x, b[c].y));
                  r[j][8 * h + 2 * c + 1] = fma2(r[j][8 * h + 2 * c + 1], o2, pack2(b[c].z, b[c].w));
                }
              }
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                float y0, y1;
                unpack2(r[j][c], y0, y1);
                am = fmaxf(am, fmaxf(fabsf(y0), fabsf(y1)));
              }
            }
            if (tr_w) CH_TRACE_AT(5, trn);
            const int k = scale_exp_from_amax(__float_as_uint(am));
            const int dslot = cs.dst_slot0 + cg;
            rsexp[dslot * 128 + rrow] = (int8_t)k;
            const float as = pow2f(-k);
            const f32x2 as2 = pack2(as, as);
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                // (y0, y1) * 2^-k -> packed fp16 pair hi, and lo = fp16(y * 2^-k - hi)   (f16_split_pack on pairs)
                const f32x2 ys = mul2(r[j][c], as2);
                float y0, y1;
                unpack2(ys, y0, y1);
                const __half2 hh = __floats2half2_rn(y0, y1);
                const float2 hf = __half22float2(hh);
                float l0, l1;
                unpack2(fma2(pack2(hf.x, hf.y), pack2(-1.f, -1.f), ys), l0, l1);   // y - hi, one rounding
                const __half2 ll = __floats2half2_rn(l0, l1);
                hi[16 * j + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hh));
                lo[16 * j + c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&ll));
              }
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(CH_ACT_COL0 + dslot * 64);
            if (tr_w) CH_TRACE_AT(6, trn);
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&aready[dslot]);
          }
          if (tr_w) CH_TRACE_AT(3, trn);
        }
        CH_TRACE_NEXT();
      }
    }
    p.epi.flush(est);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, CH_TMEM_COLS);
}

}  // namespace

int launch_chain(const ChainParams& p, cudaStream_t s, int num_sms) {
  if (p.M <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    DCCN_CUDA_OK(cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
    attr_set = true;
  }
  // flatten the (stage, n-subtile, k-block) nest of one tile
  ChainSched sc;
  memset(&sc, 0, sizeof(sc));
  for (int st = 0; st < p.nst; ++st) {
    const ChainStage& cs = p.st[st];
    DCCN_CHECK(cs.nkb >= 1 && cs.slot0 >= 0 && cs.slot0 + cs.nkb <= kChainSlots, "chain stage %d: k-blocks do not fit the slots", st);
    DCCN_CHECK(cs.dst_slot0 < 0 || (cs.nsub == 1 && cs.dst_slot0 + 2 <= kChainSlots && cs.bias), "chain stage %d: bad destination", st);
    for (int sub = 0; sub < cs.nsub; ++sub)
      for (int kb = 0; kb < cs.nkb; ++kb) {
        DCCN_CHECK(sc.nsteps < kChainMaxSteps, "chain has more than %d k-block steps", kChainMaxSteps);
        ChainStep& t = sc.steps[sc.nsteps++];
        t.stage = (uint8_t)st;
        t.kb = (uint8_t)kb;
        t.sub = (uint8_t)sub;
        t.slot = (uint8_t)(cs.slot0 + kb);
        t.flags = (uint8_t)((kb == 0 ? CS_FIRST : 0) | (kb == cs.nkb - 1 ? CS_LAST : 0) | (sub == 0 ? CS_READY : 0) |
                            (sub == 0 && cs.src >= 0 ? CS_HBM : 0));
      }
  }
  DCCN_CHECK(sc.nsteps >= 1 && p.st[p.nst - 1].dst_slot0 < 0 && p.st[p.nst - 1].src < 0,
             "the last chain stage reads the operand slots and writes to HBM");
  DCCN_CHECK(p.epi.act == 0, "chained kernels end in a linear layer");
  sc.steps[sc.nsteps - 1].flags |= CS_TILE_END;
  const int m_tiles = (p.M + CH_BM - 1) / CH_BM;
  chain_tc_kernel<<<m_tiles < num_sms ? m_tiles : num_sms, CH_THREADS, CH_SMEM_BYTES, s>>>(p, sc);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dccn
