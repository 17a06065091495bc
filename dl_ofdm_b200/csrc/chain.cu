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
//   warps 2..5  splitters: raw fp32 A boxes of the HBM-fed stages -> fp16 (hi, lo) -> tcgen05.st into an operand slot
//   warp 6      producer of the raw-A ring (4 x [128 x 32 fp32])
//   warps 7..14 epilogue: drain every k-block, add; at the end of a stage either (intermediate) add the bias, pick a
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
constexpr int CH_STAGES = 4;                       // weight ring
constexpr int CH_W_PLANE = CH_BN * CH_KB * 2;      // 16 KB: one fp16 plane of a weight stage
constexpr int CH_W_STAGE = 2 * CH_W_PLANE;         // 32 KB
constexpr int CH_SA = 4;                           // raw-A ring slots
constexpr int CH_A_BYTES = CH_BM * 32 * 4;         // 16 KB: [128 x 32 fp32]
constexpr int CH_EPI_WARP0 = 7;
constexpr int CH_THREADS = 32 * CH_EPI_WARP0 + 256;   // 480
constexpr int CH_PATCH_BYTES = 8 * 4096;
constexpr int CH_SMALL_BYTES = 1024;               // row-scale exponents (4 x 128 int8) + barriers + TMEM pointer
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

__global__ void __launch_bounds__(CH_THREADS, 1) chain_tc_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem + CH_STAGES * CH_W_STAGE;
  uint8_t* patches = a_ring + CH_SA * CH_A_BYTES;
  int8_t* rsexp = reinterpret_cast<int8_t*>(patches + CH_PATCH_BYTES);       // [kChainSlots][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(rsexp + kChainSlots * 128);   // [STAGES] weight planes landed
  uint64_t* empty = full + CH_STAGES;     // [STAGES] MMAs of the k-block retired: weight stage free + accumulator ready
  uint64_t* aready = empty + CH_STAGES;   // [kChainSlots] operand slot written (4 warps: splitters or one column group)
  uint64_t* tempty = aready + kChainSlots;   // [2] accumulator drained (8 epilogue warps)
  uint64_t* fullA = tempty + 2;           // [SA] raw A box landed
  uint64_t* emptyA = fullA + CH_SA;       // [SA] raw A box consumed (4 splitter warps)
  uint64_t* sfree = emptyA + CH_SA;       // [1] every MMA of the tile that reads the operand slots has retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // timeline of CTA 0 (role r: trace[r * 1024 + i]); 0 MMA step issued, 1 drain starts, 2 drain done, 3 stage finalised,
  // 4 splitter k-block staged
  long long* const trc = (blockIdx.x == 0 && lane == 0) ? p.trace : nullptr;
  int trn = 0;
#define CH_TRACE(role)                                              \
  do {                                                              \
    if (trc && trn < 1024) trc[(role) * 1024 + trn++] = clock64();  \
  } while (0)
  const int m_tiles = (p.M + CH_BM - 1) / CH_BM;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) tma_prefetch_desc(&p.tmA[i]);
    for (int i = 0; i < p.nst; ++i) {
      tma_prefetch_desc(&p.tmW[i][0]);
      tma_prefetch_desc(&p.tmW[i][1]);
    }
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

  if (warp == 0) {
    // =============================== weight producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int s = 0; s < p.nst; ++s) {
        const ChainStage& cs = p.st[s];
        for (int sub = 0; sub < cs.nsub; ++sub) {
          for (int kb = 0; kb < cs.nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], CH_W_STAGE);
              uint8_t* st = smem + stage * CH_W_STAGE;
              tma_load_2d(st, &p.tmW[s][0], &full[stage], kb * CH_KB, sub * CH_BN);
              tma_load_2d(st + CH_W_PLANE, &p.tmW[s][1], &full[stage], kb * CH_KB, sub * CH_BN);
            }
            __syncwarp();
            if (++stage == CH_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
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
      for (int s = 0; s < p.nst; ++s) {
        const ChainStage& cs = p.st[s];
        for (int sub = 0; sub < cs.nsub; ++sub) {
          for (int kb = 0; kb < cs.nkb; ++kb) {
            const int slot = cs.slot0 + kb;
            // weight planes landed + accumulator drained + (first pass over the slots of this stage) operand slot written
            mbar_wait_multi(&full[stage], phase, &tempty[acc], acc_phase ^ 1, sub == 0 ? &aready[slot] : nullptr,
                            (arph >> slot) & 1u);
            if (sub == 0) arph ^= 1u << slot;
            CH_TRACE(0);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(smem + stage * CH_W_STAGE);
            const uint32_t b_lo = b_hi + CH_W_PLANE;
            const uint32_t d = tmem_base + (uint32_t)(acc * CH_BN);
            const uint32_t ta_hi = tmem_base + (uint32_t)(CH_ACT_COL0 + slot * 64);
            const bool last_of_tile = (s == p.nst - 1) && (sub == cs.nsub - 1) && (kb == cs.nkb - 1);
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
              if (last_of_tile) umma_commit(sfree);   // operand slots reusable by the next tile's staging
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
      }
    }
  } else if (warp == 6) {
    // =============================== raw-A producer =================================
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      for (int s = 0; s < p.nst; ++s) {
        const ChainStage& cs = p.st[s];
        if (cs.src < 0) continue;
        for (int kb = 0; kb < cs.nkb; ++kb) {
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            mbar_wait(&emptyA[sa], pa ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&fullA[sa], CH_A_BYTES);
              tma_load_2d(a_ring + sa * CH_A_BYTES, &p.tmA[cs.src], &fullA[sa], kb * CH_KB + bx * 32, tile * CH_BM);
            }
            __syncwarp();
            if (++sa == CH_SA) {
              sa = 0;
              pa ^= 1;
            }
          }
        }
      }
    }
  } else if (warp < 6) {
    // =============================== splitters ======================================
    int sa = 0;
    uint32_t pa = 0;
    const int r = (warp & 3) * 32 + lane;
    int iter = 0;
    for (int tile = tile0; tile < m_tiles; tile += tile_step, ++iter) {
      bool gated = (iter == 0);                 // the previous tile's MMAs no longer read the slots
      for (int s = 0; s < p.nst; ++s) {
        const ChainStage& cs = p.st[s];
        if (cs.src < 0) continue;
        float a_scale = 1.f;
        if (cs.amax_in) a_scale = pow2f(-scale_exp_from_amax(__ldg(cs.amax_in)));
        for (int kb = 0; kb < cs.nkb; ++kb) {
          float hi[32], lo[32];
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            mbar_wait(&fullA[sa], pa);
            const uint32_t rowp = smem_u32(a_ring + sa * CH_A_BYTES + r * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = lds128(rowp + ((c ^ (r & 7)) << 4));   // undo the 128B swizzle: chunk c of row r
              f16_split_pack(v.x * a_scale, v.y * a_scale, hi[16 * bx + 2 * c], lo[16 * bx + 2 * c]);
              f16_split_pack(v.z * a_scale, v.w * a_scale, hi[16 * bx + 2 * c + 1], lo[16 * bx + 2 * c + 1]);
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
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(CH_ACT_COL0 + (cs.slot0 + kb) * 64);
          tmem_st_32x32(ta, hi);
          tmem_st_32x32(ta + 32, lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aready[cs.slot0 + kb]);
          if (warp == 2) CH_TRACE(4);
        }
      }
    }
  } else {
    // =============================== epilogue warps =================================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int cg = (warp - CH_EPI_WARP0) >> 2;       // column group: accumulator columns cg * 64 .. + 63
    const int rrow = q * 32 + lane;                  // row of the tile this thread owns
    EpiStore::State est;
    int acc = 0;
    int estage = 0;
    uint32_t ephase = 0;
    float r[2][32];
    for (int tile = tile0; tile < m_tiles; tile += tile_step) {
      const int row_base = tile * CH_BM + q * 32;
      for (int s = 0; s < p.nst; ++s) {
        const ChainStage& cs = p.st[s];
        // scale that undoes the operand scales of this stage: weights always; HBM operand: the uniform activation scale
        float o_sc = cs.w_scale_inv;
        if (cs.src >= 0 && cs.amax_in) o_sc *= pow2f(scale_exp_from_amax(__ldg(cs.amax_in)));
        for (int sub = 0; sub < cs.nsub; ++sub) {
          for (int kb = 0; kb < cs.nkb; ++kb) {
            mbar_wait(&empty[estage], ephase);
            if (++estage == CH_STAGES) {
              estage = 0;
              ephase ^= 1;
            }
            if (warp == CH_EPI_WARP0) CH_TRACE(1);
            tc_fence_after();
            // slot-fed stage: the operand of this k-block carried a per-row power-of-two scale; undo it here (exact)
            float rs = 1.f;
            if (cs.src < 0) rs = pow2f((int)rsexp[(cs.slot0 + kb) * 128 + rrow]);
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CH_BN + cg * 64);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float v[32];
              tmem_ld_32x32(t0 + j * 32, v);
              if (kb == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[j][i] = v[i] * rs;
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[j][i] = __fadd_rn(r[j][i], v[i] * rs);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (warp == CH_EPI_WARP0 && trc && trn < 1024) trc[2 * 1024 + trn - 1] = clock64();
            acc ^= 1;
          }
          // ---- end of a (stage, n-subtile): r = the layer's pre-bias output columns sub*128 + cg*64 .. +63 of row rrow ----
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 32; ++i) r[j][i] *= o_sc;
          if (cs.dst_slot0 < 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              p.epi.run_warp(est, row_base, lane, sub * CH_BN + cg * 64 + j * 32, r[j],
                             smem_u32(patches + (warp - CH_EPI_WARP0) * 4096));
          } else {
            // intermediate layer: y = r + bias (fp32, what the layer-by-layer schedule stored), then the fp16 (hi, lo) pair of
            // y * 2^-k with k chosen from this row's 64 values -- the next stage's k-block `cg`
            const float4* bp = reinterpret_cast<const float4*>(cs.bias + cg * 64);
            unsigned am = 0u;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float4 b[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) b[i] = __ldg(bp + j * 8 + i);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                r[j][4 * i + 0] += b[i].x;
                r[j][4 * i + 1] += b[i].y;
                r[j][4 * i + 2] += b[i].z;
                r[j][4 * i + 3] += b[i].w;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) am = max(am, __float_as_uint(r[j][i]) & 0x7FFFFFFFu);
            }
            const int k = scale_exp_from_amax(am);
            const int dslot = cs.dst_slot0 + cg;
            rsexp[dslot * 128 + rrow] = (int8_t)k;
            const float as = pow2f(-k);
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int c = 0; c < 16; ++c)
                f16_split_pack(r[j][2 * c] * as, r[j][2 * c + 1] * as, hi[16 * j + c], lo[16 * j + c]);
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(CH_ACT_COL0 + dslot * 64);
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&aready[dslot]);
          }
          if (warp == CH_EPI_WARP0 && trc && trn < 1024) trc[3 * 1024 + trn - 1] = clock64();
        }
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
  const int m_tiles = (p.M + CH_BM - 1) / CH_BM;
  chain_tc_kernel<<<m_tiles < num_sms ? m_tiles : num_sms, CH_THREADS, CH_SMEM_BYTES, s>>>(p);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dccn
