// common.cuh -- shared helpers + hand-written sm_100a PTX wrappers (mbarrier, TMA, tcgen05).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#ifndef DCCN_DEVINL
#define DCCN_DEVINL __device__ __forceinline__
#endif

namespace dccn {

// ---------------------------------------------------------------------------------
// error plumbing (thread-local message surfaced through dccn_last_error())
// ---------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...);

#define DCCN_CUDA_OK(expr)                                                            \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return dccn::set_error(-1, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                     \
  } while (0)

#define DCCN_CHECK(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return dccn::set_error(-2, __VA_ARGS__);      \
  } while (0)

// ---------------------------------------------------------------------------------
// tf32 split:  v = hi + lo,  hi = rna_tf32(v), lo = rna_tf32(v - hi)
// (both have their low 13 mantissa bits cleared, so the tensor core's own operand
//  truncation is a no-op and  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  carries ~22 bits).
// ---------------------------------------------------------------------------------
DCCN_DEVINL float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
DCCN_DEVINL void tf32_split_cvt(float v, float& hi, float& lo) {
  hi = tf32_rna(v);
  lo = tf32_rna(v - hi);
}
// Same values with full-rate integer ops: rna = add half an ulp of the 10-bit mantissa to the magnitude bits and
// clear the low 13 bits (cvt.rna.tf32.f32 issues at a fraction of the ALU rate and was the longest piece of the
// splitter warps' loop).  Inf / NaN inputs are not expected on this path; a finite value whose rounding overflows the
// exponent becomes inf in both forms.
DCCN_DEVINL float tf32_rna_int(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
DCCN_DEVINL void tf32_split(float v, float& hi, float& lo) {
  hi = tf32_rna_int(v);
  lo = tf32_rna_int(v - hi);
}
// host version (round to nearest, ties away from zero in magnitude like cvt.rna)
inline float tf32_rna_host(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return v;
  u += 0x1000u;
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// ---------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------
DCCN_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DCCN_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DCCN_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DCCN_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DCCN_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware, do not spin
      : "memory");
  return ok != 0;
}
DCCN_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Non-blocking probes of up to three barriers issued back to back (one round trip of ~150-200 clk to the barrier unit
// instead of three in series); the blocking waits only run for a barrier that was not complete yet.  bar1 / bar2 may
// be nullptr.
DCCN_DEVINL void mbar_wait_multi(uint64_t* bar0, uint32_t par0, uint64_t* bar1, uint32_t par1, uint64_t* bar2,
                                 uint32_t par2) {
  uint64_t* b1 = bar1 ? bar1 : bar0;
  uint64_t* b2 = bar2 ? bar2 : bar0;
  const uint32_t q1 = bar1 ? par1 : par0, q2 = bar2 ? par2 : par0;
  uint32_t ok0, ok1, ok2;
  asm volatile(
      "{\n\t.reg .pred p0, p1, p2;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p0, [%3], %4;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p1, [%5], %6;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p2, [%7], %8;\n\t"
      "selp.u32 %0, 1, 0, p0;\n\t"
      "selp.u32 %1, 1, 0, p1;\n\t"
      "selp.u32 %2, 1, 0, p2;\n\t}"
      : "=r"(ok0), "=r"(ok1), "=r"(ok2)
      : "r"(smem_u32(bar0)), "r"(par0), "r"(smem_u32(b1)), "r"(q1), "r"(smem_u32(b2)), "r"(q2)
      : "memory");
  if (!ok0) mbar_wait(bar0, par0);
  if (!ok1) mbar_wait(b1, q1);
  if (!ok2) mbar_wait(b2, q2);
}
// cluster-scope acquire: for barriers that threads of the peer CTA arrive on
DCCN_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
  }
}
// named CTA barriers (bar.sync / bar.arrive): sub-block hand-offs that do not touch the mbarrier unit
DCCN_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
DCCN_DEVINL void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
DCCN_DEVINL float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
DCCN_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DCCN_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) -- 2-D tile load, completion on an mbarrier
// ---------------------------------------------------------------------------------
DCCN_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
DCCN_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tile store, shared -> global (bulk async group; rows / columns outside the tensor are clipped by the hardware)
DCCN_DEVINL void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
DCCN_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the source may be reused)
DCCN_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent bulk store group (two alternating source patches: the latest store reads the OTHER one)
DCCN_DEVINL void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// multicast variant: the box is written to the same shared-memory offset of every CTA in
// cta_mask and completes bytes on the mbarrier at the same offset in each of them
DCCN_DEVINL void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
DCCN_DEVINL uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa_shared)
DCCN_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DCCN_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DCCN_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM load
// ---------------------------------------------------------------------------------
DCCN_DEVINL void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
}
DCCN_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DCCN_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
DCCN_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DCCN_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread for the CTA
DCCN_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]  ("TS" form: the A operand is read from tensor memory,
// 128 lanes = rows, one tf32 element per 32-bit column)
DCCN_DEVINL void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cta_group::2 (CTA pair) forms: one MMA spans both SMs of the pair (M = 256) ----
DCCN_DEVINL void tmem_alloc_pair(uint32_t* smem_out, uint32_t ncols) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
}
DCCN_DEVINL void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
DCCN_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
DCCN_DEVINL void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
DCCN_DEVINL void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// make the mbarrier track completion of all previously issued tcgen05.mma of this thread
DCCN_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of cta_mask
DCCN_DEVINL void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane i <-> TMEM lane base+i)
DCCN_DEVINL void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: lane i of the warp writes 32 consecutive 32-bit columns of TMEM lane base+i
DCCN_DEVINL void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
DCCN_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand tile, rows of 128 bytes (32 fp32),
// TMA SWIZZLE_128B layout, 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100).
// Bit layout per the PTX ISA "tcgen05 shared memory descriptor":
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1         [61,64) layout type: 2 = SWIZZLE_128B
DCCN_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 1024 B between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::tf32: fp32 accumulate, A and B both K-major, M=128, N=n.
//   [4,6) c_format=1 (F32)  [7,10) a_format=2 (TF32)  [10,13) b_format=2 (TF32)
//   [15] a_major=0 [16] b_major=0  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n, int m = 128) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- fp16 hi/lo operands (DCCN_F16X3, staged: not yet run on a GPU) ---------------------------------------------
// kind::f16 with fp16 inputs and fp32 accumulate: a_format = b_format = 0 (F16); K = 16 per instruction, i.e. the same
// 32 bytes of K per k-step as kind::tf32 at twice the MACs.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n, int m = 128) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// TS form with a 16-bit A operand: TMEM lane = row, each 32-bit column holds two consecutive K elements (even k in the
// low half), so one K = 16 instruction reads 8 columns -- the column arithmetic of the tf32 form carries over unchanged.
DCCN_DEVINL void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// (v0, v1) -> packed fp16 pairs hi = fp16(v), lo = fp16(v - hi); v0 in the low half.  hi + lo carries 22 significand
// bits like the tf32 pair as long as lo stays a normal fp16 (|v| >= 2^-3); below that the error is bounded by the
// fp16 subnormal quantum, 2^-25 absolute.  |v| > 65504 overflows: the caller's scale must keep operands below that.
DCCN_DEVINL void f16_split_pack(float v0, float v1, float& hi, float& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
  lo = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2: two IEEE fp32 operations per issue slot, each lane rounded exactly like its
// scalar form).  A pair is a 64-bit register; a scalar operand built as pack2(x, x) is folded into the instruction's
// broadcast form, a pair of adjacent kernel-parameter words into its uniform-register form.
typedef unsigned long long f32x2;
DCCN_DEVINL f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
DCCN_DEVINL void unpack2(f32x2 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
DCCN_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
DCCN_DEVINL f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
DCCN_DEVINL f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// f16_split_pack of (v0, v1) * s on packed pairs: FMUL2, F2FP, 2 x cvt, FFMA2 (y - hi as hi * -1 + y: one rounding, the same
// value as the subtraction), F2FP -- 6 issue slots instead of 8 (the splitter warps are issue-bound)
DCCN_DEVINL void f16_split_pack_scaled(float v0, float v1, f32x2 s2, float& hi, float& lo) {
  const f32x2 ys = mul2(pack2(v0, v1), s2);
  float y0, y1;
  unpack2(ys, y0, y1);
  const __half2 h = __floats2half2_rn(y0, y1);
  const float2 hf = __half22float2(h);
  float l0, l1;
  unpack2(fma2(pack2(hf.x, hf.y), pack2(-1.f, -1.f), ys), l0, l1);
  const __half2 l = __floats2half2_rn(l0, l1);
  hi = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
  lo = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
}

DCCN_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------
// Philox-4x32-10 counter RNG (stateless; key = seed, counter = element index)
// ---------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ explicit Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ inline void operator()(uint64_t ctr, uint32_t stream, uint32_t (&out)[4]) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x5DCC0FD1u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, c0, h0, l0);
      mulhilo(0xCD9E8D57u, c2, h1, l1);
      c0 = h1 ^ c1 ^ a;
      c1 = l1;
      c2 = h0 ^ c3 ^ b;
      c3 = l0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
// two uniforms in (0,1] -> two independent N(0,1) (Box-Muller, fp32)
DCCN_DEVINL void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
  float a = ((float)u0 + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  float b = (float)u1 * 2.3283064365386963e-10f;           // [0,1)
  float r = sqrtf(-2.0f * __logf(a));
  float s, c;
  __sincosf(6.283185307179586f * b, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

}  // namespace dccn
