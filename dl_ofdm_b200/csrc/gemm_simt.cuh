// gemm_simt.cuh -- fp32 FFMA GEMM on CUDA cores (DCCN_PREC_EXACT) with the fused epilogues.
//   C[M,N] = epilogue( A[M,K] * W[K,N] ),  A = A0 (+ A1 when the activation is stored hi/lo)
// 64x64 output tile, BK = 16, 256 threads, 4x4 micro-tile per thread (4 consecutive columns
// of 4 rows, so an (I,Q) column pair stays inside one thread for the pairwise epilogues).
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace dccn {

template <class Epi>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A0, const float* __restrict__ A1,
                                                        int lda, const float* __restrict__ W, int M, int N, int K,
                                                        const __grid_constant__ Epi epi) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 2, a_k = (tid & 3) * 4;   // A tile: 64 rows x 16 k
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;    // W tile: 16 k x 64 n

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      const int r = m0 + a_row;
      if (r < M) {
        const float* p0 = A0 + (size_t)r * lda + k0 + a_k;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + a_k + i < K) {
            a[i] = __ldg(p0 + i);
            if (A1) a[i] += __ldg(A1 + (size_t)r * lda + k0 + a_k + i);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = a[i];
    }
    {
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      const int kk = k0 + b_k;
      if (kk < K) {
        const float* p = W + (size_t)kk * N + n0 + b_n;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (n0 + b_n + i < N) b[i] = __ldg(p + i);
      }
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(b[0], b[1], b[2], b[3]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  typename Epi::State st;
#pragma unroll
  for (int i = 0; i < 4; ++i) epi.template run<4>(st, m0 + ty * 4 + i, n0 + tx * 4, acc[i]);
  epi.flush(st);
}

template <class Epi>
inline int launch_gemm_simt(const float* A0, const float* A1, int lda, const float* W, int M, int N, int K,
                            const Epi& epi, cudaStream_t s) {
  if (M <= 0) return 0;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  gemm_simt_kernel<Epi><<<grid, 256, 0, s>>>(A0, A1, lda, W, M, N, K, epi);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dccn
