// train.cu -- BASELINE config 4: one transfer-learning step of equalizer_ofdm in front of the frozen
// ofdm_dense_rx, on the GPU.
//
// Reference (zhongyuanzhao/dl_ofdm @ 5665b50):
//   total_loss = ce_mean + 0.001 * sum(REGULARIZATION_LOSSES)        dev/py/ofdmreceiver_np_mp.py:335-341
//   ce_mean    = mean softmax-xent applied ON the softmax outputs      dev/py/ofdmreceiver_np.py:154-162
//   regulariser tf.keras.regularizers.l2(0.01) on kernel+bias of the six tf.layers.dense of
//   equalizer_ofdm (dev/py/model.py:370-461); conv3d layers have none
//   AdamOptimizer(exponential_decay(lr0, step, 500, 0.98, staircase)).minimize(total_loss,
//   var_list = Equalizer/*)                                           dev/py/ofdmreceiver_np_mp.py:343-347
// TF's autodiff is replaced by a hand-derived backward pass (restated and pinned in
// oracle/dccn_train_oracle.py):
//   * data gradients (dgrad) reuse the forward GEMM engines (tcgen05 3xTF32 in parity mode, fp32 FFMA in
//     exact mode) on transposed weight operands that are re-derived on the device after every update;
//   * weight gradients (wgrad) are  X^T * dY  contractions over the batch: a split-K fp32 FFMA kernel
//     with deterministic partial sums; an extra all-ones row of X^T yields the bias gradient for free;
//   * every layer's packed GEMM operand is a signed gather of the reference-layout variable (dead taps,
//     [[a,b],[-b,-a]] tiling, Toeplitz expansion, row permutation): the gradient of the variable is the
//     adjoint signed scatter-sum, done with one CSR kernel for all layer kinds;
//   * pointwise backward kernels: demodulation head + loss, phase-only equaliser, tanh.
//
// mode DCCN_TRAIN_RX: training of the basic receiver itself (dev/py/ofdmreceiver_np.py:154-198): every variable of
// ofdm_dense_rx is trainable, total_loss = ce_mean + berlin * 1e-4 * sum(l2) (+ ber, no gradient), berlin = the BER of the
// minibatch (a constant for the gradient: it comes from tf.confusion_matrix(argmax)).  Same machinery: the two GEMM
// layers (fft_like/conv3d, demodulation/dense) go through run_wgrad / the CSR maps, the 200 weights of the
// per-subcarrier head get their gradient from head_wgrad_kernel (warp-shuffle reductions, fixed order).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "handle.cuh"

namespace dccn {

struct TrainParam {
  std::string name;
  int64_t n = 0;
  float *w = nullptr, *m = nullptr, *v = nullptr, *g = nullptr;
  float l2g = 0.f;   // d(REG_COEFF * l * sum w^2)/dw = l2g * w
};

struct TrainLayer {
  GemmLayer* L = nullptr;   // forward layer (lives in the handle)
  GemmLayer bw;             // dgrad companion: C[M, K] = dY[M, N] * W^T
  int kp = -1, bp = -1;     // indices of kernel / bias in TrainState::params
  int bias_kind = 0;        // 0 dense, 1 complex (ba-bb, bb-ba), 2 one complex filter broadcast (Toeplitz layer),
                            // 3 / 4: the layers_conv2d_vector forms of 1 / 2 (re | im channel blocks; (b0, b1) broadcast)
  float* Wp = nullptr;      // [K, N] packed fp32 operand (exact mode: the forward operand itself)
  int32_t* map = nullptr;   // [K*N]  +-(param index + 1), 0 = structural zero
  int32_t* csr_off = nullptr;
  int32_t* csr_idx = nullptr;   // +-(operand index + 1) grouped by parameter
  int max_list = 1;
  float* dWp = nullptr;     // [(K+1), N]: gradient of the packed operand; last row = packed bias gradient
};

struct TrainState {
  dccn_train_cfg cfg;
  int64_t maxB = 0;
  int64_t step = 0;       // global_step: drives the learning-rate schedule on the host, can be set (resume)
  int64_t adam_t = 0;     // updates applied to the CURRENT m / v slots: Adam's bias-correction exponent (beta^t); the
                          // slots start at zero with the training state, so this never follows a set global_step
  std::vector<TrainParam> params;
  int mode = 0;            // DCCN_TRAIN_EQ / DCCN_TRAIN_RX
  int n_layers = 10;
  TrainLayer tl[10];
  // DCCN_TRAIN_RX: the head's variables (conv2d kernel / bias, dense_1 kernel / bias) in `params`, partial sums of
  // head_wgrad_kernel, pinned staging for the D2H refresh of dccn_handle::hw
  // generic equalizer graphs (--opt 1..5): positions of the roles in tl[] (-1 = the graph has no such layer)
  int ix_front1 = -1, ix_front2 = -1, ix_pilot = -1, ix_chain[4] = {-1, -1, -1, -1}, ix_toep = -1, ix_tail1 = -1, ix_tail2 = -1;
  GemmLayer bw_ifft;       // dgrad companion of the constant inverse-DFT layer (tail 'ifft')
  int hp[4] = {-1, -1, -1, -1};
  float* head_partial = nullptr;
  int head_warps = 0;
  float* head_host = nullptr;
  GemmLayer bw_r1, bw_r2;
  Act d_oiq, d_r1o, d_oeq, d_cat, d_eq, d_corr, d_f, d_f2, d_ch, dA, dB, dC, d_p32, d_t1;
  float* partial = nullptr;
  size_t partial_floats = 0;
  // tensor-core wgrad (parity mode): transposed operands  X^T [Kin, M] (fp32)  and  dY^T [Nout, M] (tf32 hi/lo)
  int wgrad_tc = 0;
  float *xT = nullptr, *yT0 = nullptr, *yT1 = nullptr;
  size_t xT_floats = 0, yT_floats = 0;
};

static const char* kLayerVar[10] = {"Equalizer/dense",    "Equalizer/conv3d",   "Equalizer/dense_1", "Equalizer/dense_2",
                                    "Equalizer/dense_3",  "Equalizer/dense_4",  "Equalizer/conv3d_1", "Equalizer/conv3d_2",
                                    "Equalizer/conv3d_3", "Equalizer/dense_5"};
static const int kBiasKind[10] = {0, 1, 0, 0, 0, 0, 2, 1, 1, 0};
static const int kPerSymbol[10] = {1, 1, 0, 0, 0, 0, 0, 1, 1, 1};   // layer contracts over B*S rows (else B)
static const char* kRxLayerVar[2] = {"fft_like/conv3d", "demodulation/dense"};
static const int kRxBiasKind[2] = {1, 0};
static const int kRxPerSymbol[2] = {1, 0};
static const char* kRxHeadVar[4] = {"demodulation/conv2d/kernel", "demodulation/conv2d/bias", "demodulation/dense_1/kernel",
                                    "demodulation/dense_1/bias"};

// =========================================================================================
// kernels
// =========================================================================================

// ---- wgrad: P[z][i][j] = sum_{m in split z} Xaug[m][i] * Y[m][j],  Xaug = [X | 1]  (i <= Kin) -----
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const float* __restrict__ X, int ldx, int Kin, const float* __restrict__ Y, int ldy, int Nout,
                  long long M, int rps, float* __restrict__ P) {
  constexpr int BK = 16;
  __shared__ __align__(16) float As[BK][64];
  __shared__ __align__(16) float Bs[BK][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const long long m_begin = (long long)blockIdx.z * rps;
  const long long m_end = (m_begin + rps < M) ? m_begin + rps : M;
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  const int gi = i0 + lc, gj = j0 + lc;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long m0 = m_begin; m0 < m_end; m0 += BK) {
    const long long m = m0 + lr;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m < m_end) {
      if (gi + 3 < Kin) a = __ldg(reinterpret_cast<const float4*>(X + (size_t)m * ldx + gi));
      else if (gi == Kin) a.x = 1.f;                               // the all-ones row -> column sums of dY
      if (gj + 3 < Nout) b = __ldg(reinterpret_cast<const float4*>(Y + (size_t)m * ldy + gj));
    }
    *reinterpret_cast<float4*>(&As[lr][lc]) = a;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = b;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(aa[i], bv.x, acc[i][0]);
        acc[i][1] = fmaf(aa[i], bv.y, acc[i][1]);
        acc[i][2] = fmaf(aa[i], bv.z, acc[i][2]);
        acc[i][3] = fmaf(aa[i], bv.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  float* Pz = P + (size_t)blockIdx.z * (size_t)(Kin + 1) * Nout;
  const int oj = j0 + tx * 4;
  if (oj < Nout) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oi = i0 + ty * 4 + i;
      if (oi <= Kin)
        *reinterpret_cast<float4*>(Pz + (size_t)oi * Nout + oj) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// deterministic sum of the split-K partials (slice z starts at z * stride)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ P, int splits, long long n, long long stride, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += P[(size_t)z * stride + i];
  out[i] = s;
}

// X [M, C] (row pitch ld) -> T [C, M] (row pitch ldt); split != 0 writes tf32 hi / lo planes
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ X, int ld, int C, long long M, float* __restrict__ T0, float* __restrict__ T1,
                 long long ldt, int split) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32;
  const long long m0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long m = m0 + r;
    const int c = c0 + tx;
    tile[r][tx] = (m < M && c < C) ? X[(size_t)m * ld + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long m = m0 + tx;
    if (c < C && m < M) {
      const float v = tile[tx][r];
      if (split) {
        float hi, lo;
        tf32_split(v, hi, lo);
        T0[(size_t)c * ldt + m] = hi;
        T1[(size_t)c * ldt + m] = lo;
      } else T0[(size_t)c * ldt + m] = v;
    }
  }
}

// out[n] = sum_m (T0[n][m] + T1[n][m]): the bias gradient (column sums of dY) from the transposed planes
__global__ void __launch_bounds__(256)
rowsum_kernel(const float* __restrict__ T0, const float* __restrict__ T1, long long M, long long ldt, float* __restrict__ out) {
  __shared__ float red[256];
  const float* a = T0 + (size_t)blockIdx.x * ldt;
  const float* b = T1 + (size_t)blockIdx.x * ldt;
  float s = 0.f;
  for (long long m = threadIdx.x; m < M; m += 256) s += a[m] + b[m];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// ---- packed operand <-> reference-layout variable ---------------------------------------------------
__global__ void __launch_bounds__(256)
pack_map_kernel(const float* __restrict__ w, const int32_t* __restrict__ map, long long n, float* __restrict__ Wp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = map[i];
  Wp[i] = m == 0 ? 0.f : (m > 0 ? w[m - 1] : -w[-m - 1]);
}

// gradient of the variable = signed sum of the operand gradient over every place the variable was gathered to
__global__ void __launch_bounds__(256)
grad_map_thread_kernel(const float* __restrict__ dWp, const int32_t* __restrict__ off, const int32_t* __restrict__ idx,
                       int n_param, float* __restrict__ g) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_param) return;
  float s = 0.f;
  for (int e = off[p]; e < off[p + 1]; ++e) {
    const int v = idx[e];
    s += v > 0 ? dWp[v - 1] : -dWp[-v - 1];
  }
  g[p] = s;
}

__global__ void __launch_bounds__(256)
grad_map_warp_kernel(const float* __restrict__ dWp, const int32_t* __restrict__ off, const int32_t* __restrict__ idx,
                     int n_param, float* __restrict__ g) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n_param) return;
  float s = 0.f;
  for (int e = off[p] + lane; e < off[p + 1]; e += 32) {
    const int v = idx[e];
    s += v > 0 ? dWp[v - 1] : -dWp[-v - 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) g[p] = s;
}

__global__ void __launch_bounds__(256)
pack_bias_kernel(const float* __restrict__ b, int kind, int N, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  if (kind == 0) out[j] = b[j];
  else if (kind == 1) {
    const int F = N >> 1, f = j >> 1;
    out[j] = (j & 1) ? b[F + f] - b[f] : b[f] - b[F + f];
  } else if (kind == 2) out[j] = (j & 1) ? b[1] - b[0] : b[0] - b[1];
  else if (kind == 3) out[j] = (j & 1) ? b[(N >> 1) + (j >> 1)] : b[j >> 1];   // layers_conv2d_vector: re / im channel blocks
  else out[j] = b[j & 1];                                                     // vector (S,K) layer: (b0, b1) broadcast
}

// one block; dbp = packed bias gradient [N]
__global__ void __launch_bounds__(256)
grad_bias_kernel(const float* __restrict__ dbp, int kind, int N, float* __restrict__ g) {
  if (kind == 0) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) g[j] = dbp[j];
  } else if (kind == 3) {
    const int F = N >> 1;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      g[f] = dbp[2 * f];
      g[F + f] = dbp[2 * f + 1];
    }
  } else if (kind == 4) {
    __shared__ float red0[256], red1[256];
    float s0 = 0.f, s1 = 0.f;
    for (int f = threadIdx.x; f < (N >> 1); f += blockDim.x) {
      s0 += dbp[2 * f];
      s1 += dbp[2 * f + 1];
    }
    red0[threadIdx.x] = s0;
    red1[threadIdx.x] = s1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) {
        red0[threadIdx.x] += red0[threadIdx.x + o];
        red1[threadIdx.x] += red1[threadIdx.x + o];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      g[0] = red0[0];
      g[1] = red1[0];
    }
  } else if (kind == 1) {
    const int F = N >> 1;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const float ga = dbp[2 * f] - dbp[2 * f + 1];
      g[f] = ga;
      g[F + f] = -ga;
    }
  } else {
    __shared__ float red[256];
    float s = 0.f;
    for (int f = threadIdx.x; f < (N >> 1); f += blockDim.x) s += dbp[2 * f] - dbp[2 * f + 1];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      g[0] = red[0];
      g[1] = -red[0];
    }
  }
}

// Wp [K,N] -> plain (P*) and transposed (T*) copies, optionally split into tf32 hi/lo planes
__global__ void __launch_bounds__(256)
repack_kernel(const float* __restrict__ Wp, int K, int N, float* __restrict__ T0, float* __restrict__ T1,
              float* __restrict__ P0, float* __restrict__ P1, int split) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    float v = 0.f;
    if (k < K && n < N) {
      v = Wp[(size_t)k * N + n];
      if (P0) {
        if (split) {
          float hi, lo;
          tf32_split(v, hi, lo);
          P0[(size_t)k * N + n] = hi;
          if (P1) P1[(size_t)k * N + n] = lo;
        } else if (P0 != Wp) P0[(size_t)k * N + n] = v;
      }
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  if (!T0) return;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    if (k < K && n < N) {
      const float v = tile[tx][r];
      if (split) {
        float hi, lo;
        tf32_split(v, hi, lo);
        T0[(size_t)n * K + k] = hi;
        if (T1) T1[(size_t)n * K + k] = lo;
      } else T0[(size_t)n * K + k] = v;
    }
  }
}

// g += l2g * w  (gradient of the regulariser), so that g is d total_loss / d var
__global__ void __launch_bounds__(256) add_l2_kernel(float* __restrict__ g, const float* __restrict__ w, long long n, float l2g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] = fmaf(l2g, w[i], g[i]);
}

// tf.train.AdamOptimizer (TF 1.15): epsilon outside the square root, lr_t carries the bias corrections
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, long long n,
            float lr_t, float b1, float b2, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gt = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gt;
  const float vi = b2 * v[i] + (1.f - b2) * gt * gt;
  m[i] = mi;
  v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// ---- pointwise backward kernels --------------------------------------------------------------------
// demodulation head (dev/py/model.py:1275-1291) + loss (dev/py/ofdmreceiver_np.py:154-162):
// d ce_mean / d out_iq per data subcarrier; the head is recomputed from out_iq.
template <int NB>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ out_iq, const uint8_t* __restrict__ bits, const __grid_constant__ HeadWeights hw,
                long long total, float inv_n, float* __restrict__ d_oiq) {
  constexpr int MO = 1 << NB;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float2 iq = __ldg(reinterpret_cast<const float2*>(out_iq) + i);
    const float I = iq.x, Q = iq.y;
    float hpre[MO], hh[MO], dh[MO];
#pragma unroll
    for (int m = 0; m < MO; ++m) {
      hpre[m] = I * hw.Wc[0][m] + Q * hw.Wc[1][m] + hw.bc[m];
      hh[m] = fmaxf(0.2f * hpre[m], hpre[m]);
      dh[m] = 0.f;
    }
    float dI = 0.f, dQ = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      float l0 = hw.b1[2 * k], l1 = hw.b1[2 * k + 1];
#pragma unroll
      for (int m = 0; m < MO; ++m) {
        l0 += hh[m] * hw.W1[m][2 * k];
        l1 += hh[m] * hw.W1[m][2 * k + 1];
      }
      l0 += I * hw.W1[MO][2 * k] + Q * hw.W1[MO + 1][2 * k];
      l1 += I * hw.W1[MO][2 * k + 1] + Q * hw.W1[MO + 1][2 * k + 1];
      const float s0 = l0 > 0.f ? 1.f : 0.2f, s1 = l1 > 0.f ? 1.f : 0.2f;   // LeakyReluGrad
      const float a0 = fmaxf(0.2f * l0, l0), a1 = fmaxf(0.2f * l1, l1);
      const float t = expf(-fabsf(a1 - a0));
      const float pb = 1.0f / (1.0f + t), ps = t / (1.0f + t);
      const bool one_big = a1 > a0;
      const float p0 = one_big ? ps : pb, p1 = one_big ? pb : ps;
      // loss = logsumexp(p) - p_y  ->  d/dp = softmax(p) - onehot(y)
      const float q1 = 1.0f / (1.0f + expf(p0 - p1)), q0 = 1.0f - q1;
      const unsigned y = bits[i * NB + k] & 1u;
      const float dp0 = (q0 - (y ? 0.f : 1.f)) * inv_n, dp1 = (q1 - (y ? 1.f : 0.f)) * inv_n;
      // through the model's softmax: da_j = p_j (dp_j - sum dp p)  =>  da0 = p0 p1 (dp0 - dp1) = -da1
      const float da0 = p0 * p1 * (dp0 - dp1);
      const float dl0 = da0 * s0, dl1 = -da0 * s1;
#pragma unroll
      for (int m = 0; m < MO; ++m) dh[m] += hw.W1[m][2 * k] * dl0 + hw.W1[m][2 * k + 1] * dl1;
      dI += hw.W1[MO][2 * k] * dl0 + hw.W1[MO][2 * k + 1] * dl1;
      dQ += hw.W1[MO + 1][2 * k] * dl0 + hw.W1[MO + 1][2 * k + 1] * dl1;
    }
#pragma unroll
    for (int m = 0; m < MO; ++m) {
      const float d = dh[m] * (hpre[m] > 0.f ? 1.f : 0.2f);
      dI += hw.Wc[0][m] * d;
      dQ += hw.Wc[1][m] * d;
    }
    reinterpret_cast<float2*>(d_oiq)[i] = make_float2(dI, dQ);
  }
}

// Gradients of the head's own weights (DCCN_TRAIN_RX).  Layout of one partial / of the result, NQ = 2*MO + MO +
// (MO+2)*2NB + 2NB floats:  [dWc (2 x MO) | dbc (MO) | dW1 ((MO+2) x 2NB) | db1 (2NB)].
// One thread per (frame, data subcarrier) position recomputes the head and its backward; every quantity is summed over
// the warp with a shuffle butterfly and kept by lane q % 32 (slot q / 32), so a warp carries its NQ running sums in 7
// registers per lane; partial[warp][q] is reduced in a fixed order by head_wgrad_reduce_kernel (deterministic).
template <int NB>
__global__ void __launch_bounds__(256)
head_wgrad_kernel(const float* __restrict__ out_iq, const uint8_t* __restrict__ bits, const __grid_constant__ HeadWeights hw,
                  long long total, float inv_n, float* __restrict__ partial) {
  constexpr int MO = 1 << NB;
  constexpr int NQ = 2 * MO + MO + (MO + 2) * 2 * NB + 2 * NB;
  constexpr int NS = (NQ + 31) / 32;
  const int lane = threadIdx.x & 31;
  float acc[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) acc[i] = 0.f;
  auto add = [&](int q, float v) {   // q is a compile-time constant after unrolling
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == (q & 31)) acc[q >> 5] += v;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long iters = (total + stride - 1) / stride;       // every lane runs every iteration (shuffles)
  for (long long it = 0; it < iters; ++it) {
    const long long i = i0 + it * stride;
    const bool live = i < total;
    const float2 iq = live ? __ldg(reinterpret_cast<const float2*>(out_iq) + i) : make_float2(0.f, 0.f);
    const float I = iq.x, Q = iq.y;
    float hpre[MO], hh[MO], dh[MO], dl[2 * NB];
#pragma unroll
    for (int m = 0; m < MO; ++m) {
      hpre[m] = I * hw.Wc[0][m] + Q * hw.Wc[1][m] + hw.bc[m];
      hh[m] = fmaxf(0.2f * hpre[m], hpre[m]);
      dh[m] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      float l0 = hw.b1[2 * k], l1 = hw.b1[2 * k + 1];
#pragma unroll
      for (int m = 0; m < MO; ++m) {
        l0 += hh[m] * hw.W1[m][2 * k];
        l1 += hh[m] * hw.W1[m][2 * k + 1];
      }
      l0 += I * hw.W1[MO][2 * k] + Q * hw.W1[MO + 1][2 * k];
      l1 += I * hw.W1[MO][2 * k + 1] + Q * hw.W1[MO + 1][2 * k + 1];
      const float s0 = l0 > 0.f ? 1.f : 0.2f, s1 = l1 > 0.f ? 1.f : 0.2f;
      const float a0 = fmaxf(0.2f * l0, l0), a1 = fmaxf(0.2f * l1, l1);
      const float t = expf(-fabsf(a1 - a0));
      const float pb = 1.0f / (1.0f + t), ps = t / (1.0f + t);
      const bool one_big = a1 > a0;
      const float p0 = one_big ? ps : pb, p1 = one_big ? pb : ps;
      const float q1 = 1.0f / (1.0f + expf(p0 - p1)), q0 = 1.0f - q1;
      const unsigned y = live ? (bits[i * NB + k] & 1u) : 0u;
      const float dp0 = (q0 - (y ? 0.f : 1.f)) * inv_n, dp1 = (q1 - (y ? 1.f : 0.f)) * inv_n;
      const float da0 = live ? p0 * p1 * (dp0 - dp1) : 0.f;
      dl[2 * k] = da0 * s0;
      dl[2 * k + 1] = -da0 * s1;
#pragma unroll
      for (int m = 0; m < MO; ++m) dh[m] += hw.W1[m][2 * k] * dl[2 * k] + hw.W1[m][2 * k + 1] * dl[2 * k + 1];
    }
    // dWc, dbc
#pragma unroll
    for (int m = 0; m < MO; ++m) {
      const float d = dh[m] * (hpre[m] > 0.f ? 1.f : 0.2f);
      add(m, I * d);
      add(MO + m, Q * d);
      add(2 * MO + m, d);
    }
    // dW1 (rows: hh[0..MO), I, Q), db1
#pragma unroll
    for (int j = 0; j < 2 * NB; ++j) {
#pragma unroll
      for (int m = 0; m < MO; ++m) add(3 * MO + m * 2 * NB + j, hh[m] * dl[j]);
      add(3 * MO + MO * 2 * NB + j, I * dl[j]);
      add(3 * MO + (MO + 1) * 2 * NB + j, Q * dl[j]);
      add(3 * MO + (MO + 2) * 2 * NB + j, dl[j]);
    }
  }
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int q = sl * 32 + lane;
    if (q < NQ) partial[warp_id * NQ + q] = acc[sl];
  }
}

// g[q] = sum over warps of partial[w][q], fixed order; the four head variables are consecutive slices of the NQ vector
__global__ void __launch_bounds__(256)
head_wgrad_reduce_kernel(const float* __restrict__ partial, int warps, int NQ, float* __restrict__ g_wc, int n_wc,
                         float* __restrict__ g_bc, int n_bc, float* __restrict__ g_w1, int n_w1, float* __restrict__ g_b1) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= NQ) return;
  float s = 0.f;
  for (int w = 0; w < warps; ++w) s += partial[(size_t)w * NQ + q];
  if (q < n_wc) g_wc[q] = s;
  else if (q < n_wc + n_bc) g_bc[q - n_wc] = s;
  else if (q < n_wc + n_bc + n_w1) g_w1[q - n_wc - n_bc] = s;
  else g_b1[q - n_wc - n_bc - n_w1] = s;
}

// g += (coef * berlin) * w with berlin = (c01 + c10) / sum of this minibatch's confusion matrix (ofdmreceiver_np.py:173)
__global__ void __launch_bounds__(256)
add_l2_ber_kernel(float* __restrict__ g, const float* __restrict__ w, long long n, float coef,
                  const unsigned long long* __restrict__ conf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double tot = (double)(conf[0] + conf[1] + conf[2] + conf[3]);
  const float berlin = tot > 0.0 ? (float)((double)(conf[1] + conf[2]) / tot) : 0.f;
  g[i] = fmaf(coef * berlin, w[i], g[i]);
}

// phase-only equaliser eq = f * conj(c)/|c|, corr = |eq|^2 (dev/py/model.py:430-437), per complex point
__global__ void __launch_bounds__(256)
phaseeq_bwd_kernel(const float2* __restrict__ deq, const float* __restrict__ dcorr, const float2* __restrict__ f,
                   const float2* __restrict__ ch, const float2* __restrict__ eq, long long total,
                   float2* __restrict__ df, float2* __restrict__ dch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float2 e = eq[i], c = ch[i], ff = f[i], de = deq[i];
  const float dc = dcorr ? dcorr[i] : 0.f;     // graphs without the correlation branch pass nullptr
  const float der = fmaf(2.f * e.x, dc, de.x), dei = fmaf(2.f * e.y, dc, de.y);
  const float inv = rsqrtf(c.x * c.x + c.y * c.y);
  const float nr = c.x * inv, ni = -c.y * inv;
  df[i] = make_float2(der * nr + dei * ni, -der * ni + dei * nr);
  const float dnr = der * ff.x + dei * ff.y, dni = -der * ff.y + dei * ff.x;
  const float com = (dnr * c.y + dni * c.x) * inv * inv * inv;
  dch[i] = make_float2(c.y * com, -c.x * com);
}

__global__ void __launch_bounds__(256) tanh_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] *= (1.f - y[i] * y[i]);
}

__global__ void __launch_bounds__(256) add_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// =========================================================================================
// host side
// =========================================================================================
static inline unsigned blocks_for(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

// tile width and split-K factor of the tensor-core wgrad GEMM  [Kin, M] x [M, Nout]
static int wgrad_bn(int Nout) { return Nout <= 32 ? 32 : 128; }
static int wgrad_ksplit(int Kin, int Nout, int64_t M, int num_sms) {
  const int m_tiles = (Kin + 127) / 128, n_tiles = (Nout + wgrad_bn(Nout) - 1) / wgrad_bn(Nout);
  const int num_kb = (int)((M + 31) / 32);
  int ksplit = (2 * num_sms + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);   // ~two waves of tiles
  if (ksplit > num_kb / 2) ksplit = num_kb / 2;
  if (ksplit < 1) ksplit = 1;
  const int per = (num_kb + ksplit - 1) / ksplit;
  return (num_kb + per - 1) / per;                                             // every slice owns >= 1 k-block
}

static int rows_per_split(int64_t M) {
  int64_t rps = 512;
  if ((M + rps - 1) / rps > 32) rps = ((M + 31) / 32 + 15) / 16 * 16;
  return (int)rps;
}

// wgrad of one layer + reduction to the variable gradients
static int run_wgrad(dccn_handle* h, TrainLayer& t, const float* X, int ldx, const float* Y, int ldy, int64_t M,
                     cudaStream_t s) {
  TrainState* tr = h->tr;
  const int Kin = t.L->K, Nout = t.L->N;
  DCCN_CHECK(Kin % 4 == 0 && Nout % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "wgrad operands must be float4-aligned");
  if (tr->wgrad_tc) {
    // ---- tcgen05 path: dWp[Kin, Nout] = X^T[Kin, M] * (dY^T[Nout, M])^T, split-K over the batch ----------------
    const long long ldt = (M + 3) & ~3LL;
    DCCN_CHECK((size_t)Kin * ldt <= tr->xT_floats && (size_t)Nout * ldt <= tr->yT_floats, "transpose scratch too small");
    GemmLayer wl;
    wl.K = (int)M;
    wl.N = Nout;
    wl.BN = wgrad_bn(Nout);
    wl.dWt0 = tr->yT0;
    wl.dWt1 = tr->yT1;
    int rc = make_tmap(&wl.tmB0, tr->yT0, Nout, M, ldt, wl.BN);
    if (rc) return rc;
    if ((rc = make_tmap(&wl.tmB1, tr->yT1, Nout, M, ldt, wl.BN))) return rc;
    const int m_tiles = (Kin + 127) / 128;
    const int ksplit = wgrad_ksplit(Kin, Nout, M, h->num_sms);
    const size_t stride = (size_t)m_tiles * 128 * Nout;
    DCCN_CHECK(stride * ksplit <= tr->partial_floats, "wgrad scratch too small (%zu > %zu)", stride * ksplit,
               tr->partial_floats);
    {
      LaunchScope ls(h, SLOT_T_WGRAD, s, 3);
      dim3 gx((Kin + 31) / 32, (unsigned)((M + 31) / 32)), gy((Nout + 31) / 32, (unsigned)((M + 31) / 32));
      transpose_kernel<<<gx, 256, 0, s>>>(X, ldx, Kin, (long long)M, tr->xT, nullptr, ldt, 0);
      transpose_kernel<<<gy, 256, 0, s>>>(Y, ldy, Nout, (long long)M, tr->yT0, tr->yT1, ldt, 1);
      rowsum_kernel<<<Nout, 256, 0, s>>>(tr->yT0, tr->yT1, (long long)M, ldt, t.dWp + (size_t)Kin * Nout);
      DCCN_CUDA_OK(cudaGetLastError());
    }
    Act A;
    A.p0 = tr->xT;
    A.ld = (int)ldt;
    Act P;
    P.p0 = tr->partial;
    P.ld = Nout;
    EpiStore e = store_epi(wl, P, 0, (int64_t)ksplit * m_tiles * 128);
    e.bias = nullptr;
    if ((rc = run_gemm_store(h, SLOT_T_WGRAD, wl, A, 0, Kin, e, s, ksplit))) return rc;
    LaunchScope ls(h, SLOT_T_WGRAD, s, 1);
    reduce_partials_kernel<<<blocks_for((long long)Kin * Nout), 256, 0, s>>>(tr->partial, ksplit, (long long)Kin * Nout,
                                                                           (long long)stride, t.dWp);
  } else {
    const int rps = rows_per_split(M);
    const int splits = (int)((M + rps - 1) / rps);
    const size_t n = (size_t)(Kin + 1) * Nout;
    DCCN_CHECK(n * splits <= tr->partial_floats, "wgrad scratch too small (%zu > %zu)", n * splits, tr->partial_floats);
    LaunchScope ls(h, SLOT_T_WGRAD, s, 2);
    dim3 grid((Nout + 63) / 64, (Kin + 1 + 63) / 64, splits);
    wgrad_simt_kernel<<<grid, 256, 0, s>>>(X, ldx, Kin, Y, ldy, Nout, (long long)M, rps, tr->partial);
    reduce_partials_kernel<<<blocks_for((long long)n), 256, 0, s>>>(tr->partial, splits, (long long)n, (long long)n, t.dWp);
  }
  LaunchScope ls(h, SLOT_T_ADAM, s, 2);
  TrainParam& kp = tr->params[t.kp];
  TrainParam& bp = tr->params[t.bp];
  if (t.max_list <= 4)
    grad_map_thread_kernel<<<blocks_for(kp.n), 256, 0, s>>>(t.dWp, t.csr_off, t.csr_idx, (int)kp.n, kp.g);
  else
    grad_map_warp_kernel<<<blocks_for(kp.n * 32), 256, 0, s>>>(t.dWp, t.csr_off, t.csr_idx, (int)kp.n, kp.g);
  grad_bias_kernel<<<1, 256, 0, s>>>(t.dWp + (size_t)Kin * Nout, t.bias_kind, Nout, bp.g);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_dgrad(dccn_handle* h, const GemmLayer& bw, const Act& dY, int col_off, int64_t M, const Act& dst,
                     int dst_col_off, cudaStream_t s) {
  return run_gemm_store(h, SLOT_T_DGRAD, bw, dY, col_off, M, store_epi(bw, dst, dst_col_off, M), s);
}

// variables -> packed operands -> forward / dgrad operand copies (after an update or a commit)
static int repack_all(dccn_handle* h, cudaStream_t s) {
  TrainState* tr = h->tr;
  const bool split = h->cfg.precision == DCCN_PREC_PARITY;
  for (int i = 0; i < tr->n_layers; ++i) {
    TrainLayer& t = tr->tl[i];
    const int K = t.L->K, N = t.L->N;
    LaunchScope ls(h, SLOT_T_REPACK, s, 3);
    pack_map_kernel<<<blocks_for((long long)K * N), 256, 0, s>>>(tr->params[t.kp].w, t.map, (long long)K * N, t.Wp);
    pack_bias_kernel<<<blocks_for(N), 256, 0, s>>>(tr->params[t.bp].w, t.bias_kind, N, t.L->dBias);
    dim3 grid((N + 31) / 32, (K + 31) / 32);
    if (split)   // forward B operand [N,K] hi/lo (transposed), dgrad B operand [K,N] hi/lo (plain)
      repack_kernel<<<grid, 256, 0, s>>>(t.Wp, K, N, t.L->dWt0, t.L->dWt1, t.bw.dWt0, t.bw.dWt1, 1);
    else         // forward SIMT operand is Wp itself; dgrad SIMT operand W^T [N,K]
      repack_kernel<<<grid, 256, 0, s>>>(t.Wp, K, N, t.bw.dW, nullptr, nullptr, nullptr, 0);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

static int make_bw_layer(dccn_handle* h, const GemmLayer& L, GemmLayer* bw, cudaStream_t s) {
  bw->K = L.N;
  bw->N = L.K;
  bw->W.resize((size_t)L.K * L.N);
  for (int k = 0; k < L.K; ++k)
    for (int n = 0; n < L.N; ++n) bw->W[(size_t)n * L.K + k] = L.W[(size_t)k * L.N + n];
  bw->bias.assign(bw->N, 0.f);
  bw->fused = false;
  return upload_layer(h, bw, s);
}

static int upload_params(dccn_handle* h, cudaStream_t s) {
  for (TrainParam& p : h->tr->params) {
    const HostTensor* t = find(h, p.name);
    DCCN_CHECK(t && (int64_t)t->data.size() == p.n, "weight '%s' missing or resized", p.name.c_str());
    DCCN_CUDA_OK(cudaMemcpyAsync(p.w, t->data.data(), (size_t)p.n * 4, cudaMemcpyHostToDevice, s));
  }
  DCCN_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

void train_free(dccn_handle* h) {
  if (h->tr && h->tr->head_host) cudaFreeHost(h->tr->head_host);
  delete h->tr;   // device memory is owned by the handle's allocation list
  h->tr = nullptr;
}

int train_on_commit(dccn_handle* h, cudaStream_t s) {
  int rc = upload_params(h, s);
  if (rc) return rc;
  return repack_all(h, s);
}

int train_fetch_weight(dccn_handle* h, const char* tf_name, HostTensor* t) {
  for (TrainParam& p : h->tr->params)
    if (p.name == tf_name) {
      DCCN_CUDA_OK(cudaMemcpy(t->data.data(), p.w, (size_t)p.n * 4, cudaMemcpyDeviceToHost));
      return 1;
    }
  return 0;
}

static int backward(dccn_handle* h, int64_t B, const uint8_t* bits, cudaStream_t s) {
  TrainState* tr = h->tr;
  const int S = h->S, K = h->K, T = h->T, Tin = h->Tin, D = h->D, NB = h->NB, F = h->F;
  const int64_t MS = B * S;
  const int cp_off = (T - Tin) * 2;
  const int SK2 = S * K * 2;
  int rc;
  // ---- loss + demodulation head ------------------------------------------------------------------
  {
    LaunchScope ls(h, SLOT_T_HEAD, s);
    const long long total = (long long)B * D;
    const float inv_n = (float)(1.0 / ((double)B * D * NB));
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    switch (NB) {
      case 1: head_bwd_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      case 2: head_bwd_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      case 3: head_bwd_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      default: head_bwd_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
    }
    DCCN_CUDA_OK(cudaGetLastError());
  }
  // ---- frozen receiver: data gradients only ------------------------------------------------------
  if ((rc = run_dgrad(h, tr->bw_r2, tr->d_oiq, 0, B, tr->d_r1o, 0, s))) return rc;
  Act d_r1v = tr->d_r1o;  d_r1v.ld = 2 * F;
  Act d_oeqv = tr->d_oeq; d_oeqv.ld = 2 * T;
  if (cp_off) DCCN_CUDA_OK(cudaMemsetAsync(tr->d_oeq.p0, 0, (size_t)B * h->P * 4, s));
  if ((rc = run_dgrad(h, tr->bw_r1, d_r1v, 0, MS, d_oeqv, cp_off, s))) return rc;
  // ---- equalizer_ofdm ------------------------------------------------------------------------------
  TrainLayer* tl = tr->tl;
  Act d_catv = tr->d_cat;                       // [MS, 4K] = [d eq_out | d corr_out]
  Act d_eqv = tr->d_eq;   d_eqv.ld = 2 * K;
  Act d_corrv = tr->d_corr; d_corrv.ld = K;
  Act d_fv = tr->d_f;     d_fv.ld = 2 * K;
  Act d_t1v = tr->d_t1;   d_t1v.ld = 2 * K;
  // dense_5                                                                            model.py:457
  if ((rc = run_wgrad(h, tl[9], h->cat.p0, 4 * K, d_oeqv.p0, 2 * T, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[9].bw, d_oeqv, 0, MS, d_catv, 0, s))) return rc;
  // conv3d_3 on eq, conv3d_2 on corr (real rows only)                                  model.py:437-448
  if ((rc = run_wgrad(h, tl[8], h->eq.p0, 2 * K, d_catv.p0, 4 * K, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[8].bw, d_catv, 0, MS, d_eqv, 0, s))) return rc;
  if ((rc = run_wgrad(h, tl[7], h->corr.p0, K, d_catv.p0 + 2 * K, 4 * K, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[7].bw, d_catv, 2 * K, MS, d_corrv, 0, s))) return rc;
  // phase-only equaliser                                                               model.py:430-437
  {
    LaunchScope ls(h, SLOT_T_POINT, s);
    const long long total = (long long)B * S * K;
    phaseeq_bwd_kernel<<<blocks_for(total), 256, 0, s>>>((const float2*)tr->d_eq.p0, tr->d_corr.p0, (const float2*)h->f.p0,
                                                         (const float2*)h->chest_buf, (const float2*)h->eq.p0, total,
                                                         (float2*)tr->d_f.p0, (float2*)tr->d_ch.p0);
  }
  // conv3d_1 (Toeplitz), tanh, dense_4 .. dense_1                                      model.py:393-426
  if ((rc = run_wgrad(h, tl[6], h->u3.p0, SK2, tr->d_ch.p0, SK2, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[6].bw, tr->d_ch, 0, B, tr->dA, 0, s))) return rc;
  {
    LaunchScope ls(h, SLOT_T_POINT, s);
    tanh_bwd_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(tr->dA.p0, h->u3.p0, (long long)B * SK2);
  }
  if ((rc = run_wgrad(h, tl[5], h->u2.p0, SK2, tr->dA.p0, SK2, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[5].bw, tr->dA, 0, B, tr->dB, 0, s))) return rc;
  if (h->eqs.chain_act[1]) {   // equalizer_separateIQ: dense_3 has a tanh (model.py:1146-1150); its output is u2
    LaunchScope ls(h, SLOT_T_POINT, s);
    tanh_bwd_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(tr->dB.p0, h->u2.p0, (long long)B * SK2);
  }
  if ((rc = run_wgrad(h, tl[4], h->u1.p0, SK2, tr->dB.p0, SK2, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[4].bw, tr->dB, 0, B, tr->dC, 0, s))) return rc;
  if (h->eqs.chain_act[0]) {   // ... and dense_2 (model.py:1140-1144); its output is u1
    LaunchScope ls(h, SLOT_T_POINT, s);
    tanh_bwd_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(tr->dC.p0, h->u1.p0, (long long)B * SK2);
  }
  if ((rc = run_wgrad(h, tl[3], h->p32.p0, h->p32.ld, tr->dC.p0, SK2, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[3].bw, tr->dC, 0, B, tr->d_p32, 0, s))) return rc;
  if ((rc = run_wgrad(h, tl[2], h->f.p0, SK2, tr->d_p32.p0, tr->d_p32.ld, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[2].bw, tr->d_p32, 0, B, tr->d_f2, 0, s))) return rc;
  {
    LaunchScope ls(h, SLOT_T_POINT, s);
    add_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(tr->d_f.p0, tr->d_f2.p0, (long long)B * SK2);
  }
  // learned DFT (conv3d) and the per-symbol input dense                                 model.py:370-379
  if ((rc = run_wgrad(h, tl[1], h->t1.p0, 2 * K, tr->d_f.p0, 2 * K, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[1].bw, d_fv, 0, MS, d_t1v, 0, s))) return rc;
  if ((rc = run_wgrad(h, tl[0], h->a0.p0 + cp_off, 2 * T, tr->d_t1.p0, 2 * K, MS, s))) return rc;
  // ---- regulariser -------------------------------------------------------------------------------
  {
    LaunchScope ls(h, SLOT_T_ADAM, s, 12);
    for (TrainParam& p : tr->params)
      if (p.l2g != 0.f) add_l2_kernel<<<blocks_for(p.n), 256, 0, s>>>(p.g, p.w, p.n, p.l2g);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

// --opt 1..5: backward of ce_mean + reg through the generic equalizer wiring (run_chunk's generic path kept the chain
// outputs in u1, u2, u3 and the channel estimate -- after its tanh, if the last chain layer has one -- in chest_buf).
static int backward_generic(dccn_handle* h, int64_t B, const uint8_t* bits, cudaStream_t s) {
  TrainState* tr = h->tr;
  const EqSpec& sp = h->eqs;
  const int S = h->S, K = h->K, T = h->T, Tin = h->Tin, D = h->D, NB = h->NB, F = h->F;
  const int64_t MS = B * S;
  const int cp_off = (T - Tin) * 2;
  const int SK2 = S * K * 2;
  int rc;
  {
    LaunchScope ls(h, SLOT_T_HEAD, s);
    const long long total = (long long)B * D;
    const float inv_n = (float)(1.0 / ((double)B * D * NB));
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    switch (NB) {
      case 1: head_bwd_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      case 2: head_bwd_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      case 3: head_bwd_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
      default: head_bwd_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0); break;
    }
    DCCN_CUDA_OK(cudaGetLastError());
  }
  // frozen receiver: data gradients only
  if ((rc = run_dgrad(h, tr->bw_r2, tr->d_oiq, 0, B, tr->d_r1o, 0, s))) return rc;
  Act d_r1v = tr->d_r1o;  d_r1v.ld = 2 * F;
  Act d_oeqv = tr->d_oeq; d_oeqv.ld = 2 * T;
  if (cp_off) DCCN_CUDA_OK(cudaMemsetAsync(tr->d_oeq.p0, 0, (size_t)B * h->P * 4, s));
  if ((rc = run_dgrad(h, tr->bw_r1, d_r1v, 0, MS, d_oeqv, cp_off, s))) return rc;
  TrainLayer* tl = tr->tl;
  Act d_midv = tr->d_cat; d_midv.ld = 2 * K;       // [MS, 2K] inside the [MS, 4K] buffer
  Act d_eqv = tr->d_eq;   d_eqv.ld = 2 * K;
  Act d_fv = tr->d_f;     d_fv.ld = 2 * K;
  Act d_t1v = tr->d_t1;   d_t1v.ld = 2 * K;
  // tail: dense(2T) <- [dense(2K) | constant inverse DFT] <- eq
  if ((rc = run_wgrad(h, tl[tr->ix_tail2], h->cat.p0, 2 * K, d_oeqv.p0, 2 * T, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[tr->ix_tail2].bw, d_oeqv, 0, MS, d_midv, 0, s))) return rc;
  if (sp.tail == 1) {
    if ((rc = run_wgrad(h, tl[tr->ix_tail1], h->eq.p0, 2 * K, d_midv.p0, 2 * K, MS, s))) return rc;
    if ((rc = run_dgrad(h, tl[tr->ix_tail1].bw, d_midv, 0, MS, d_eqv, 0, s))) return rc;
  } else if ((rc = run_dgrad(h, tr->bw_ifft, d_midv, 0, MS, d_eqv, 0, s))) return rc;
  // phase-only equaliser (no correlation branch in these graphs)
  {
    LaunchScope ls(h, SLOT_T_POINT, s);
    const long long total = (long long)B * S * K;
    phaseeq_bwd_kernel<<<blocks_for(total), 256, 0, s>>>((const float2*)tr->d_eq.p0, nullptr, (const float2*)h->f.p0,
                                                         (const float2*)h->chest_buf, (const float2*)h->eq.p0, total,
                                                         (float2*)tr->d_f.p0, (float2*)tr->d_ch.p0);
  }
  const Act* chain_out[3] = {&h->u1, &h->u2, &h->u3};
  Act* ring[3] = {&tr->dA, &tr->dB, &tr->dC};
  int rp = 0;
  const Act* cur = &tr->d_ch;
  if (sp.toeplitz) {
    if ((rc = run_wgrad(h, tl[tr->ix_toep], chain_out[sp.n_chain - 1]->p0, SK2, cur->p0, SK2, B, s))) return rc;
    if ((rc = run_dgrad(h, tl[tr->ix_toep].bw, *cur, 0, B, *ring[rp], 0, s))) return rc;
    cur = ring[rp];
    rp = (rp + 1) % 3;
  }
  for (int i = sp.n_chain - 1; i >= 0; --i) {
    const bool fused = (i == sp.n_chain - 1) && !sp.toeplitz;          // this layer's output is the channel estimate
    if (sp.chain_act[i]) {
      LaunchScope ls(h, SLOT_T_POINT, s);
      tanh_bwd_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(cur->p0, fused ? h->chest_buf : chain_out[i]->p0,
                                                                     (long long)B * SK2);
    }
    const float* X = i == 0 ? h->p32.p0 : chain_out[i - 1]->p0;
    const int ldx = i == 0 ? h->p32.ld : SK2;
    if ((rc = run_wgrad(h, tl[tr->ix_chain[i]], X, ldx, cur->p0, SK2, B, s))) return rc;
    const Act* nxt = i == 0 ? &tr->d_p32 : ring[rp];
    if ((rc = run_dgrad(h, tl[tr->ix_chain[i]].bw, *cur, 0, B, *nxt, 0, s))) return rc;
    cur = nxt;
    if (i != 0) rp = (rp + 1) % 3;
  }
  // pilot bottleneck, then the two per-symbol front layers
  if ((rc = run_wgrad(h, tl[tr->ix_pilot], h->f.p0, SK2, tr->d_p32.p0, tr->d_p32.ld, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[tr->ix_pilot].bw, tr->d_p32, 0, B, tr->d_f2, 0, s))) return rc;
  {
    LaunchScope ls(h, SLOT_T_POINT, s);
    add_kernel<<<blocks_for((long long)B * SK2), 256, 0, s>>>(tr->d_f.p0, tr->d_f2.p0, (long long)B * SK2);
  }
  if ((rc = run_wgrad(h, tl[tr->ix_front2], h->t1.p0, 2 * K, tr->d_f.p0, 2 * K, MS, s))) return rc;
  if ((rc = run_dgrad(h, tl[tr->ix_front2].bw, d_fv, 0, MS, d_t1v, 0, s))) return rc;
  if ((rc = run_wgrad(h, tl[tr->ix_front1], h->a0.p0 + cp_off, 2 * T, tr->d_t1.p0, 2 * K, MS, s))) return rc;
  {
    LaunchScope ls(h, SLOT_T_ADAM, s, 12);
    for (TrainParam& p : tr->params)
      if (p.l2g != 0.f) add_l2_kernel<<<blocks_for(p.n), 256, 0, s>>>(p.g, p.w, p.n, p.l2g);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int NB>
static void launch_head_grads(dccn_handle* h, TrainState* tr, const uint8_t* bits, long long total, float inv_n,
                              unsigned blocks, cudaStream_t s) {
  constexpr int MO = 1 << NB;
  constexpr int NQ = 2 * MO + MO + (MO + 2) * 2 * NB + 2 * NB;
  head_bwd_kernel<NB><<<blocks, 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n, tr->d_oiq.p0);
  head_wgrad_kernel<NB><<<(unsigned)(tr->head_warps / 8), 256, 0, s>>>(h->out_iq.p0, bits, h->hw, total, inv_n,
                                                                      tr->head_partial);
  head_wgrad_reduce_kernel<<<blocks_for(NQ), 256, 0, s>>>(tr->head_partial, tr->head_warps, NQ, tr->params[tr->hp[0]].g,
                                                         2 * MO, tr->params[tr->hp[1]].g, MO, tr->params[tr->hp[2]].g,
                                                         (MO + 2) * 2 * NB, tr->params[tr->hp[3]].g);
}

// DCCN_TRAIN_RX: d total_loss / d (every ofdm_dense_rx variable)      dev/py/ofdmreceiver_np.py:154-189
static int backward_rx(dccn_handle* h, int64_t B, const uint8_t* bits, cudaStream_t s) {
  TrainState* tr = h->tr;
  const int S = h->S, T = h->T, Tin = h->Tin, D = h->D, NB = h->NB, F = h->F;
  const int64_t MS = B * S;
  const int cp_off = (T - Tin) * 2;
  int rc;
  {
    LaunchScope ls(h, SLOT_T_HEAD, s, 3);
    const long long total = (long long)B * D;
    const float inv_n = (float)(1.0 / ((double)B * D * NB));
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    switch (NB) {
      case 1: launch_head_grads<1>(h, tr, bits, total, inv_n, (unsigned)blocks, s); break;
      case 2: launch_head_grads<2>(h, tr, bits, total, inv_n, (unsigned)blocks, s); break;
      case 3: launch_head_grads<3>(h, tr, bits, total, inv_n, (unsigned)blocks, s); break;
      default: launch_head_grads<4>(h, tr, bits, total, inv_n, (unsigned)blocks, s); break;
    }
    DCCN_CUDA_OK(cudaGetLastError());
  }
  TrainLayer* tl = tr->tl;
  // demodulation/dense [S*F*2 -> 2D], contraction over the B frames                                  model.py:1269
  if ((rc = run_wgrad(h, tl[1], h->r1o.p0, S * F * 2, tr->d_oiq.p0, 2 * D, B, s))) return rc;
  if ((rc = run_dgrad(h, tl[1].bw, tr->d_oiq, 0, B, tr->d_r1o, 0, s))) return rc;
  // fft_like/conv3d: the live tap as a per-symbol [2Tin -> 2F] layer, contraction over the B*S symbols   model.py:1249
  if ((rc = run_wgrad(h, tl[0], h->a0.p0 + cp_off, 2 * T, tr->d_r1o.p0, 2 * F, MS, s))) return rc;
  // regulariser: berlin * REG_COEFF * l2 * sum(w^2) on the two tf.layers.dense (kernel + bias)           :162-173
  {
    LaunchScope ls(h, SLOT_T_ADAM, s, 4);
    for (TrainParam& p : tr->params)
      if (p.l2g != 0.f) add_l2_ber_kernel<<<blocks_for(p.n), 256, 0, s>>>(p.g, p.w, p.n, p.l2g, h->d_conf);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

// the head kernels take their 200 weights by value (dccn_handle::hw, host memory): fetch the updated variables.
// Synchronises the stream (a receiver-training step is not asynchronous).
static int refresh_head(dccn_handle* h, cudaStream_t s) {
  TrainState* tr = h->tr;
  const int NB = h->NB, MO = 1 << NB;
  size_t off[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    off[i + 1] = off[i] + (size_t)tr->params[tr->hp[i]].n;
    DCCN_CUDA_OK(cudaMemcpyAsync(tr->head_host + off[i], tr->params[tr->hp[i]].w, (size_t)tr->params[tr->hp[i]].n * 4,
                                 cudaMemcpyDeviceToHost, s));
  }
  DCCN_CUDA_OK(cudaStreamSynchronize(s));
  const float *kc = tr->head_host + off[0], *bc = tr->head_host + off[1], *k1 = tr->head_host + off[2],
              *b1 = tr->head_host + off[3];
  for (int i = 0; i < 2; ++i)
    for (int m = 0; m < MO; ++m) h->hw.Wc[i][m] = kc[i * MO + m];
  for (int m = 0; m < MO; ++m) h->hw.bc[m] = bc[m];
  for (int m = 0; m < MO + 2; ++m)
    for (int j = 0; j < 2 * NB; ++j) h->hw.W1[m][j] = k1[m * 2 * NB + j];
  for (int j = 0; j < 2 * NB; ++j) h->hw.b1[j] = b1[j];
  return 0;
}

}  // namespace dccn

using namespace dccn;

static int train_init_impl(dccn_handle* h, const dccn_train_cfg* cfg, void* stream) {
  const bool rx_mode = cfg->mode == DCCN_TRAIN_RX;
  DCCN_CHECK(cfg->mode == DCCN_TRAIN_EQ || cfg->mode == DCCN_TRAIN_RX, "unknown training mode %d", (int)cfg->mode);
  DCCN_CHECK(rx_mode || h->cfg.equalizer, "DCCN_TRAIN_EQ updates the Equalizer/* variables: the handle has no equalizer");
  DCCN_CHECK(!rx_mode || !h->cfg.equalizer, "DCCN_TRAIN_RX trains the basic receiver: create the handle without equalizer");
  DCCN_CHECK(h->committed, "weights not committed (dccn_commit_weights)");
  DCCN_CHECK(h->cfg.precision == DCCN_PREC_EXACT || h->cfg.precision == DCCN_PREC_PARITY,
             "training needs fp32-class arithmetic (precision exact or parity)");
  DCCN_CHECK(!h->fused_head, "training needs the stored out_iq (DCCN_FUSED_HEAD=0)");
  DCCN_CHECK(h->cfg.head == DCCN_HEAD_DEV, "training is defined for the dev head (dev/py/model.py:1275-1288)");
  DCCN_CHECK(cfg->max_batch > 0 && cfg->max_batch <= h->chunk, "max_batch must be in 1..chunk_frames (%d)", h->chunk);
  cudaStream_t s = (cudaStream_t)stream;
  const int S = h->S, K = h->K, T = h->T, F = h->F, D = h->D;
  const int64_t MB = cfg->max_batch;
  int rc = 0;
  TrainState* tr = new TrainState();
  h->tr = tr;
  tr->cfg = *cfg;
  tr->maxB = MB;
  tr->mode = cfg->mode;
  const bool generic = !rx_mode && h->eqs.generic;
  std::vector<GemmLayer*> Ls;
  std::vector<std::string> layer_var;
  std::vector<int> bias_kind, per_symbol;
  if (rx_mode) {
    Ls = {&h->r1, &h->r2};
    for (int i = 0; i < 2; ++i) {
      layer_var.push_back(kRxLayerVar[i]);
      bias_kind.push_back(kRxBiasKind[i]);
      per_symbol.push_back(kRxPerSymbol[i]);
    }
  } else if (!generic) {
    Ls = {&h->g1, &h->g2, &h->g3, &h->g4, &h->g5, &h->g6, &h->g7, &h->g8, &h->g9, &h->g10};
    for (int i = 0; i < 10; ++i) {
      layer_var.push_back(kLayerVar[i]);
      // --opt 7: the four conv3d layers are layers_conv2d_vector (bias kinds 3 / 4 instead of 1 / 2)
      bias_kind.push_back(h->eqs.vector && kBiasKind[i] ? kBiasKind[i] + 2 : kBiasKind[i]);
      per_symbol.push_back(kPerSymbol[i]);
    }
  } else {
    // the trainable layers of the --opt graph in creation order, named by TF-1's per-scope auto-numbering
    const EqSpec& sp = h->eqs;
    int nd = 0, ncv = 0;
    auto add = [&](GemmLayer* L, bool conv, int kind, int persym) {
      int& ctr = conv ? ncv : nd;
      layer_var.push_back(std::string("Equalizer/") + (conv ? "conv3d" : "dense") + (ctr == 0 ? "" : "_" + std::to_string(ctr)));
      ++ctr;
      Ls.push_back(L);
      bias_kind.push_back(kind);
      per_symbol.push_back(persym);
      return (int)Ls.size() - 1;
    };
    tr->ix_front1 = add(&h->g1, false, 0, 1);
    tr->ix_front2 = sp.front2_cconv ? add(&h->g2, true, 1, 1) : add(&h->g2, false, 0, 1);
    tr->ix_pilot = add(&h->g3, false, 0, 0);
    GemmLayer* chain[4] = {&h->g4, &h->g5, &h->g6, &h->gx0};
    for (int i = 0; i < sp.n_chain; ++i) tr->ix_chain[i] = add(chain[i], false, 0, 0);
    if (sp.toeplitz) tr->ix_toep = add(&h->g7, true, 2, 0);
    if (sp.tail == 1) tr->ix_tail1 = add(&h->g9, false, 0, 1);
    tr->ix_tail2 = add(&h->g10, false, 0, 1);
  }
  const int NL = tr->n_layers = (int)Ls.size();
  DCCN_CHECK(NL <= 10, "too many trainable layers");
  if (!rx_mode && !h->ws_train) h->ws_train = h->ws_dirty = true;   // the backward pass needs u3 (tanh output) and chest
  if ((rc = ensure_workspace(h, MB))) return rc;
  // ---- gather maps of the layers: run the host packers on index-valued variables ------------------------
  std::vector<std::vector<float>> backup(NL);
  for (int i = 0; i < NL; ++i) {
    HostTensor& t = h->raw[layer_var[i] + "/kernel"];
    DCCN_CHECK(t.data.size() < (1u << 24), "variable too large for the index trick");
    backup[i] = t.data;
    for (size_t p = 0; p < t.data.size(); ++p) t.data[p] = (float)(p + 1);
  }
  rc = pack_layers_host(h);
  std::vector<std::vector<int32_t>> maps(NL);
  if (!rc)
    for (int i = 0; i < NL; ++i) {
      maps[i].resize(Ls[i]->W.size());
      for (size_t e = 0; e < maps[i].size(); ++e) maps[i][e] = (int32_t)Ls[i]->W[e];
    }
  for (int i = 0; i < NL; ++i) h->raw[layer_var[i] + "/kernel"].data = backup[i];
  if (rc) return rc;
  if ((rc = pack_layers_host(h))) return rc;
  // ---- variables, optimiser slots, maps, operands ----------------------------------------------------------
  const float l2g = 2.0f * cfg->reg_coeff * cfg->l2;
  tr->params.reserve(24);
  size_t max_partial = 0;
  for (int i = 0; i < NL; ++i) {
    TrainLayer& t = tr->tl[i];
    t.L = Ls[i];
    t.bias_kind = bias_kind[i];
    const bool dense = bias_kind[i] == 0;
    for (int kb = 0; kb < 2; ++kb) {
      TrainParam p;
      p.name = layer_var[i] + (kb == 0 ? "/kernel" : "/bias");
      const HostTensor* ht = find(h, p.name);
      DCCN_CHECK(ht, "weight '%s' was not set", p.name.c_str());
      p.n = (int64_t)ht->data.size();
      p.l2g = dense ? l2g : 0.f;
      rc |= dev_alloc(h, (void**)&p.w, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.m, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.v, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.g, (size_t)p.n * 4);
      if (rc) return rc;
      DCCN_CUDA_OK(cudaMemsetAsync(p.m, 0, (size_t)p.n * 4, s));
      DCCN_CUDA_OK(cudaMemsetAsync(p.v, 0, (size_t)p.n * 4, s));
      DCCN_CUDA_OK(cudaMemsetAsync(p.g, 0, (size_t)p.n * 4, s));
      (kb == 0 ? t.kp : t.bp) = (int)tr->params.size();
      tr->params.push_back(p);
    }
    const int Kl = t.L->K, Nl = t.L->N;
    const size_t kn = (size_t)Kl * Nl;
    const int64_t n_param = tr->params[t.kp].n;
    // CSR: parameter -> signed operand positions
    std::vector<int32_t> off(n_param + 1, 0), idx;
    for (size_t e = 0; e < kn; ++e)
      if (maps[i][e]) off[std::abs(maps[i][e])]++;            // count at p+1
    for (int64_t p = 0; p < n_param; ++p) {
      if (off[p + 1] > t.max_list) t.max_list = off[p + 1];
      off[p + 1] += off[p];
    }
    idx.resize(off[n_param] > 0 ? off[n_param] : 1);
    std::vector<int32_t> cur(off.begin(), off.end() - 1);
    for (size_t e = 0; e < kn; ++e) {
      const int32_t mv = maps[i][e];
      if (!mv) continue;
      const int32_t p = std::abs(mv) - 1;
      idx[cur[p]++] = mv > 0 ? (int32_t)(e + 1) : -(int32_t)(e + 1);
    }
    rc |= dev_alloc(h, (void**)&t.map, kn * 4);
    rc |= dev_alloc(h, (void**)&t.csr_off, off.size() * 4);
    rc |= dev_alloc(h, (void**)&t.csr_idx, idx.size() * 4);
    rc |= dev_alloc(h, (void**)&t.dWp, (kn + Nl) * 4);
    if (h->cfg.precision == DCCN_PREC_EXACT) t.Wp = t.L->dW;
    else rc |= dev_alloc(h, (void**)&t.Wp, kn * 4);
    if (rc) return rc;
    DCCN_CUDA_OK(cudaMemcpyAsync(t.map, maps[i].data(), kn * 4, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaMemcpyAsync(t.csr_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaMemcpyAsync(t.csr_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaStreamSynchronize(s));
    if ((rc = make_bw_layer(h, *t.L, &t.bw, s))) return rc;
    // wgrad scratch: splits x (K+1) x N for the batch this layer contracts over
    const int64_t M = per_symbol[i] ? MB * S : MB;
    const int rps = rows_per_split(M);
    size_t need = (size_t)((M + rps - 1) / rps) * (kn + Nl);
    if (need > max_partial) max_partial = need;
    // tensor-core wgrad: split-K slices of padded [m_tiles*128, N] + the transposed operands
    // (split-K grows as the batch shrinks only up to num_kb/2, so the largest batch bounds the scratch)
    need = (size_t)wgrad_ksplit(Kl, Nl, M, h->num_sms) * ((Kl + 127) / 128) * 128 * Nl;
    if (need > max_partial) max_partial = need;
    const size_t ldt = (size_t)((M + 3) & ~3LL);
    if ((size_t)Kl * ldt > tr->xT_floats) tr->xT_floats = (size_t)Kl * ldt;
    if ((size_t)Nl * ldt > tr->yT_floats) tr->yT_floats = (size_t)Nl * ldt;
  }
  tr->wgrad_tc = h->cfg.precision == DCCN_PREC_PARITY;
  if (const char* e = getenv("DCCN_WGRAD_SIMT")) if (atoi(e)) tr->wgrad_tc = 0;
  if (tr->wgrad_tc) {
    rc |= dev_alloc(h, (void**)&tr->xT, tr->xT_floats * 4);
    rc |= dev_alloc(h, (void**)&tr->yT0, tr->yT_floats * 4);
    rc |= dev_alloc(h, (void**)&tr->yT1, tr->yT_floats * 4);
    if (rc) return rc;
  }
  if (!rx_mode) {
    if ((rc = make_bw_layer(h, h->r1, &tr->bw_r1, s))) return rc;
    if ((rc = make_bw_layer(h, h->r2, &tr->bw_r2, s))) return rc;
    if (generic && h->eqs.tail == 2 && (rc = make_bw_layer(h, h->g9, &tr->bw_ifft, s))) return rc;
  }
  tr->partial_floats = max_partial;
  rc |= dev_alloc(h, (void**)&tr->partial, max_partial * 4);
  if (rx_mode) {
    // the head's four variables: plain parameters (no packed operand), l2 on dense_1 only
    const int MO = 1 << h->NB;
    const int NQ = 2 * MO + MO + (MO + 2) * 2 * h->NB + 2 * h->NB;
    for (int i = 0; i < 4; ++i) {
      TrainParam p;
      p.name = kRxHeadVar[i];
      const HostTensor* ht = find(h, p.name);
      DCCN_CHECK(ht, "weight '%s' was not set", p.name.c_str());
      p.n = (int64_t)ht->data.size();
      p.l2g = i >= 2 ? l2g : 0.f;
      rc |= dev_alloc(h, (void**)&p.w, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.m, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.v, (size_t)p.n * 4);
      rc |= dev_alloc(h, (void**)&p.g, (size_t)p.n * 4);
      if (rc) return rc;
      DCCN_CUDA_OK(cudaMemsetAsync(p.m, 0, (size_t)p.n * 4, s));
      DCCN_CUDA_OK(cudaMemsetAsync(p.v, 0, (size_t)p.n * 4, s));
      DCCN_CUDA_OK(cudaMemsetAsync(p.g, 0, (size_t)p.n * 4, s));
      tr->hp[i] = (int)tr->params.size();
      tr->params.push_back(p);
    }
    DCCN_CHECK(tr->params[tr->hp[0]].n == 2 * MO && tr->params[tr->hp[1]].n == MO &&
                   tr->params[tr->hp[2]].n == (MO + 2) * 2 * h->NB && tr->params[tr->hp[3]].n == 2 * h->NB,
               "head variables do not have the dev-head shapes");
    tr->head_warps = h->num_sms * 8 * 8;                      // 8 blocks of 8 warps per SM
    rc |= dev_alloc(h, (void**)&tr->head_partial, (size_t)tr->head_warps * NQ * 4);
    if (cudaMallocHost((void**)&tr->head_host, (size_t)NQ * 4) != cudaSuccess) rc = set_error(-1, "cudaMallocHost failed");
    if (rc) return rc;
  }
  // ---- gradient activations -------------------------------------------------------------------------------
  const int SK2 = S * K * 2;
  rc |= alloc_act(h, &tr->d_oiq, MB, 2 * D, false);
  rc |= alloc_act(h, &tr->d_r1o, MB, S * F * 2, false);
  if (rx_mode) {
    if (rc) return rc;
    if ((rc = upload_params(h, s))) return rc;
    return repack_all(h, s);
  }
  rc |= alloc_act(h, &tr->d_oeq, MB, S * T * 2, false);
  rc |= alloc_act(h, &tr->d_cat, MB * S, 4 * K, false);
  rc |= alloc_act(h, &tr->d_eq, MB, SK2, false);
  rc |= alloc_act(h, &tr->d_corr, MB, S * K, false);
  rc |= alloc_act(h, &tr->d_f, MB, SK2, false);
  rc |= alloc_act(h, &tr->d_f2, MB, SK2, false);
  rc |= alloc_act(h, &tr->d_ch, MB, SK2, false);
  rc |= alloc_act(h, &tr->dA, MB, SK2, false);
  rc |= alloc_act(h, &tr->dB, MB, SK2, false);
  rc |= alloc_act(h, &tr->dC, MB, SK2, false);
  rc |= alloc_act(h, &tr->d_p32, MB, 2 * h->cfg.pilot_size, false);
  rc |= alloc_act(h, &tr->d_t1, MB, SK2, false);
  if (rc) return rc;
  if ((rc = upload_params(h, s))) return rc;
  return repack_all(h, s);
}

extern "C" {

int dccn_train_init(dccn_handle* h, const dccn_train_cfg* cfg, void* stream) {
  DCCN_CHECK(h && cfg, "null argument");
  DCCN_CHECK(!h->tr, "training state already initialised");
  const int rc = train_init_impl(h, cfg, stream);
  if (rc) train_free(h);   // a half-built state must not be used
  return rc;
}

int dccn_train_step(dccn_handle* h, const float* x_dev, int64_t B, const uint8_t* bits_dev, float learning_rate,
                    int apply_update, int64_t* conf_dev, double* ce_sum_dev, int flags, void* stream) {
  DCCN_CHECK(h && x_dev && bits_dev, "null argument");
  DCCN_CHECK(h->tr, "dccn_train_init was not called");
  DCCN_CHECK(B > 0 && B <= h->tr->maxB, "batch %lld outside 1..max_batch (%lld)", (long long)B, (long long)h->tr->maxB);
  DCCN_CHECK(!(flags & (DCCN_FWD_EQ_ONLY | DCCN_FWD_SKIP_EQ)), "a training step runs the whole graph");
  cudaStream_t s = (cudaStream_t)stream;
  TrainState* tr = h->tr;
  const bool rx_mode = tr->mode == DCCN_TRAIN_RX;
  int rc = 0;
  // ---- forward, keeping every activation ------------------------------------------------------------------
  if (!(flags & DCCN_FWD_NO_NORM) && (rc = run_moments(h, x_dev, B, h->d_mean, h->d_rstd, s))) return rc;
  // (the receiver's loss needs this minibatch's confusion matrix: berlin scales the regulariser)
  if (conf_dev || rx_mode) DCCN_CUDA_OK(cudaMemsetAsync(h->d_conf, 0, 4 * sizeof(unsigned long long), s));
  h->train_fwd = !rx_mode;
  rc = run_chunk(h, x_dev, B, bits_dev, nullptr, nullptr, nullptr, nullptr, (conf_dev || rx_mode) ? h->d_conf : nullptr,
                 ce_sum_dev, flags, s);
  h->train_fwd = false;
  if (rc) return rc;
  if (conf_dev) conf_accumulate(h->d_conf, conf_dev, s);
  // ---- backward ------------------------------------------------------------------------------------------------
  if ((rc = rx_mode ? backward_rx(h, B, bits_dev, s)
                    : (h->eqs.generic ? backward_generic(h, B, bits_dev, s) : backward(h, B, bits_dev, s))))
    return rc;
  if (!apply_update) return 0;
  // ---- Adam (dev/py/ofdmreceiver_np_mp.py:345-347) ------------------------------------------------------------
  tr->step += 1;
  tr->adam_t += 1;
  const double b1 = tr->cfg.beta1, b2 = tr->cfg.beta2;
  const float lr_t = (float)((double)learning_rate * std::sqrt(1.0 - std::pow(b2, (double)tr->adam_t)) /
                             (1.0 - std::pow(b1, (double)tr->adam_t)));
  {
    LaunchScope ls(h, SLOT_T_ADAM, s, (int)tr->params.size());
    for (TrainParam& p : tr->params)
      adam_kernel<<<blocks_for(p.n), 256, 0, s>>>(p.w, p.m, p.v, p.g, p.n, lr_t, tr->cfg.beta1, tr->cfg.beta2, tr->cfg.eps);
    DCCN_CUDA_OK(cudaGetLastError());
  }
  if ((rc = repack_all(h, s))) return rc;
  return rx_mode ? refresh_head(h, s) : 0;
}

int64_t dccn_train_get_grad(dccn_handle* h, const char* tf_name, float* host, int64_t capacity) {
  if (!h || !tf_name || !h->tr) return set_error(-2, "bad argument / training not initialised");
  for (TrainParam& p : h->tr->params)
    if (p.name == tf_name) {
      if (host) {
        if (capacity < p.n) return set_error(-2, "buffer too small for the gradient of '%s'", tf_name);
        if (cudaMemcpy(host, p.g, (size_t)p.n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
          return set_error(-1, "cudaMemcpy failed");
      }
      return p.n;
    }
  return set_error(-2, "'%s' is not a trainable variable", tf_name);
}

int64_t dccn_train_global_step(const dccn_handle* h) { return (h && h->tr) ? h->tr->step : -1; }

int dccn_train_set_global_step(dccn_handle* h, int64_t step) {
  DCCN_CHECK(h && h->tr && step >= 0, "bad argument / training not initialised");
  h->tr->step = step;
  return 0;
}

}  // extern "C"
