// kernels.cuh -- the HBM-bound kernels around the GEMM stack:
//   batch moments (a2), per-frame normalisation/layer-norm + tf32 split (a2 + model.py:363),
//   Rayleigh FIR (warp-shuffle) + AWGN (a6, a7), OFDM transmitter, BER accumulation (a5),
//   and the op-level layers_conv2d_complex (a1).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "epilogue.cuh"

namespace dccn {

// =====================================================================================
// a2: tf.nn.moments(x, [0])  -- per-position sum / sum-of-squares over the batch axis.
// x [B, P] fp32 (P = S*T*2, multiple of 4).  grid = (ceil(P/4/128), Gy); each thread owns
// 4 adjacent positions and strides over frames; fp64 accumulation, one fp64 atomic per
// position per CTA row-group.  sums[0:P] = sum x, sums[P:2P] = sum x^2.
// =====================================================================================
__global__ void __launch_bounds__(128) moments_partial_kernel(const float* __restrict__ x, long long B, int P,
                                                              double* __restrict__ sums) {
  const int p4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (p4 * 4 >= P) return;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  auto acc = [&](const float4& v) {
    s[0] += v.x; q[0] += (double)v.x * v.x;
    s[1] += v.y; q[1] += (double)v.y * v.y;
    s[2] += v.z; q[2] += (double)v.z * v.z;
    s[3] += v.w; q[3] += (double)v.w * v.w;
  };
  // four frames in flight per thread: with one 16-byte load per iteration the kernel had ~16 KB outstanding per SM, a third of
  // what the HBM latency x bandwidth product asks for (measured 4.1 TB/s)
  const long long gy = gridDim.y;
  long long b = blockIdx.y;
  for (; b + 3 * gy < B; b += 4 * gy) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * P) + p4);
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b + gy) * P) + p4);
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b + 2 * gy) * P) + p4);
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b + 3 * gy) * P) + p4);
    acc(v0); acc(v1); acc(v2); acc(v3);
  }
  for (; b < B; b += gy) acc(__ldg(reinterpret_cast<const float4*>(x + (size_t)b * P) + p4));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(&sums[p4 * 4 + i], s[i]);
    atomicAdd(&sums[P + p4 * 4 + i], q[i]);
  }
}

// mean / rstd in fp32 like the TF graph: rstd = rsqrt(var + 1e-9)
__global__ void moments_final_kernel(const double* __restrict__ sums, long long B, int P, float* __restrict__ mean,
                                     float* __restrict__ rstd) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double m = sums[p] / (double)B;
  double var = sums[P + p] / (double)B - m * m;
  if (var < 0) var = 0;
  mean[p] = (float)m;
  rstd[p] = (float)(1.0 / sqrt(var + 1e-9));
}

// =====================================================================================
// prep: one warp per frame.
//   z  = (x*rstd + (-mean*rstd)) / sqrt(2)          ofdmreceiver_np.py:129  (TF op order, no FMA)
//   y  = layer_norm(z) over the whole frame          model.py:363 (only in front of the equalizer)
// and stores y (or z) as activation planes (fp32, or tf32 hi/lo for the 3xTF32 GEMMs).
// =====================================================================================
template <int MAXV>   // MAXV float4 per lane: P <= MAXV*128
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ x, long long B, int P,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   int do_norm, int do_ln, ActOut out, unsigned* __restrict__ amax) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int nv = P >> 2;
  const float4* xp = reinterpret_cast<const float4*>(x + (size_t)warp * P);
  float4 z[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 v = __ldg(xp + idx);
      float4 t = v;
      if (do_norm) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + idx);
      const float4 r = __ldg(reinterpret_cast<const float4*>(rstd) + idx);
      t.x = __fdiv_rn(__fadd_rn(__fmul_rn(v.x, r.x), __fmul_rn(-m.x, r.x)), 1.41421356237f);
      t.y = __fdiv_rn(__fadd_rn(__fmul_rn(v.y, r.y), __fmul_rn(-m.y, r.y)), 1.41421356237f);
      t.z = __fdiv_rn(__fadd_rn(__fmul_rn(v.z, r.z), __fmul_rn(-m.z, r.z)), 1.41421356237f);
      t.w = __fdiv_rn(__fadd_rn(__fmul_rn(v.w, r.w), __fmul_rn(-m.w, r.w)), 1.41421356237f);
      }
      z[i] = t;
      sum += (t.x + t.y) + (t.z + t.w);
    } else {
      z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (do_ln) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mu = sum / (float)P;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nv) {
        const float a = z[i].x - mu, b = z[i].y - mu, c = z[i].z - mu, d = z[i].w - mu;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float inv = (float)(1.0 / sqrt((double)(sq / (float)P) + 1e-12));
    const float sh = __fmul_rn(-mu, inv);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      z[i].x = __fadd_rn(__fmul_rn(z[i].x, inv), sh);
      z[i].y = __fadd_rn(__fmul_rn(z[i].y, inv), sh);
      z[i].z = __fadd_rn(__fmul_rn(z[i].z, inv), sh);
      z[i].w = __fadd_rn(__fmul_rn(z[i].w, inv), sh);
    }
  }
  unsigned mx = 0u;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      float y[4] = {z[i].x, z[i].y, z[i].z, z[i].w};
      store_act<4>(out, warp, idx * 4, y);
#pragma unroll
      for (int j = 0; j < 4; ++j) mx = max(mx, __float_as_uint(y[j]) & 0x7FFFFFFFu);
    }
  }
  if (amax) {   // max |y| of the pass: operand scale of the fp16 hi/lo GEMM that reads this buffer (amax_update_warp)
    mx = __reduce_max_sync(0xffffffffu, mx);
    // one RED per row at ONE address would serialise in its L2 slice (4.6e5 rows per pass): look first (an L2 read that
    // any number of warps share), fire only when this row raises the maximum -- a handful of times per pass
    if (lane == 0 && mx > *reinterpret_cast<volatile unsigned*>(amax)) atomicMax(amax, mx);
  }
}

// =====================================================================================
// a6: static Rayleigh FIR, one warp per frame, neighbours exchanged with warp shuffles.
//   g  = (z * ch_coeff) @ alpha                                 radio.py:433-435
//   rx = np.convolve(tx, g, 'same')   (centred, zero history)    radio.py:436
// complex128 arithmetic like NumPy, result rounded to complex64 (radio.py:492).  Also
// accumulates the batch power sum_b,n |rx|^2 that AWGN_channel_np normalises by.
// =====================================================================================
constexpr int kMaxFir = 32;

// acc += g * x (complex128) with a FIXED contraction, shared by every FIR kernel so that the fused feeder
// (tx_fade_kernel) and the stand-alone chan_fir_kernel round identically
DCCN_DEVINL void fir_cmac(double& ar, double& ai, const double2 g, const double xr, const double xi) {
  ar = fma(-g.y, xi, fma(g.x, xr, ar));
  ai = fma(g.y, xr, fma(g.x, xi, ai));
}

__global__ void __launch_bounds__(256) chan_fir_kernel(const float2* __restrict__ tx, long long B, int n_samp,
                                                       const double* __restrict__ alpha,
                                                       const double* __restrict__ coeff, int n_taps, int n_fir,
                                                       const double* __restrict__ z_in, uint64_t seed,
                                                       long long frame0, long long fstride,
                                                       float2* __restrict__ rx, double* __restrict__ power_sum) {
  __shared__ double2 gsm[8][kMaxFir];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long frame = frame0 + ((long long)blockIdx.x * 8 + wib) * fstride;
  if (frame >= B) return;

  // path gains -> sample-spaced FIR g[j] (lane j owns tap j)
  double2 a = make_double2(0.0, 0.0);
  if (n_taps == 0) {
    if (lane == 0) gsm[wib][0] = make_double2(1.0, 0.0);
  } else {
    if (lane < n_taps) {
      double zr, zi;
      if (z_in) {
        zr = z_in[((size_t)frame * n_taps + lane) * 2];
        zi = z_in[((size_t)frame * n_taps + lane) * 2 + 1];
      } else {
        uint32_t r[4];
        Philox{seed}((uint64_t)frame * 64 + lane, 0xA11CEu, r);
        float n0, n1;
        box_muller(r[0], r[1], n0, n1);
        zr = (double)n0 * 0.70710678118654752;
        zi = (double)n1 * 0.70710678118654752;
      }
      const double c = coeff[lane];
      a = make_double2(zr * c, zi * c);
    }
    double2 g = make_double2(0.0, 0.0);
    for (int t = 0; t < n_taps; ++t) {
      const double ar = __shfl_sync(0xffffffffu, a.x, t);
      const double ai = __shfl_sync(0xffffffffu, a.y, t);
      if (lane < n_fir) {
        const double al = alpha ? alpha[t * n_fir + lane] : 1.0;
        g.x = fma(ar, al, g.x);
        g.y = fma(ai, al, g.y);
      }
    }
    if (lane < n_fir) gsm[wib][lane] = g;
  }
  __syncwarp();
  const int M = n_taps == 0 ? 1 : n_fir;
  const int off = (M - 1) - (M >> 1);         // np.convolve 'same': full[n + off]
  const float2* txf = tx + (size_t)frame * n_samp;
  float2* rxf = rx + (size_t)frame * n_samp;
  double pw = 0.0;
  float2 prev = make_float2(0.f, 0.f);
  float2 cur = lane < n_samp ? __ldg(txf + lane) : make_float2(0.f, 0.f);
  for (int base = 0; base < n_samp; base += 32) {
    const int nidx = base + 32 + lane;
    const float2 next = nidx < n_samp ? __ldg(txf + nidx) : make_float2(0.f, 0.f);
    double accr = 0.0, acci = 0.0;
    for (int j = 0; j < M; ++j) {
      const int dlt = off - j;                 // need tx[base + lane + dlt]  (dlt is warp-uniform)
      // the lane that will be READ decides what it supplies: its sample of the previous /
      // current / next 32-block (the requesting ranges are disjoint)
      const float2 supply = dlt < 0 ? (lane >= 32 + dlt ? prev : cur) : (lane < dlt ? next : cur);
      const int sl = (lane + dlt) & 31;
      const double xr = __shfl_sync(0xffffffffu, supply.x, sl);
      const double xi = __shfl_sync(0xffffffffu, supply.y, sl);
      fir_cmac(accr, acci, gsm[wib][j], xr, xi);
    }
    if (base + lane < n_samp) {
      const float2 o = make_float2((float)accr, (float)acci);
      rxf[base + lane] = o;
      pw += (double)o.x * o.x + (double)o.y * o.y;
    }
    prev = cur;
    cur = next;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
  if (lane == 0) atomicAdd(power_sum, pw);
}

// =====================================================================================
// Mobile (Doppler) fading, dev/py/radio.py:387-422: per frame 48 sinusoids per path with random
// phases (sum-of-sinusoids Jakes model), path gains re-evaluated at every OFDM symbol, and a
// per-symbol 'same' convolution over [symbol + n_taps samples of history] (no look-ahead past the
// symbol end, zero history before the frame).  One warp per frame; lanes share the sinusoids
// (warp reduction per path), then the FIR taps; float64 like NumPy, output rounded to complex64.
//   frames handled: frame0 + idx*fstride  (profile cycling of 'mixRayleigh', radio.py:450-467)
//   theta_in [B, 2, ss, n_taps] uniform(0,2pi) phases or nullptr -> Philox
// =====================================================================================
constexpr int kMaxPaths = 16;
constexpr int kSinusoids = 48;

__global__ void __launch_bounds__(256) chan_doppler_kernel(const float2* __restrict__ tx, long long B, int n_sym,
                                                           int n_sc, const double* __restrict__ alpha,
                                                           const double* __restrict__ coeff, int n_taps, int n_fir,
                                                           double Fd, double t_sym, const double* __restrict__ theta_in,
                                                           uint64_t seed, long long frame0, long long fstride,
                                                           float2* __restrict__ rx, double* __restrict__ power_sum) {
  __shared__ double2 gsm[8][kMaxFir];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long frame = frame0 + ((long long)blockIdx.x * 8 + wib) * fstride;
  if (frame >= B) return;
  const double kPi = 3.14159265358979323846;
  // sinusoid frequencies / phases of this lane: sinusoids n = lane and lane + 32 (< 48)
  double f_re[2][kMaxPaths], f_im[2][kMaxPaths], th_re[2][kMaxPaths], th_im[2][kMaxPaths];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = lane + 32 * h;
    for (int t = 0; t < n_taps; ++t) {
      if (n < kSinusoids) {
        const double nv = ((double)(n + 1) - 0.5) * kPi / (4.0 * kSinusoids);
        const double a0 = (double)(t + 1) * kPi / (4.0 * kSinusoids);
        f_re[h][t] = Fd * cos(nv + a0);
        f_im[h][t] = Fd * cos(nv - a0);
        if (theta_in) {
          const double* th = theta_in + (size_t)frame * 2 * kSinusoids * n_taps;
          th_re[h][t] = th[(size_t)n * n_taps + t];
          th_im[h][t] = th[(size_t)(kSinusoids + n) * n_taps + t];
        } else {
          uint32_t r[4];
          Philox{seed}((uint64_t)frame * 1024 + (uint64_t)n * kMaxPaths + t, 0xD099u, r);
          th_re[h][t] = 2.0 * kPi * ((double)r[0] * 2.3283064365386963e-10);
          th_im[h][t] = 2.0 * kPi * ((double)r[1] * 2.3283064365386963e-10);
        }
      } else {
        f_re[h][t] = f_im[h][t] = th_re[h][t] = th_im[h][t] = 0.0;
      }
    }
  }
  const double const1 = sqrt(1.0 / kSinusoids);
  const int M = n_fir;
  const int off = (M - 1) - (M >> 1);
  const float2* txf = tx + (size_t)frame * n_sym * n_sc;
  float2* rxf = rx + (size_t)frame * n_sym * n_sc;
  double pw = 0.0;
  for (int i = 0; i < n_sym; ++i) {
    const double tw = 2.0 * kPi * ((double)i * t_sym);
    double2 g = make_double2(0.0, 0.0);
    for (int t = 0; t < n_taps; ++t) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lane + 32 * h < kSinusoids) {
          sr += cos(tw * f_re[h][t] + th_re[h][t]);
          si += cos(tw * f_im[h][t] + th_im[h][t]);
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      const double c = coeff[t];
      const double ar = const1 * sr * c, ai = const1 * si * c;       // zck * ch_coeff
      if (lane < n_fir) {
        const double al = alpha ? alpha[t * n_fir + lane] : 1.0;
        g.x = fma(ar, al, g.x);
        g.y = fma(ai, al, g.y);
      }
    }
    __syncwarp();
    if (lane < n_fir) gsm[wib][lane] = g;
    __syncwarp();
    // window of this symbol: local index m in [-n_taps, n_sc) <-> frame sample i*n_sc + m (zero before the frame)
    for (int base = 0; base < n_sc; base += 32) {
      const int q = base + lane;
      if (q < n_sc) {
        double accr = 0.0, acci = 0.0;
        for (int j = 0; j < M; ++j) {
          const int m = q + off - j;
          const long long gi = (long long)i * n_sc + m;
          if (m >= -n_taps && m < n_sc && gi >= 0) {
            const float2 x = __ldg(txf + gi);
            const double2 gg = gsm[wib][j];
            accr += gg.x * x.x - gg.y * x.y;
            acci += gg.x * x.y + gg.y * x.x;
          }
        }
        const float2 o = make_float2((float)accr, (float)acci);
        rxf[(size_t)i * n_sc + q] = o;
        pw += (double)o.x * o.x + (double)o.y * o.y;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
  if (lane == 0) atomicAdd(power_sum, pw);
}

// =====================================================================================
// a7: AWGN_channel_np -- x / sqrt(mean power of the whole batch) + N(0,1)*sqrt(.5)*10^(-SNR/20)
// float64 arithmetic like the reference, output rounded to fp32 (the TF feed dtype).
// One warp per frame (grid-stride over frames): the per-frame noise deviation -- a float64 pow -- and the batch scale are
// evaluated once per frame instead of once per sample, and the flat index needs no 64-bit divide; a thread handles two
// adjacent complex samples (one 16-byte load / store) with ONE Philox-4x32 call (4 uniforms -> 2 Box-Muller pairs).
// Round 1 did pow() + a 64-bit divide + a half-used Philox call per sample: 17 % of the HBM roofline.
// =====================================================================================
__global__ void __launch_bounds__(256) awgn_kernel(const float2* __restrict__ xin, long long B, int n_samp,
                                                   const double* __restrict__ power_sum,
                                                   const float* __restrict__ snr_db,
                                                   const double* __restrict__ normals, uint64_t seed,
                                                   float2* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double inv = 1.0 / sqrt(*power_sum / (double)(B * n_samp));
  const int npair = n_samp >> 1;                       // n_samp is even for every frame geometry (S * T, T even)
  for (long long b = warp0; b < B; b += nwarps) {
    double std_b = 0.0;
    if (lane == 0) std_b = sqrt(0.5) * pow(10.0, -(double)__ldg(snr_db + b) / 20.0);
    std_b = __shfl_sync(0xffffffffu, std_b, 0);
    const float4* xi = reinterpret_cast<const float4*>(xin + b * n_samp);
    float4* xo = reinterpret_cast<float4*>(out + b * n_samp);
    for (int p = lane; p < npair; p += 32) {
      const float4 v = __ldg(xi + p);
      double n0, n1, n2, n3;
      if (normals) {
        const double2* np_ = reinterpret_cast<const double2*>(normals + (b * n_samp + 2 * p) * 2);
        const double2 t0 = np_[0], t1 = np_[1];
        n0 = t0.x; n1 = t0.y; n2 = t1.x; n3 = t1.y;
      } else {
        uint32_t r[4];
        Philox{seed}((uint64_t)(b * npair + p), 0xB0B0u, r);
        float f0, f1, f2, f3;
        box_muller(r[0], r[1], f0, f1);
        box_muller(r[2], r[3], f2, f3);
        n0 = f0; n1 = f1; n2 = f2; n3 = f3;
      }
      xo[p] = make_float4((float)((double)v.x * inv + n0 * std_b), (float)((double)v.y * inv + n1 * std_b),
                          (float)((double)v.z * inv + n2 * std_b), (float)((double)v.w * inv + n3 * std_b));
    }
    if ((n_samp & 1) && lane == 0) {                   // odd tail (not reached by the LTE geometries)
      const int i = n_samp - 1;
      const float2 v = xin[b * n_samp + i];
      double n0, n1;
      if (normals) {
        n0 = normals[(b * n_samp + i) * 2];
        n1 = normals[(b * n_samp + i) * 2 + 1];
      } else {
        uint32_t r[4];
        Philox{seed}((uint64_t)(B * npair + b), 0xB0B1u, r);
        float f0, f1;
        box_muller(r[0], r[1], f0, f1);
        n0 = f0; n1 = f1;
      }
      out[b * n_samp + i] = make_float2((float)((double)v.x * inv + n0 * std_b), (float)((double)v.y * inv + n1 * std_b));
    }
  }
}

// =====================================================================================
// OFDM transmitter (dev/py/ofdm.py:328-380): one warp per OFDM symbol.
//   grid[k] = constellation[bits] on data carriers, pilot value on pilots, 0 elsewhere
//   time    = ifft(grid)  (naive 64-point DFT, fp64 accumulate, fp32 out), CP prepended.
// sc_map [S*K]: -1 guard, -2 pilot, >=0 data index inside the frame.
// =====================================================================================
__global__ void __launch_bounds__(256) tx_kernel(const uint8_t* __restrict__ bits, long long B, int S, int K, int CP,
                                                 int nbits, int D, const int* __restrict__ sc_map,
                                                 const float2* __restrict__ constellation, float2 pilot,
                                                 float2* __restrict__ tx) {
  extern __shared__ double2 tx_sm[];     // [8 warps][K] grid + [K] twiddles
  double2* tw = tx_sm + 8 * K;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    double s, c;
    sincospi(2.0 * i / K, &s, &c);
    tw[i] = make_double2(c, s);
  }
  __syncthreads();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long sym = (long long)blockIdx.x * 8 + wib;
  if (sym >= B * S) return;
  const long long frame = sym / S;
  const int s = (int)(sym % S);
  double2* g = tx_sm + wib * K;
  for (int k = lane; k < K; k += 32) {
    const int m = sc_map[s * K + k];
    float2 v = make_float2(0.f, 0.f);
    if (m == -2) v = pilot;
    else if (m >= 0) {
      int idx = 0;
      const uint8_t* bp = bits + ((size_t)frame * D + m) * nbits;
      for (int b = 0; b < nbits; ++b) idx = (idx << 1) | bp[b];
      v = constellation[idx];
    }
    g[k] = make_double2(v.x, v.y);
  }
  __syncwarp();
  const int T = K + CP;
  float2* o = tx + (size_t)sym * T;
  for (int n = lane; n < K; n += 32) {
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < K; ++k) {
      const double2 w = tw[(n * k) % K];
      const double2 x = g[k];
      ar += x.x * w.x - x.y * w.y;
      ai += x.x * w.y + x.y * w.x;
    }
    const float2 r = make_float2((float)(ar / K), (float)(ai / K));
    o[CP + n] = r;
    if (n >= K - CP) o[n - (K - CP)] = r;
  }
}

// subcarrier role map of a frame: -1 guard / unused, -2 pilot, >= 0 index of the data symbol carried (one block)
__global__ void txmap_kernel(const int* __restrict__ data_sc, int n_data, const int* __restrict__ pilot_sc, int n_pilot,
                             int n, int* __restrict__ map) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) map[i] = -1;
  __syncthreads();
  for (int i = threadIdx.x; i < n_data; i += blockDim.x) {
    const int k = data_sc[i];
    if (k >= 0 && k < n) map[k] = i;        // out-of-range indices are ignored (the host validates ofdmobj.dataSc)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pilot; i += blockDim.x) {
    const int k = pilot_sc[i];
    if (k >= 0 && k < n) map[k] = -2;
  }
}

// ---- shared by tx64_kernel and tx_fade_kernel: FOUR OFDM symbols of K = 64 subcarriers per warp ------------------------
// 64-point inverse DFT as 8 x 8 (n = 8 n1 + n2, k = k1 + 8 k2):
//   Y[k1][n2] = sum_k2 X[k1 + 8 k2] W8^(n2 k2),   Z = Y * W64^(n2 k1),   x[8 n1 + n2] = sum_k1 Z[k1][n2] W8^(n1 k1) / 64
// with each 8-point transform done by ONE lane as radix-2 butterflies in registers (48 complex additions and two
// multiplications by (+-1 + j)/sqrt(2) instead of the 64 complex MACs of the matrix form -- the transmitter is fp64-bound),
// lane = 8 * slot + q: slot = which of the warp's four symbols, q = k1 in the first pass and n2 in the second.  The
// frequency grid never touches shared memory (a lane maps the 8 subcarriers q + 8 k2 itself); Z goes through a padded
// [8][9] tile per slot (conflict-free both ways).  fp64 like NumPy's pocketfft; results are rounded to fp32 by the caller.
constexpr int kTxZ = 72;                       // double2 per slot: Z[k1 * 9 + n2]

// inverse 8-point DFT in place: a[n] <- sum_k a[k] exp(+2 pi j n k / 8)
DCCN_DEVINL void idft8(double2 (&a)[8]) {
  const double r = 0.70710678118654752440;
  const double2 b0 = make_double2(a[0].x + a[4].x, a[0].y + a[4].y), b1 = make_double2(a[0].x - a[4].x, a[0].y - a[4].y);
  const double2 b2 = make_double2(a[2].x + a[6].x, a[2].y + a[6].y), b3 = make_double2(a[2].x - a[6].x, a[2].y - a[6].y);
  const double2 b4 = make_double2(a[1].x + a[5].x, a[1].y + a[5].y), b5 = make_double2(a[1].x - a[5].x, a[1].y - a[5].y);
  const double2 b6 = make_double2(a[3].x + a[7].x, a[3].y + a[7].y), b7 = make_double2(a[3].x - a[7].x, a[3].y - a[7].y);
  // j * (x + j y) = -y + j x
  const double2 c0 = make_double2(b0.x + b2.x, b0.y + b2.y), c2 = make_double2(b0.x - b2.x, b0.y - b2.y);
  const double2 c1 = make_double2(b1.x - b3.y, b1.y + b3.x), c3 = make_double2(b1.x + b3.y, b1.y - b3.x);
  const double2 c4 = make_double2(b4.x + b6.x, b4.y + b6.y), c6 = make_double2(b4.x - b6.x, b4.y - b6.y);
  const double2 c5 = make_double2(b5.x - b7.y, b5.y + b7.x), c7 = make_double2(b5.x + b7.y, b5.y - b7.x);
  // W8 c5 = (1 + j)/sqrt2 (x + j y) = ((x - y) + j (x + y)) r ;  W8^3 c7 = (-1 + j)/sqrt2 (x + j y) = ((-x - y) + j (x - y)) r
  const double2 w5 = make_double2((c5.x - c5.y) * r, (c5.x + c5.y) * r);
  const double2 w7 = make_double2((-c7.x - c7.y) * r, (c7.x - c7.y) * r);
  a[0] = make_double2(c0.x + c4.x, c0.y + c4.y);
  a[4] = make_double2(c0.x - c4.x, c0.y - c4.y);
  a[1] = make_double2(c1.x + w5.x, c1.y + w5.y);
  a[5] = make_double2(c1.x - w5.x, c1.y - w5.y);
  a[2] = make_double2(c2.x - c6.y, c2.y + c6.x);
  a[6] = make_double2(c2.x + c6.y, c2.y - c6.x);
  a[3] = make_double2(c3.x + w7.x, c3.y + w7.y);
  a[7] = make_double2(c3.x - w7.x, c3.y - w7.y);
}

// W64^i = exp(+2 pi j i / 64), i = 0..63, once per block
DCCN_DEVINL void tx64_twiddle_table(double2* tw64) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    double sn, cs;
    sincospi((double)i / 32.0, &sn, &cs);
    tw64[i] = make_double2(cs, sn);
  }
}

// Symbol `s` of the frame whose label bytes start at `fbits` (this lane's slot; `active` = the slot holds a symbol) ->
// x[n1] = time sample 8 n1 + q of the symbol, scaled by 1/64, no cyclic prefix.  Whole warp must call (two __syncwarp
// inside).  The eight role / label / constellation lookups of a lane are issued as three batches of independent
// (predicated) loads -- written with branches they formed eight serial map -> label -> constellation chains, which was
// where the fused feeder spent a third of its stall samples.
DCCN_DEVINL void tx64_group(const uint8_t* fbits, int s, bool active, int nbits,
                            const int* __restrict__ sc_map, const float2* __restrict__ constellation, float2 pilot,
                            int lane, const double2* tw64, double2* z_slot, double2 (&x)[8]) {
  constexpr int K = 64;
  const int q = lane & 7;
  int m[8];
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) m[k2] = active ? sc_map[s * K + q + 8 * k2] : -1;
  int idx[8];
  if (nbits == 4) {            // the symbol's label bytes in one aligned word (MSB-first index, ofdm.py:121-153)
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      const uint32_t w = m[k2] >= 0 ? *reinterpret_cast<const uint32_t*>(fbits + 4 * m[k2]) : 0u;
      idx[k2] = (int)(((w & 1u) << 3) | ((w >> 6) & 4u) | ((w >> 15) & 2u) | ((w >> 24) & 1u));
    }
  } else if (nbits == 2) {
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      const uint32_t w = m[k2] >= 0 ? *reinterpret_cast<const uint16_t*>(fbits + 2 * m[k2]) : 0u;
      idx[k2] = (int)(((w & 1u) << 1) | ((w >> 8) & 1u));
    }
  } else {
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      int v = 0;
      for (int b = 0; b < nbits; ++b) v = (v << 1) | (m[k2] >= 0 ? fbits[m[k2] * nbits + b] : 0);
      idx[k2] = v;
    }
  }
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const float2 c = constellation[idx[k2]];
    const float2 v = m[k2] >= 0 ? c : (m[k2] == -2 ? pilot : make_float2(0.f, 0.f));
    x[k2] = make_double2(v.x, v.y);
  }
  idft8(x);                                            // x[n2] = Y[k1 = q][n2]
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) {
    const double2 w = tw64[(n2 * q) & 63];
    z_slot[q * 9 + n2] = make_double2(fma(-x[n2].y, w.y, x[n2].x * w.x), fma(x[n2].y, w.x, x[n2].x * w.y));
  }
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) x[k1] = z_slot[k1 * 9 + q];      // Z[k1][n2 = q]
  __syncwarp();
  idft8(x);                                            // x[n1] = 64 * sample 8 n1 + q
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) x[n1] = make_double2(x[n1].x * (1.0 / K), x[n1].y * (1.0 / K));
}

// The same transmitter for K = 64 (the default since round 2; DCCN_TX_V2=0 selects tx_kernel).
__global__ void __launch_bounds__(256) tx64_kernel(const uint8_t* __restrict__ bits, long long B, int S, int CP, int nbits,
                                                   int D, const int* __restrict__ sc_map,
                                                   const float2* __restrict__ constellation, float2 pilot,
                                                   float2* __restrict__ tx) {
  constexpr int K = 64;
  __shared__ double2 sm_tw[64];
  __shared__ double2 sm_z[8][4][kTxZ];
  tx64_twiddle_table(sm_tw);
  __syncthreads();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane >> 3, q = lane & 7;
  const int T = K + CP;
  const long long total = B * S;
  for (long long sym0 = ((long long)blockIdx.x * 8 + wib) * 4; sym0 < total; sym0 += (long long)gridDim.x * 32) {
    const long long sym = sym0 + slot;
    const bool active = sym < total;
    const long long frame = active ? sym / S : 0;
    const int s = active ? (int)(sym - frame * S) : 0;
    double2 x[8];
    tx64_group(bits + (size_t)frame * D * nbits, s, active, nbits, sc_map, constellation, pilot, lane, sm_tw, sm_z[wib][slot], x);
    if (active) {
      float2* o = tx + (size_t)sym * T;
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        const int n = 8 * n1 + q;
        const float2 v = make_float2((float)x[n1].x, (float)x[n1].y);
        o[CP + n] = v;
        if (n >= K - CP) o[n - (K - CP)] = v;           // cyclic prefix = the last CP samples
      }
    }
  }
}

// =====================================================================================
// Fused feeder: bits -> constellation -> 8 x 8 IDFT -> CP (tx64_kernel's arithmetic) -> static Rayleigh FIR
// (chan_fir_kernel's arithmetic) for nfft = 64, one warp per frame.  The transmitted frame lives in shared memory as the
// fp32 samples tx64_kernel would have written (the reference's complex64 `iq_tx_cmpx`, dev/py/ofdm.py:380), so the
// result is bit-identical to chan_fir_kernel(tx64_kernel(bits)) while the 4 480-byte frame is neither written to nor
// read back from HBM (a sweep cell generates its frames once and never looks at the unfaded signal):
// replaces the call pair ofdm_tx_frame_np + rayleigh_chan_lte.run of dev/py/ofdmreceiver_np.py:227-228.
//   tx_out: optional [B, S*T] copy of the transmitted frames (nullptr = not wanted)
// =====================================================================================
constexpr int kGenWarps = 4;
constexpr int kGenMaxSamp = 8 * 80;     // S <= 8 symbols of K + CP <= 80 samples
constexpr int kGenBitBytes = 1536;      // label bytes of a frame staged per warp (368 data cells x 4 bits max)

// Shared-memory frame buffers of the fused feeder use a padded layout: sample n of a frame sits at n + n / L (L = samples
// per lane in the FIR's blocked mapping), i.e. lane l's run starts at (L + 1) l -- an odd stride in 8-byte words, so the
// 32 lanes of a warp hit 32 different bank pairs when each reads "its" i-th sample.
template <int L>
DCCN_DEVINL int fr_pos(int n) { return n + n / L; }

// FIR of one frame, blocked mapping: lane l produces outputs [L l, L l + L) from a register window that slides over
// fr[] (each input sample is loaded from shared memory and converted to double ONCE per lane instead of once per tap),
// taps in registers (MT >= M of them, zero-padded).  out[n] = sum_j g[j] x[n + off - j], zero outside the frame
// (np.convolve(tx, g, 'same'), radio.py:436), complex128, same accumulation order as chan_fir_kernel (fir_cmac, j ascending).
template <int L, int MT>
DCCN_DEVINL double fir_blocked(const float2* fr, float2* fo, const double2* gs, int M, int off, int n_samp, int lane) {
  double2 g[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j) g[j] = j < M ? gs[j] : make_double2(0.0, 0.0);
  // window w[t] = x[n0 + t - (MT - 1)] for the current output n0 + i; element for tap j of output i: x[n0 + i + off - j]
  const int n0 = L * lane;
  double2 win[MT];                                   // win[j] = x[n + off - j] for the current n
  auto load = [&](int n) -> double2 {
    if (n < 0 || n >= n_samp) return make_double2(0.0, 0.0);
    const float2 v = fr[fr_pos<L>(n)];
    return make_double2((double)v.x, (double)v.y);
  };
#pragma unroll
  for (int j = 1; j < MT; ++j) win[j] = load(n0 + off - j);      // history of the first output (j = 0 is loaded in the loop)
  double pw = 0.0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const int n = n0 + i;
    win[0] = load(n + off);
    double accr = 0.0, acci = 0.0;
#pragma unroll
    for (int j = 0; j < MT; ++j) fir_cmac(accr, acci, g[j], win[j].x, win[j].y);
    if (n < n_samp) {
      const float2 o = make_float2((float)accr, (float)acci);
      fo[fr_pos<L>(n)] = o;
      pw += (double)o.x * o.x + (double)o.y * o.y;
    }
#pragma unroll
    for (int j = MT - 1; j > 0; --j) win[j] = win[j - 1];       // renaming only: the loops are fully unrolled
  }
  return pw;
}

template <int L>
__global__ void __launch_bounds__(32 * kGenWarps) tx_fade_kernel(
    const uint8_t* __restrict__ bits, long long B, int S, int CP, int nbits, int D, const int* __restrict__ sc_map,
    const float2* __restrict__ constellation, float2 pilot, const double* __restrict__ alpha,
    const double* __restrict__ coeff, int n_taps, int n_fir, const double* __restrict__ z_in, uint64_t seed,
    float2* __restrict__ tx_out, float2* __restrict__ rx, double* __restrict__ power_sum) {
  constexpr int K = 64;
  constexpr int FR = 32 * (L + 1);                   // padded frame buffer (float2)
  extern __shared__ __align__(16) uint8_t gen_smem[];
  double2* sm_tw = reinterpret_cast<double2*>(gen_smem);                                  // [64]
  double2* sm_z_all = sm_tw + 64;                                                          // [warps][4][kTxZ]
  double2* gsm_all = sm_z_all + kGenWarps * 4 * kTxZ;                                      // [warps][kMaxFir]
  float2* sm_fr_all = reinterpret_cast<float2*>(gsm_all + kGenWarps * kMaxFir);            // [warps][FR]
  float2* sm_fo_all = sm_fr_all + kGenWarps * FR;                                          // [warps][FR]
  float2* sm_fo_end = sm_fo_all + kGenWarps * FR;
  int* sm_map = reinterpret_cast<int*>(sm_fo_end);                                         // [S * K] subcarrier roles
  float2* sm_const = reinterpret_cast<float2*>(sm_map + 8 * K);                            // [16] constellation
  uint8_t* sm_bits_all = reinterpret_cast<uint8_t*>(sm_const + 16);                        // [warps][kGenBitBytes] labels of the frame
  tx64_twiddle_table(sm_tw);
  for (int i = threadIdx.x; i < S * K; i += blockDim.x) sm_map[i] = sc_map[i];
  for (int i = threadIdx.x; i < (1 << nbits); i += blockDim.x) sm_const[i] = constellation[i];
  __syncthreads();
  sc_map = sm_map;
  constellation = sm_const;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = lane >> 3, q = lane & 7;
  double2* gs = gsm_all + wib * kMaxFir;
  float2* fr = sm_fr_all + wib * FR;
  float2* fo = sm_fo_all + wib * FR;
  uint8_t* fb = sm_bits_all + wib * kGenBitBytes;
  const int nbytes = D * nbits;
  const bool stage_bits = nbytes <= kGenBitBytes && (nbytes & 15) == 0;
  const int T = K + CP, n_samp = S * T;
  const int M = n_taps == 0 ? 1 : n_fir;
  const int off = (M - 1) - (M >> 1);
  double pw = 0.0;
  for (long long frame = (long long)blockIdx.x * kGenWarps + wib; frame < B; frame += (long long)gridDim.x * kGenWarps) {
    // the frame's labels: one coalesced copy into shared memory (the per-subcarrier lookups then never wait on HBM)
    const uint8_t* fbits = bits + (size_t)frame * nbytes;
    if (stage_bits) {
      for (int i = lane; i < (nbytes >> 4); i += 32)
        reinterpret_cast<uint4*>(fb)[i] = __ldg(reinterpret_cast<const uint4*>(fbits) + i);
      fbits = fb;
    }
    // ---- path gains -> sample-spaced FIR (chan_fir_kernel) ----
    if (n_taps == 0) {
      if (lane == 0) gs[0] = make_double2(1.0, 0.0);
    } else {
      double2 pa = make_double2(0.0, 0.0);
      if (lane < n_taps) {
        double zr, zi;
        if (z_in) {
          zr = z_in[((size_t)frame * n_taps + lane) * 2];
          zi = z_in[((size_t)frame * n_taps + lane) * 2 + 1];
        } else {
          uint32_t r[4];
          Philox{seed}((uint64_t)frame * 64 + lane, 0xA11CEu, r);
          float n0, n1;
          box_muller(r[0], r[1], n0, n1);
          zr = (double)n0 * 0.70710678118654752;
          zi = (double)n1 * 0.70710678118654752;
        }
        const double c = coeff[lane];
        pa = make_double2(zr * c, zi * c);
      }
      double2 gt = make_double2(0.0, 0.0);
      for (int t = 0; t < n_taps; ++t) {
        const double ar = __shfl_sync(0xffffffffu, pa.x, t);
        const double ai = __shfl_sync(0xffffffffu, pa.y, t);
        if (lane < n_fir) {
          const double al = alpha ? alpha[t * n_fir + lane] : 1.0;
          gt.x = fma(ar, al, gt.x);
          gt.y = fma(ai, al, gt.y);
        }
      }
      if (lane < n_fir) gs[lane] = gt;
    }
    __syncwarp();                                      // staged labels and taps visible to every lane
    // ---- transmitter: the S symbols, four at a time, into the shared frame buffer as fp32 (tx64_kernel's values) ----
    for (int s0 = 0; s0 < S; s0 += 4) {
      const int s = s0 + slot;
      const bool active = s < S;
      double2 x[8];
      tx64_group(fbits, active ? s : 0, active, nbits, sc_map, constellation, pilot, lane, sm_tw,
                 sm_z_all + (wib * 4 + slot) * kTxZ, x);
      if (active) {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int n = 8 * n1 + q;
          const float2 v = make_float2((float)x[n1].x, (float)x[n1].y);
          fr[fr_pos<L>(s * T + CP + n)] = v;
          if (n >= K - CP) fr[fr_pos<L>(s * T + n - (K - CP))] = v;
        }
      }
    }
    __syncwarp();
    if (tx_out) {
      float2* o = tx_out + (size_t)frame * n_samp;
      for (int n = lane; n < n_samp; n += 32) o[n] = fr[fr_pos<L>(n)];
    }
    // ---- centred 'same' FIR with zero history, complex128 ----
    if (M <= 9) pw += fir_blocked<L, 9>(fr, fo, gs, M, off, n_samp, lane);
    else if (M <= 13) pw += fir_blocked<L, 13>(fr, fo, gs, M, off, n_samp, lane);
    else {                                             // long filters: interleaved mapping, taps and samples from shared memory
      for (int n = lane; n < n_samp; n += 32) {
        double accr = 0.0, acci = 0.0;
        for (int j = 0; j < M; ++j) {
          const int i = n + off - j;
          const float2 xv = (i >= 0 && i < n_samp) ? fr[fr_pos<L>(i)] : make_float2(0.f, 0.f);
          fir_cmac(accr, acci, gs[j], (double)xv.x, (double)xv.y);
        }
        const float2 o = make_float2((float)accr, (float)acci);
        fo[fr_pos<L>(n)] = o;
        pw += (double)o.x * o.x + (double)o.y * o.y;
      }
    }
    __syncwarp();
    float2* rxf = rx + (size_t)frame * n_samp;
    for (int n = lane; n < n_samp; n += 32) rxf[n] = fo[fr_pos<L>(n)];     // coalesced copy-out
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
  if (lane == 0 && pw != 0.0) atomicAdd(power_sum, pw);
}

template <int L>
constexpr size_t tx_fade_smem() {
  return (64 + kGenWarps * 4 * kTxZ + kGenWarps * kMaxFir) * sizeof(double2) + 2 * kGenWarps * 32 * (L + 1) * sizeof(float2) +
         8 * 64 * sizeof(int) + 16 * sizeof(float2) + kGenWarps * kGenBitBytes;
}

// util.bit_source (dev/py/util.py:25-34): n uniform bits, one per byte.  A thread expands ONE Philox-4x32 call (128 random
// bits) into 128 bytes with eight 16-byte stores (round 1 drew one call per 16 bytes and stored them a byte at a time:
// 0.22 ms per 65 536-frame cell, now store-bound).
__global__ void __launch_bounds__(256) bit_source_kernel(uint8_t* __restrict__ bits, long long n, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 128 bytes per thread
  if (i * 128 >= n) return;
  uint32_t r[4];
  Philox{seed}((uint64_t)i, 0xB175u, r);
  auto spread = [](uint32_t v) {      // 4 bits -> 4 bytes of 0 / 1
    return (v & 1u) | ((v & 2u) << 7) | ((v & 4u) << 14) | ((v & 8u) << 21);
  };
  if (i * 128 + 128 <= n && (reinterpret_cast<uintptr_t>(bits) & 15) == 0) {
    uint4* o = reinterpret_cast<uint4*>(bits + i * 128);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t v = r[w] >> (16 * h);
        o[2 * w + h] = make_uint4(spread(v), spread(v >> 4), spread(v >> 8), spread(v >> 12));
      }
    }
  } else {
    for (int j = 0; j < 128 && i * 128 + j < n; ++j) bits[i * 128 + j] = (uint8_t)((r[j >> 5] >> (j & 31)) & 1u);
  }
}

// =====================================================================================
// Demodulation head as a stand-alone, fully parallel kernel: one thread per (frame, data
// subcarrier).  Same arithmetic as EpiHead (it calls it): the GEMM epilogue variant is bound
// by the latency of 8 epilogue warps, this one runs at full occupancy.
// out_iq [M, 2D] fp32 (bias already added by the GEMM epilogue).
// =====================================================================================
template <int NB, bool V1>
DCCN_DEVINL void head_emit(const EpiHead<NB, V1>& epi, typename EpiHead<NB, V1>::State& st, size_t o0, unsigned hbits,
                           const float (&p)[2 * NB]) {
  if (epi.soft) {
    float* sp = epi.soft + o0 * 2;
    if constexpr ((2 * NB) % 4 == 0) {
#pragma unroll
      for (int q = 0; q < 2 * NB; q += 4) *reinterpret_cast<float4*>(sp + q) = make_float4(p[q], p[q + 1], p[q + 2], p[q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 2 * NB; q += 2) *reinterpret_cast<float2*>(sp + q) = make_float2(p[q], p[q + 1]);
    }
  }
  // decision bit k -> byte k (0 / 1) and label byte k -> bit k, each with one multiply: bit k times (1 + 2^7 + 2^14 + 2^21) lands
  // on bit 8k of byte k (every other copy falls between the byte LSBs); byte j's LSB times 2^(24 - 7j) lands on bit 24 + j
  unsigned y = 0u;
  const unsigned hw = (hbits * 0x00204081u) & 0x01010101u;
  if constexpr (NB == 4) {
    if (epi.hard) *reinterpret_cast<uint32_t*>(epi.hard + o0) = hw;
    if (epi.bits) y = ((__ldg(reinterpret_cast<const uint32_t*>(epi.bits + o0)) & 0x01010101u) * 0x01020408u) >> 24;
  } else if constexpr (NB == 2) {
    if (epi.hard) *reinterpret_cast<uint16_t*>(epi.hard + o0) = (uint16_t)hw;
    if (epi.bits) y = (((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(epi.bits + o0)) & 0x0101u) * 0x01020408u) >> 24;
  } else {
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (epi.hard) epi.hard[o0 + k] = (uint8_t)((hbits >> k) & 1u);
      if (epi.bits) y |= (uint32_t)(__ldg(epi.bits + o0 + k) & 1u) << k;
    }
  }
  if (epi.bits) epi.account(st, y, hbits, p);
}

template <int NB, bool V1, int SUBS>
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ out_iq, const __grid_constant__ EpiHead<NB, V1> epi) {
  // block <-> frame (grid-stride), thread <-> SUBS data subcarriers (d, d + blockDim.x): no index division anywhere; a warp
  // reads 32 adjacent (I, Q) pairs and writes 32 adjacent 8*NB-byte probability records (full 32-byte sectors) and NB-byte
  // decision records.  SUBS = 2: the two subcarriers share every weight fetch (the 200 head weights live in the constant
  // bank and go through the uniform datapath) and give the scheduler two independent dependency chains.  The
  // per-subcarrier arithmetic is EpiHead's (shared with the fused GEMM-epilogue form).
  const int D = epi.N >> 1;
  typename EpiHead<NB, V1>::State st;
  for (int row = blockIdx.x; row < epi.M; row += gridDim.x) {
    const float2* iqp = reinterpret_cast<const float2*>(out_iq + (size_t)row * epi.N);
    if constexpr (SUBS == 2) {
      for (int d = threadIdx.x; d < D; d += 2 * blockDim.x) {
        const int d2 = d + blockDim.x;
        const bool two = d2 < D;
        const float2 iq0 = __ldg(iqp + d);
        const float2 iq1 = two ? __ldg(iqp + d2) : make_float2(0.f, 0.f);
        float p0[2 * NB], p1[2 * NB];
        const unsigned h0 = epi.subcarrier(iq0.x, iq0.y, p0);
        const unsigned h1 = epi.subcarrier(iq1.x, iq1.y, p1);
        head_emit<NB, V1>(epi, st, ((size_t)row * D + d) * NB, h0, p0);
        if (two) head_emit<NB, V1>(epi, st, ((size_t)row * D + d2) * NB, h1, p1);
      }
    } else {
      for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float2 iq0 = __ldg(iqp + d);
        float p0[2 * NB];
        const unsigned h0 = epi.subcarrier(iq0.x, iq0.y, p0);
        head_emit<NB, V1>(epi, st, ((size_t)row * D + d) * NB, h0, p0);
      }
    }
  }
  epi.flush(st);
}

// labels packed 8 per byte (bit j of byte i = label 8 i + j, i.e. numpy.packbits(bitorder='little')) -> one uint8 per
// label, the layout the head reads (dccn_forward_host_begin_packed: 160 instead of 1 280 label bytes per 16-QAM frame
// over PCIe)
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint8_t* __restrict__ packed, long long n_bytes,
                                                          uint8_t* __restrict__ bits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t b = packed[i];
    uint2 o;
    o.x = (b & 1u) | (((b >> 1) & 1u) << 8) | (((b >> 2) & 1u) << 16) | (((b >> 3) & 1u) << 24);
    o.y = ((b >> 4) & 1u) | (((b >> 5) & 1u) << 8) | (((b >> 6) & 1u) << 16) | (((b >> 7) & 1u) << 24);
    *reinterpret_cast<uint2*>(bits + 8 * i) = o;
  }
}

// =====================================================================================
// Monitor tensors of the reference graph (dev/py/ofdmreceiver_np.py:125-149,172-183) -- NOT on the receiver's data path:
//   input    = batch-moment norm of tx_ofdm / sqrt(2)                       ('input:0', what the receiver consumes)
//   iq_layer = tf.clip_by_norm(input, 8, axes=[-1])   (complex.py:21-27)   -> tx_power = mean(I^2 + Q^2), iq_tx (fp16)
//   the in-graph AWGN (radio.py:62-88, built but bypassed, :136-138): noise = |level * N(0,1)| * (sin phi, cos phi) with
//   level = sqrt(.5) 10^(-SNR/20), phi ~ U(0, 2 pi)                         -> noise_power = mean(|noise|^2), iq_rx (fp16)
// The second batch normalisation inside AWGN_channel (eps 1e-8, / sqrt 2) maps the already normalised batch onto itself to
// 1e-8 relative, so iq_rx = iq_layer + noise.  One warp per frame; sums[0] = sum |iq_layer|^2, sums[1] = sum |noise|^2.
// TF's Philox stream cannot be reproduced, so noise_power / iq_rx agree with the reference in distribution only.
// =====================================================================================
__global__ void __launch_bounds__(256) monitor_kernel(const float2* __restrict__ x, long long B, int n_samp,
                                                      const float2* __restrict__ mean, const float2* __restrict__ rstd,
                                                      const float* __restrict__ snr_db, uint64_t seed,
                                                      double* __restrict__ sums, float2* __restrict__ input_out,
                                                      __half2* __restrict__ iq_tx, __half2* __restrict__ iq_rx) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double p_tx = 0.0, p_n = 0.0;
  for (long long b = warp0; b < B; b += nwarps) {
    const float level = snr_db ? 0.70710678118f * exp10f(-__ldg(snr_db + b) / 20.0f) : 0.f;
    for (int n = lane; n < n_samp; n += 32) {
      const float2 v = x[b * n_samp + n], m = __ldg(mean + n), r = __ldg(rstd + n);
      float2 z;
      z.x = __fdiv_rn(__fadd_rn(__fmul_rn(v.x, r.x), __fmul_rn(-m.x, r.x)), 1.41421356237f);
      z.y = __fdiv_rn(__fadd_rn(__fmul_rn(v.y, r.y), __fmul_rn(-m.y, r.y)), 1.41421356237f);
      if (input_out) input_out[b * n_samp + n] = z;
      const float l2 = sqrtf(z.x * z.x + z.y * z.y);
      const float den = fmaxf(l2, 8.0f);
      const float2 c = make_float2(z.x * 8.0f / den, z.y * 8.0f / den);
      p_tx += (double)(c.x * c.x + c.y * c.y);
      if (iq_tx) iq_tx[b * n_samp + n] = __floats2half2_rn(c.x, c.y);
      if (snr_db) {
        uint32_t rr[4];
        Philox{seed}((uint64_t)(b * n_samp + n), 0x1107u, rr);
        float n0, n1, s, cs;
        box_muller(rr[0], rr[1], n0, n1);
        const float amp = fabsf(level * n0);
        __sincosf(6.283185307179586f * ((float)rr[2] * 2.3283064365386963e-10f), &s, &cs);
        const float2 nz = make_float2(amp * s, amp * cs);
        p_n += (double)(nz.x * nz.x + nz.y * nz.y);
        if (iq_rx) iq_rx[b * n_samp + n] = __floats2half2_rn(c.x + nz.x, c.y + nz.y);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    p_tx += __shfl_xor_sync(0xffffffffu, p_tx, o);
    p_n += __shfl_xor_sync(0xffffffffu, p_n, o);
  }
  if (lane == 0) {
    atomicAdd(sums, p_tx);
    atomicAdd(sums + 1, p_n);
  }
}

// equalizer_ofdm's SNR monitor (dev/py/model.py:464-475): over the S * P pilot-carrier points of the phase-equalised
// frequency-domain frame, signal = mean |eq|^2, noise = variance of |eq|^2, snr_db = log10(clip(signal / noise, 1e-3, 1e4)).
// (The reference names it snr_db but takes a plain log10 -- kept.)  One warp per frame; eq [B, S, K] complex.
__global__ void __launch_bounds__(256) snr_monitor_kernel(const float2* __restrict__ eq, long long B, int S, int K,
                                                          const int* __restrict__ pilot_carriers, int P,
                                                          float* __restrict__ snr_out) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int n = S * P;
  float s1 = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float2 v = eq[(b * S + i / P) * K + pilot_carriers[i % P]];
    s1 += v.x * v.x + v.y * v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  const float mu = s1 / (float)n;
  float s2 = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float2 v = eq[(b * S + i / P) * K + pilot_carriers[i % P]];
    const float d = v.x * v.x + v.y * v.y - mu;
    s2 += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if (lane == 0) snr_out[b] = log10f(fminf(fmaxf(mu / (s2 / (float)n), 0.001f), 10000.0f));
}

// a5: confusion matrix of hard decisions vs bits (rows = truth)
__global__ void __launch_bounds__(256) ber_accum_kernel(const uint8_t* __restrict__ hard,
                                                        const uint8_t* __restrict__ bits, long long n,
                                                        unsigned long long* __restrict__ conf) {
  unsigned c[4] = {0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c[((bits[i] & 1) << 1) | (hard[i] & 1)]++;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned t = __reduce_add_sync(0xffffffffu, c[k]);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(conf + k, (unsigned long long)t);
  }
}

// =====================================================================================
// a1 op-level: direct evaluation of layers_conv2d_complex (complex.py:140-196).
// One thread per complex output element (b, l, w, f).
// =====================================================================================
__global__ void __launch_bounds__(128) cconv2d_kernel(const float2* __restrict__ x, long long B, int L, int W, int C,
                                                      const float* __restrict__ kernel, const float* __restrict__ bias,
                                                      int F, int kl, int kw, int pl, int pw, int Lo, int Wo,
                                                      float2* __restrict__ y) {
  const long long total = B * Lo * Wo * F;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % F);
  long long r = i / F;
  const int w = (int)(r % Wo);
  r /= Wo;
  const int l = (int)(r % Lo);
  const long long b = r / Lo;
  // c0 = xr*Wa, c1 = xr*Wb, c2 = xi*Wa, c3 = xi*Wb  (each + its conv3d bias)
  float c0 = bias[f], c1 = bias[F + f], c2 = bias[f], c3 = bias[F + f];
  for (int ii = 0; ii < kl; ++ii) {
    const int li = l + ii - pl;
    if (li < 0 || li >= L) continue;
    for (int jj = 0; jj < kw; ++jj) {
      const int wj = w + jj - pw;
      if (wj < 0 || wj >= W) continue;
      const float2* xp = x + (((size_t)b * L + li) * W + wj) * C;
      const float* kp = kernel + ((size_t)(ii * kw + jj) * C) * (2 * F);
      for (int c = 0; c < C; ++c) {
        const float2 xv = __ldg(xp + c);
        const float wa = __ldg(kp + (size_t)c * 2 * F + f), wb = __ldg(kp + (size_t)c * 2 * F + F + f);
        c0 = fmaf(xv.x, wa, c0);
        c1 = fmaf(xv.x, wb, c1);
        c2 = fmaf(xv.y, wa, c2);
        c3 = fmaf(xv.y, wb, c3);
      }
    }
  }
  y[i] = make_float2(c0 - c3, c1 - c2);   // complex.py:187-188
}

// =====================================================================================
// op-level: direct evaluation of layers_conv2d_vector (complex.py:199-255): one real conv3d over (length, width, IQ)
// with kernel depth 2 across IQ, 2F channels, no complex recombination.  kernel [kl,kw,2,C,2F], bias [2F].
// After the reference's reshape / slice both paddings reduce to: IQ position 0, channels [0,F) = real parts,
// [F,2F) = imaginary parts ('same' pads the size-2 IQ axis 0 before / 1 after, so position 0 sees both components).
// One thread per output element (b, l, w, f).
// =====================================================================================
__global__ void __launch_bounds__(128) vconv2d_kernel(const float2* __restrict__ x, long long B, int L, int W, int C,
                                                      const float* __restrict__ kernel, const float* __restrict__ bias,
                                                      int F, int kl, int kw, int pl, int pw, int Lo, int Wo,
                                                      float2* __restrict__ y) {
  const long long total = B * Lo * Wo * F;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % F);
  long long r = i / F;
  const int w = (int)(r % Wo);
  r /= Wo;
  const int l = (int)(r % Lo);
  const long long b = r / Lo;
  float re = bias[f], im = bias[F + f];
  for (int ii = 0; ii < kl; ++ii) {
    const int li = l + ii - pl;
    if (li < 0 || li >= L) continue;
    for (int jj = 0; jj < kw; ++jj) {
      const int wj = w + jj - pw;
      if (wj < 0 || wj >= W) continue;
      const float2* xp = x + (((size_t)b * L + li) * W + wj) * C;
      const float* k0 = kernel + ((size_t)(ii * kw + jj) * 2) * C * (2 * F);   // IQ tap 0
      const float* k1 = k0 + (size_t)C * 2 * F;                                // IQ tap 1
      for (int c = 0; c < C; ++c) {
        const float2 xv = __ldg(xp + c);
        re = fmaf(xv.x, __ldg(k0 + (size_t)c * 2 * F + f), re);
        re = fmaf(xv.y, __ldg(k1 + (size_t)c * 2 * F + f), re);
        im = fmaf(xv.x, __ldg(k0 + (size_t)c * 2 * F + F + f), im);
        im = fmaf(xv.y, __ldg(k1 + (size_t)c * 2 * F + F + f), im);
      }
    }
  }
  y[i] = make_float2(re, im);
}

}  // namespace dccn
