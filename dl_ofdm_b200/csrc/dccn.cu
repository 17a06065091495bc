// dccn.cu -- libdccn.so: handle, weight packing, layer schedule and the C ABI (include/dccn.h).
//
// Reference path being replaced (zhongyuanzhao/dl_ofdm @ 5665b50):
//   tx_ofdm -> batch-moment norm (dev/py/ofdmreceiver_np.py:128-129)
//           -> [equalizer_ofdm  dev/py/model.py:349-478]
//           -> ofdm_dense_rx     dev/py/model.py:1222-1292
//           -> argmax / confusion matrix / xent (dev/py/ofdmreceiver_np.py:154-169)
// Every complex layer is `layers_conv2d_complex` (dev/py/complex.py:140-196) and is packed
// into ONE real GEMM on IQ-interleaved activations with the reference's own sign
// convention ([[a, b], [-b, -a]] per complex weight, bias (ba-bb, bb-ba)).
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <type_traits>

#include "handle.cuh"
#include "chain.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace dccn {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ---------------------------------------------------------------------------------------
// TMA descriptor creation (driver entry point fetched through the runtime: no -lcuda)
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// fp32 matrix [rows, cols] with row pitch ld (elements); box = [box_rows x 32 cols], 128B swizzle
int make_tmap(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  DCCN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
  DCCN_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0,
             "TMA operand must be 16-byte aligned (base %p, ld %lld)", (const void*)base, (long long)ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCCN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%d", (int)r,
             (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

// fp16 matrix [rows, cols] with row pitch ld (elements); box = [box_rows x 64 cols] (128-byte rows), 128B swizzle
static int make_tmap_f16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  DCCN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
  DCCN_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
             "TMA operand must be 16-byte aligned (base %p, ld %lld)", base, (long long)ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCCN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed (%d) rows=%lld cols=%lld ld=%lld box=%d", (int)r,
             (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

#ifdef DCCN_TRACE
long long* g_trace_host_ptr = nullptr;
int g_abl_host = 0;
#endif
std::atomic<long long> g_launches{0};
const char* kSlotNames[SLOT_COUNT] = {
    "moments", "prep_norm", "eq_dense", "eq_dft", "eq_pilot", "eq_dense2", "eq_dense3", "eq_dense4_tanh",
    "eq_conv7x64_phaseeq", "eq_corr_idft", "eq_idft", "eq_dense5", "rx_fft_like", "rx_demod_head",
    "chan_fir", "chan_awgn", "rx_demod_gemm", "train_head_bwd", "train_dgrad", "train_wgrad", "train_pointwise",
    "train_reduce_adam", "train_repack", "fold_dense_dft", "fold_mlp_tanh", "fold_tail_fft", "eq_chain_front",
    "eq_chain_tail"};

}  // namespace dccn

using namespace dccn;

namespace dccn {

int dev_alloc(dccn_handle* h, void** p, size_t bytes) {
  DCCN_CUDA_OK(cudaMalloc(p, bytes));
  h->allocs.push_back(*p);
  h->ws_bytes += bytes;
  return 0;
}

// a persistent activation-shaped buffer (freed with the handle)
int alloc_act(dccn_handle* h, Act* a, int64_t rows, int ld, bool split) {
  (void)split;   // activations are one fp32 plane in every mode (hi/lo is made in shared memory)
  a->ld = ld;
  a->p1 = nullptr;
  return dev_alloc(h, (void**)&a->p0, (size_t)rows * ld * sizeof(float));
}

// an inter-layer buffer of the growable workspace (freed and re-made by ensure_workspace)
static int alloc_ws_act(dccn_handle* h, Act* a, int64_t rows, int ld) {
  a->ld = ld;
  a->p1 = nullptr;
  const size_t bytes = (size_t)rows * ld * sizeof(float);
  DCCN_CUDA_OK(cudaMalloc((void**)&a->p0, bytes));
  h->ws_allocs.push_back(a->p0);
  h->ws_bytes += bytes;
  h->ws_act_bytes += bytes;
  return 0;
}

// The inter-layer buffers are sized for the largest pass seen so far (at most h->chunk frames) and grown on demand:
// a handle that only ever sees 256-frame batches holds 10 MB, one that is fed 65 536-frame batches 2.7 GB.
int ensure_workspace(dccn_handle* h, int64_t frames) {
  if (frames > h->chunk) frames = h->chunk;
  frames = (frames + 127) / 128 * 128;
  if (frames <= h->ws_frames && !h->ws_dirty) return 0;
  if (frames < h->ws_frames) frames = h->ws_frames;
  if (h->ws_frames > 0) DCCN_CUDA_OK(cudaDeviceSynchronize());   // nothing may still be using the old buffers
  for (void* p : h->ws_allocs) cudaFree(p);
  for (Act* a : {&h->a0, &h->t1, &h->f, &h->p32, &h->u1, &h->u2, &h->eq, &h->corr, &h->cat, &h->oeq, &h->r1o,
                 &h->out_iq, &h->eqc, &h->u3})
    a->p0 = nullptr;
  h->chest_buf = nullptr;
  h->ws_bytes -= h->ws_act_bytes;
  h->ws_act_bytes = 0;
  h->ws_allocs.clear();
  h->ws_frames = 0;
  h->ws_dirty = false;
  const int64_t C = frames;
  const int S = h->S, K = h->K;
  int rc = 0;
  auto act = [&](Act* a, int64_t rows, int ld) {
    if (!rc) rc = alloc_ws_act(h, a, rows, ld);
  };
  act(&h->a0, C, h->P);
  act(&h->r1o, C, S * h->F * 2);
  act(&h->out_iq, C, 2 * h->D);
  if (h->cfg.equalizer) {
    act(&h->t1, C, S * K * 2);
    act(&h->f, C, S * K * 2);
    act(&h->p32, C, 2 * h->cfg.pilot_size);
    act(&h->u1, C, S * K * 2);
    act(&h->u2, C, S * K * 2);
    act(&h->eq, C, S * K * 2);
    act(&h->corr, C, S * K);
    act(&h->cat, C * S, 4 * K);
    act(&h->oeq, C, h->P);
    if (h->ws_eqc) act(&h->eqc, C * S, 3 * K);
    if (h->ws_train) {
      act(&h->u3, C, S * K * 2);
      if (!rc) {
        Act cb;
        rc = alloc_ws_act(h, &cb, C, S * K * 2);
        h->chest_buf = cb.p0;
      }
    }
  }
  if (rc) return rc;
  {   // one max-|activation| word per inter-layer buffer (zeroed at the start of every pass, run_chunk)
    int i = 0;
    for (Act* a : {&h->a0, &h->t1, &h->f, &h->p32, &h->u1, &h->u2, &h->eq, &h->corr, &h->cat, &h->oeq, &h->r1o,
                   &h->out_iq, &h->eqc, &h->u3})
      a->amax = h->d_amax + (i++);
    static_assert(14 <= kAmaxSlots, "amax slots");
  }
  h->ws_frames = C;
  return 0;
}

const HostTensor* find(const dccn_handle* h, const std::string& n) {
  auto it = h->raw.find(n);
  return it == h->raw.end() ? nullptr : &it->second;
}

// ---------------------------------------------------------------------------------------
// weight packing (host)
// ---------------------------------------------------------------------------------------
// complex layer, per-k rows: Wa[k][f], Wb[k][f] -> Bp [2K, 2F] + bias_p [2F]   
static void pack_complex(const float* Wa, const float* Wb, int64_t strideK, int Kc, int F, const float* bias2F,
                         GemmLayer* L) {
  L->K = 2 * Kc;
  L->N = 2 * F;
  L->W.assign((size_t)L->K * L->N, 0.f);
  for (int k = 0; k < Kc; ++k)
    for (int f = 0; f < F; ++f) {
      const float a = Wa[k * strideK + f], b = Wb[k * strideK + f];
      L->W[(size_t)(2 * k) * L->N + 2 * f] = a;
      L->W[(size_t)(2 * k) * L->N + 2 * f + 1] = b;
      L->W[(size_t)(2 * k + 1) * L->N + 2 * f] = -b;
      L->W[(size_t)(2 * k + 1) * L->N + 2 * f + 1] = -a;
    }
  L->bias.resize(L->N);
  for (int f = 0; f < F; ++f) {
    L->bias[2 * f] = bias2F[f] - bias2F[F + f];
    L->bias[2 * f + 1] = bias2F[F + f] - bias2F[f];
  }
}

static void pack_dense(const HostTensor& k, const HostTensor& b, GemmLayer* L) {
  L->K = (int)k.shape[0];
  L->N = (int)k.shape[1];
  L->W = k.data;
  L->bias = b.data;
}

static int pick_bn(int N, bool fused_epilogue, int wide) {
  if (fused_epilogue) return 128;
  if (N <= 32) return 32;
  if (N > 128 && N <= 192) return 192;
  if (wide && N >= 512) return 256;
  return 128;
}

int upload_layer(dccn_handle* h, GemmLayer* L, cudaStream_t s) {
  const int K = L->K, N = L->N;
  L->BN = pick_bn(N, L->fused, h->bn_wide);
  const int prec = h->cfg.precision;
  // 192-wide tiles only exist as the 2-stage shared-memory-split form (TMEM cannot hold 2x192 accumulator columns plus
  // the A staging); two 128-wide A-in-TMEM tiles (the second mostly empty) are faster.  DCCN_BN192=1 restores them.
  if (L->BN == 192 && prec == DCCN_PREC_PARITY && h->a_tmem && !h->bn192) L->BN = 128;
  DCCN_CHECK((K * 4) % 16 == 0, "layer K=%d is not a multiple of 4", K);
  if (!L->built) {
    const size_t bias_bytes = (size_t)((N + 127) / 128 * 128) * 4;   // epilogues read it in 32-float vector chunks
    int rc = dev_alloc(h, (void**)&L->dBias, bias_bytes);
    if (rc) return rc;
    DCCN_CUDA_OK(cudaMemsetAsync(L->dBias, 0, bias_bytes, s));
    if (prec == DCCN_PREC_EXACT) {
      rc = dev_alloc(h, (void**)&L->dW, (size_t)K * N * 4);
      if (rc) return rc;
    } else {
      rc = dev_alloc(h, (void**)&L->dWt0, (size_t)K * N * 4);
      if (rc) return rc;
      if (prec == DCCN_PREC_PARITY) {
        rc = dev_alloc(h, (void**)&L->dWt1, (size_t)K * N * 4);
        if (rc) return rc;
      }
    }
    L->built = true;
  }
  DCCN_CUDA_OK(cudaMemcpyAsync(L->dBias, L->bias.data(), (size_t)N * 4, cudaMemcpyHostToDevice, s));
  if (prec == DCCN_PREC_EXACT) {
    DCCN_CUDA_OK(cudaMemcpyAsync(L->dW, L->W.data(), (size_t)K * N * 4, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
  }
  std::vector<float> t0((size_t)K * N), t1;
  if (prec == DCCN_PREC_PARITY) t1.resize((size_t)K * N);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const float w = L->W[(size_t)k * N + n];
      if (prec == DCCN_PREC_PARITY) {
        const float hi = tf32_rna_host(w);
        t0[(size_t)n * K + k] = hi;
        t1[(size_t)n * K + k] = tf32_rna_host(w - hi);
      } else {
        t0[(size_t)n * K + k] = w;
      }
    }
  DCCN_CUDA_OK(cudaMemcpyAsync(L->dWt0, t0.data(), t0.size() * 4, cudaMemcpyHostToDevice, s));
  if (prec == DCCN_PREC_PARITY)
    DCCN_CUDA_OK(cudaMemcpyAsync(L->dWt1, t1.data(), t1.size() * 4, cudaMemcpyHostToDevice, s));
  DCCN_CUDA_OK(cudaStreamSynchronize(s));
  // multicast pairs: parity mode, A-in-TMEM tiles of width 128, and enough K to amortise the cluster syncs
  L->mc = prec == DCCN_PREC_PARITY && h->a_tmem && h->multicast && L->BN == 128 && K >= h->mc_min_k;
  const int box = L->mc ? L->BN / 2 : L->BN;
  int rc = make_tmap(&L->tmB0, L->dWt0, N, K, K, box);
  if (rc) return rc;
  if (prec == DCCN_PREC_PARITY) rc = make_tmap(&L->tmB1, L->dWt1, N, K, K, box);
  if (rc) return rc;
  // DCCN_F16X3 (staged): fp16 hi / lo planes of W * 2^e for the kind::f16 form (128- and 32-wide A-in-TMEM tiles)
  L->f16_ok = false;
  if (h->f16x3 && prec == DCCN_PREC_PARITY && h->a_tmem && !L->mc && (L->BN == 128 || L->BN == 32) && K % 8 == 0) {
    float wmax = 0.f;
    for (float w : L->W) wmax = fmaxf(wmax, fabsf(w));
    int e = 0;
    if (wmax > 0.f && std::isfinite(wmax)) e = 13 - (int)floorf(log2f(wmax));
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    const float sc = ldexpf(1.f, e);
    L->w_scale_inv = ldexpf(1.f, -e);
    std::vector<__half> h0((size_t)K * N), h1((size_t)K * N);
    for (int k = 0; k < K; ++k)
      for (int n = 0; n < N; ++n) {
        const float w = L->W[(size_t)k * N + n] * sc;
        const __half hi = __float2half_rn(w);
        h0[(size_t)n * K + k] = hi;
        h1[(size_t)n * K + k] = __float2half_rn(w - __half2float(hi));
      }
    if (!L->dWh0) {
      if ((rc = dev_alloc(h, &L->dWh0, (size_t)K * N * 2))) return rc;
      if ((rc = dev_alloc(h, &L->dWh1, (size_t)K * N * 2))) return rc;
    }
    DCCN_CUDA_OK(cudaMemcpyAsync(L->dWh0, h0.data(), h0.size() * 2, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaMemcpyAsync(L->dWh1, h1.data(), h1.size() * 2, cudaMemcpyHostToDevice, s));
    DCCN_CUDA_OK(cudaStreamSynchronize(s));
    if ((rc = make_tmap_f16(&L->tmH0, L->dWh0, N, K, K, L->BN))) return rc;
    if ((rc = make_tmap_f16(&L->tmH1, L->dWh1, N, K, K, L->BN))) return rc;
    L->f16_ok = true;
  }
  return rc;
}

#define NEED(var, name)                                           \
  const HostTensor* var = find(h, name);                          \
  DCCN_CHECK(var != nullptr, "weight '%s' was not set", name)

// (S,K) 'same' complex conv with one filter -> dense Toeplitz [S*K*2, S*K*2]      (dev/py/model.py:426)
static int pack_toeplitz_layer(dccn_handle* h, const char* kn, const char* bn, GemmLayer* L) {
  const int S = h->S, K = h->K, SK2 = S * K * 2;
  NEED(k, kn);
  NEED(b, bn);
  const bool vec = h->eqs.vector != 0;   // layers_conv2d_vector: kernel [S,K,2,1,2], re / im = channels 0 / 1 at IQ position 0
  DCCN_CHECK(k->shape.size() == 5 && k->shape[0] == S && k->shape[1] == K && k->shape[4] == 2 &&
                 k->shape[2] == (vec ? 2 : 1),
             "%s: expected [%d,%d,%d,1,2]", kn, S, K, vec ? 2 : 1);
  L->K = SK2; L->N = SK2;
  L->W.assign((size_t)SK2 * SK2, 0.f);
  const int pl = (S - 1) / 2, pw = (K - 1) / 2;
  if (vec) {
    for (int d = 0; d < S; ++d)
      for (int hh = 0; hh < K; ++hh) {
        const int co = (d * K + hh) * 2;
        for (int i = 0; i < S; ++i) {
          const int di = d + i - pl;
          if (di < 0 || di >= S) continue;
          for (int j = 0; j < K; ++j) {
            const int hj = hh + j - pw;
            if (hj < 0 || hj >= K) continue;
            const float* kp = &k->data[(size_t)(i * K + j) * 4];      // [iq][ch]
            const int ri = (di * K + hj) * 2;
            L->W[(size_t)ri * SK2 + co] = kp[0];
            L->W[(size_t)ri * SK2 + co + 1] = kp[1];
            L->W[(size_t)(ri + 1) * SK2 + co] = kp[2];
            L->W[(size_t)(ri + 1) * SK2 + co + 1] = kp[3];
          }
        }
      }
    L->bias.resize(SK2);
    for (int i = 0; i < SK2; i += 2) {
      L->bias[i] = b->data[0];
      L->bias[i + 1] = b->data[1];
    }
    return 0;
  }
  for (int d = 0; d < S; ++d)
    for (int hh = 0; hh < K; ++hh) {
      const int co = (d * K + hh) * 2;                 // output column (re)
      for (int i = 0; i < S; ++i) {
        const int di = d + i - pl;
        if (di < 0 || di >= S) continue;
        for (int j = 0; j < K; ++j) {
          const int hj = hh + j - pw;
          if (hj < 0 || hj >= K) continue;
          const float wa = k->data[(size_t)(i * K + j) * 2], wb = k->data[(size_t)(i * K + j) * 2 + 1];
          const int ri = (di * K + hj) * 2;            // input row (re)
          L->W[(size_t)ri * SK2 + co] = wa;
          L->W[(size_t)ri * SK2 + co + 1] = wb;
          L->W[(size_t)(ri + 1) * SK2 + co] = -wb;
          L->W[(size_t)(ri + 1) * SK2 + co + 1] = -wa;
        }
      }
    }
  L->bias.resize(SK2);
  for (int i = 0; i < SK2; i += 2) {
    L->bias[i] = b->data[0] - b->data[1];
    L->bias[i + 1] = b->data[1] - b->data[0];
  }
  return 0;
}

int pack_layers_host(dccn_handle* h) {
  const dccn_cfg& c = h->cfg;
  const int S = h->S, K = h->K, F = h->F, Tin = h->Tin, NB = h->NB, D = h->D;
  const int MO = 1 << NB;
  int rc;
  // ---- receiver ---------------------------------------------------------------
  {
    NEED(k, "fft_like/conv3d/kernel");
    NEED(b, "fft_like/conv3d/bias");
    // stored [1, Tk, 1, Tk, 2F]; 'same' over a width-1 axis => only tap (Tk-1)/2 is live
    DCCN_CHECK(k->shape.size() == 5 && k->shape[0] == 1 && k->shape[2] == 1 && k->shape[1] == Tin &&
                   k->shape[3] == Tin && k->shape[4] == 2 * F,
               "fft_like/conv3d/kernel: expected [1,%d,1,%d,%d]", Tin, Tin, 2 * F);
    const int tap = (Tin - 1) / 2;
    const float* base = k->data.data() + (size_t)tap * Tin * 2 * F;   // [Tin, 2F]
    pack_complex(base, base + F, 2 * F, Tin, F, b->data.data(), &h->r1);
  }
  {
    NEED(k, "demodulation/dense/kernel");
    NEED(b, "demodulation/dense/bias");
    DCCN_CHECK(k->shape.size() == 2 && k->shape[0] == S * F * 2 && k->shape[1] == 2 * D,
               "demodulation/dense/kernel: expected [%d,%d]", S * F * 2, 2 * D);
    pack_dense(*k, *b, &h->r2);
  }
  {
    NEED(kc, "demodulation/conv2d/kernel");
    NEED(bc, "demodulation/conv2d/bias");
    NEED(k1, "demodulation/dense_1/kernel");
    NEED(b1, "demodulation/dense_1/bias");
    DCCN_CHECK((int)kc->data.size() == 2 * MO && (int)k1->data.size() == (MO + 2) * 2 * NB,
               "demodulation head shapes do not match nbits=%d", NB);
    memset(&h->hw, 0, sizeof(h->hw));
    for (int i = 0; i < 2; ++i)
      for (int m = 0; m < MO; ++m) h->hw.Wc[i][m] = kc->data[i * MO + m];
    for (int m = 0; m < MO; ++m) h->hw.bc[m] = bc->data[m];
    for (int m = 0; m < MO + 2; ++m)
      for (int j = 0; j < 2 * NB; ++j) h->hw.W1[m][j] = k1->data[m * 2 * NB + j];
    for (int j = 0; j < 2 * NB; ++j) h->hw.b1[j] = b1->data[j];
    if (c.head == DCCN_HEAD_V1) {
      NEED(kc1, "demodulation/conv2d_1/kernel");
      NEED(bc1, "demodulation/conv2d_1/bias");
      DCCN_CHECK((int)kc1->data.size() == MO * MO, "demodulation/conv2d_1/kernel size");
      for (int m = 0; m < MO; ++m)
        for (int n = 0; n < MO; ++n) h->hw.Wc1[m][n] = kc1->data[m * MO + n];
      for (int m = 0; m < MO; ++m) h->hw.bc1[m] = bc1->data[m];
    }
  }
  h->r2.fused = h->fused_head != 0;
  h->g7.fused = true;
  if (!c.equalizer) return 0;
  const int SK2 = S * K * 2;
  if (h->eqs.generic) {
    // ---- ablation equalizers (dev/py/model.py:482-1084): same primitives, other wiring --------------------------
    // TF-1 auto-numbers the layers of a variable scope in creation order: dense, dense_1, ...; conv3d, conv3d_1, ...
    const EqSpec& sp = h->eqs;
    int nd = 0, ncv = 0;
    auto next_name = [](const char* base, int& ctr) {
      std::string n = std::string("Equalizer/") + base + (ctr == 0 ? "" : "_" + std::to_string(ctr));
      ++ctr;
      return n;
    };
    auto dense = [&](GemmLayer* L, int rows, int cols) -> int {
      const std::string n = next_name("dense", nd);
      const HostTensor* k = find(h, n + "/kernel");
      const HostTensor* b = find(h, n + "/bias");
      DCCN_CHECK(k && b, "weight '%s/{kernel,bias}' was not set (--opt=%d)", n.c_str(), h->eq_opt);
      DCCN_CHECK(k->shape.size() == 2 && k->shape[0] == rows && k->shape[1] == cols && (int)b->data.size() == cols,
                 "%s/kernel: expected [%d,%d] (--opt=%d)", n.c_str(), rows, cols, h->eq_opt);
      pack_dense(*k, *b, L);
      return 0;
    };
    if ((rc = dense(&h->g1, Tin * 2, 2 * K))) return rc;                                   // front1
    if (sp.front2_cconv) {                                                                 // front2
      const std::string n = next_name("conv3d", ncv);
      const HostTensor* k = find(h, n + "/kernel");
      const HostTensor* b = find(h, n + "/bias");
      DCCN_CHECK(k && b, "weight '%s/{kernel,bias}' was not set (--opt=%d)", n.c_str(), h->eq_opt);
      DCCN_CHECK(k->shape.size() == 5 && k->shape[0] == 1 && k->shape[1] == K && k->shape[3] == 1 && k->shape[4] == 2 * K,
                 "%s/kernel: expected [1,%d,1,1,%d]", n.c_str(), K, 2 * K);
      pack_complex(k->data.data(), k->data.data() + K, 2 * K, K, K, b->data.data(), &h->g2);
    } else if ((rc = dense(&h->g2, 2 * K, 2 * K))) return rc;
    if ((rc = dense(&h->g3, SK2, 2 * c.pilot_size))) return rc;                            // pilot bottleneck
    GemmLayer* chain[4] = {&h->g4, &h->g5, &h->g6, &h->gx0};
    for (int i = 0; i < sp.n_chain; ++i)
      if ((rc = dense(chain[i], i == 0 ? 2 * c.pilot_size : SK2, SK2))) return rc;
    if (sp.toeplitz) {
      const std::string n = next_name("conv3d", ncv);
      if ((rc = pack_toeplitz_layer(h, (n + "/kernel").c_str(), (n + "/bias").c_str(), &h->g7))) return rc;
    } else {
      chain[sp.n_chain - 1]->fused = true;                                                 // carries the phase equaliser
    }
    if (sp.tail == 1) {
      if ((rc = dense(&h->g9, 2 * K, 2 * K))) return rc;
    } else {
      // tf.ifft over the K subcarriers of a symbol as a constant [2K, 2K] layer (standard complex product, 1/K scale)
      GemmLayer* L = &h->g9;
      L->K = 2 * K; L->N = 2 * K;
      L->W.assign((size_t)4 * K * K, 0.f);
      L->bias.assign(2 * K, 0.f);
      const double PI2 = 6.283185307179586476925286766559;
      for (int kk = 0; kk < K; ++kk)
        for (int n = 0; n < K; ++n) {
          const double a = PI2 * (double)((kk * n) % K) / K;
          const float cs = (float)(cos(a) / K), sn = (float)(sin(a) / K);
          L->W[(size_t)(2 * kk) * 2 * K + 2 * n] = cs;
          L->W[(size_t)(2 * kk + 1) * 2 * K + 2 * n] = -sn;
          L->W[(size_t)(2 * kk) * 2 * K + 2 * n + 1] = sn;
          L->W[(size_t)(2 * kk + 1) * 2 * K + 2 * n + 1] = cs;
        }
    }
    return dense(&h->g10, 2 * K, 2 * h->T);
  }

  // ---- equalizer_ofdm -----------------------------------------------------------
  {
    NEED(k, "Equalizer/dense/kernel");
    NEED(b, "Equalizer/dense/bias");
    DCCN_CHECK(k->shape.size() == 2 && k->shape[0] == Tin * 2 && k->shape[1] == 2 * K, "Equalizer/dense/kernel shape");
    pack_dense(*k, *b, &h->g1);
  }
  const bool vec = h->eqs.vector != 0;
  auto pack_1xK = [&](const char* kn, const char* bn, GemmLayer* L, bool real_only) -> int {
    NEED(k, kn);
    NEED(b, bn);
    GemmLayer full;
    if (vec) {
      // layers_conv2d_vector, (1,K) 'valid' (complex.py:199-255): kernel [1,K,2,1,2F]; a plain real map of the 2K inputs
      // (k, iq) onto F real parts (channels [0,F)) and F imaginary parts (channels [F,2F)); no recombination
      DCCN_CHECK(k->shape.size() == 5 && k->shape[0] == 1 && k->shape[1] == K && k->shape[2] == 2 && k->shape[3] == 1 &&
                     k->shape[4] == 2 * K,
                 "%s: expected [1,%d,2,1,%d] (layers_conv2d_vector)", kn, K, 2 * K);
      full.K = 2 * K; full.N = 2 * K;
      full.W.resize((size_t)4 * K * K);
      full.bias.resize(2 * K);
      for (int kk = 0; kk < K; ++kk)
        for (int iq = 0; iq < 2; ++iq)
          for (int f = 0; f < K; ++f)
            for (int part = 0; part < 2; ++part)
              full.W[(size_t)(2 * kk + iq) * 2 * K + 2 * f + part] = k->data[((size_t)kk * 2 + iq) * 2 * K + part * K + f];
      for (int f = 0; f < K; ++f) {
        full.bias[2 * f] = b->data[f];
        full.bias[2 * f + 1] = b->data[K + f];
      }
    } else {
    DCCN_CHECK(k->shape.size() == 5 && k->shape[0] == 1 && k->shape[1] == K && k->shape[3] == 1 &&
                   k->shape[4] == 2 * K,
               "%s: expected [1,%d,1,1,%d]", kn, K, 2 * K);
    pack_complex(k->data.data(), k->data.data() + K, 2 * K, K, K, b->data.data(), &full);
    }
    if (!real_only) {
      L->K = full.K; L->N = full.N; L->W = full.W; L->bias = full.bias;
    } else {   // input imag part is identically 0: keep the xr rows only
      L->K = K; L->N = full.N; L->bias = full.bias;
      L->W.resize((size_t)K * full.N);
      for (int kk = 0; kk < K; ++kk)
        memcpy(&L->W[(size_t)kk * full.N], &full.W[(size_t)(2 * kk) * full.N], (size_t)full.N * 4);
    }
    return 0;
  };
  if ((rc = pack_1xK("Equalizer/conv3d/kernel", "Equalizer/conv3d/bias", &h->g2, false))) return rc;
  {
    NEED(k, "Equalizer/dense_1/kernel");
    NEED(b, "Equalizer/dense_1/bias");
    DCCN_CHECK(k->shape[0] == SK2 && k->shape[1] == 2 * c.pilot_size, "Equalizer/dense_1/kernel shape");
    pack_dense(*k, *b, &h->g3);
  }
  {
    NEED(k, "Equalizer/dense_2/kernel");
    NEED(b, "Equalizer/dense_2/bias");
    DCCN_CHECK(k->shape[0] == 2 * c.pilot_size && k->shape[1] == SK2, "Equalizer/dense_2/kernel shape");
    pack_dense(*k, *b, &h->g4);
  }
  {
    NEED(k, "Equalizer/dense_3/kernel");
    NEED(b, "Equalizer/dense_3/bias");
    DCCN_CHECK(k->shape[0] == SK2 && k->shape[1] == SK2, "Equalizer/dense_3/kernel shape");
    pack_dense(*k, *b, &h->g5);
  }
  {
    NEED(k, "Equalizer/dense_4/kernel");
    NEED(b, "Equalizer/dense_4/bias");
    DCCN_CHECK(k->shape[0] == SK2 && k->shape[1] == SK2, "Equalizer/dense_4/kernel shape");
    pack_dense(*k, *b, &h->g6);
  }
  if ((rc = pack_toeplitz_layer(h, "Equalizer/conv3d_1/kernel", "Equalizer/conv3d_1/bias", &h->g7))) return rc;
  if ((rc = pack_1xK("Equalizer/conv3d_2/kernel", "Equalizer/conv3d_2/bias", &h->g8, true))) return rc;
  if ((rc = pack_1xK("Equalizer/conv3d_3/kernel", "Equalizer/conv3d_3/bias", &h->g9, false))) return rc;
  {
    // dense_5 input is concat per subcarrier (eq_re, eq_im, corr_re, corr_im) -> row h*4+c.
    // Our producers write [eq_out (2K) | corr_out (2K)]; permute the rows accordingly.
    NEED(k, "Equalizer/dense_5/kernel");
    NEED(b, "Equalizer/dense_5/bias");
    DCCN_CHECK(k->shape[0] == 4 * K && k->shape[1] == 2 * h->T, "Equalizer/dense_5/kernel shape");
    GemmLayer* L = &h->g10;
    L->K = 4 * K; L->N = 2 * h->T;
    L->W.resize((size_t)L->K * L->N);
    for (int hh = 0; hh < K; ++hh)
      for (int cc = 0; cc < 4; ++cc) {
        const int src = hh * 4 + cc;
        const int dst = (cc < 2 ? 0 : 2 * K) + hh * 2 + (cc & 1);
        memcpy(&L->W[(size_t)dst * L->N], &k->data[(size_t)src * L->N], (size_t)L->N * 4);
      }
    L->bias = b->data;
  }
  return 0;
}

static int build_layers(dccn_handle* h, cudaStream_t s) {
  int rc = pack_layers_host(h);
  if (rc) return rc;
  if ((rc = upload_layer(h, &h->r1, s))) return rc;
  if ((rc = upload_layer(h, &h->r2, s))) return rc;
  if (!h->cfg.equalizer) return 0;
  if (h->eqs.generic) {
    const EqSpec& sp = h->eqs;
    std::vector<GemmLayer*> ls = {&h->g1, &h->g2, &h->g3, &h->g9, &h->g10};
    GemmLayer* chain[4] = {&h->g4, &h->g5, &h->g6, &h->gx0};
    for (int i = 0; i < sp.n_chain; ++i) ls.push_back(chain[i]);
    if (sp.toeplitz) ls.push_back(&h->g7);
    for (GemmLayer* L : ls)
      if ((rc = upload_layer(h, L, s))) return rc;
    return 0;
  }
  GemmLayer* ls[] = {&h->g1, &h->g2, &h->g3, &h->g4, &h->g5, &h->g6, &h->g7, &h->g8, &h->g9, &h->g10};
  for (GemmLayer* L : ls)
    if ((rc = upload_layer(h, L, s))) return rc;
  return 0;
}

// ---------------------------------------------------------------------------------------
// DCCN_FWD_FOLDED: pre-multiply consecutive linear layers (fp64 on the host, once per commit)
// ---------------------------------------------------------------------------------------
typedef std::vector<double> DVec;
static DVec to_d(const std::vector<float>& v) { return DVec(v.begin(), v.end()); }
// C[m,n] = A[m,k] * B[k,n]
static DVec matmul_d(const DVec& A, int m, int k, const DVec& B, int n) {
  DVec C((size_t)m * n, 0.0);
  for (int i = 0; i < m; ++i)
    for (int l = 0; l < k; ++l) {
      const double a = A[(size_t)i * k + l];
      if (a == 0.0) continue;
      const double* b = &B[(size_t)l * n];
      double* c = &C[(size_t)i * n];
      for (int j = 0; j < n; ++j) c[j] += a * b[j];
    }
  return C;
}
static void set_folded(GemmLayer* L, int K, int N, const DVec& W, const DVec& b) {
  L->K = K;
  L->N = N;
  L->W.assign(W.begin(), W.end());
  L->bias.assign(b.begin(), b.end());
  L->fused = false;
}

static int build_folded(dccn_handle* h, cudaStream_t s) {
  const int S = h->S, K = h->K, T = h->T, Tin = h->Tin, F = h->F;
  const int SK2 = S * K * 2, cp_off = (T - Tin) * 2, P2 = 2 * h->cfg.pilot_size;
  // f1 = dense . conv3d  (model.py:370-379):  [2Tin] -> [2K]
  {
    DVec W = matmul_d(to_d(h->g1.W), 2 * Tin, 2 * K, to_d(h->g2.W), 2 * K);
    DVec b = matmul_d(to_d(h->g1.bias), 1, 2 * K, to_d(h->g2.W), 2 * K);
    for (int j = 0; j < 2 * K; ++j) b[j] += h->g2.bias[j];
    set_folded(&h->f1, 2 * Tin, 2 * K, W, b);
  }
  // f4 = dense_2 . dense_3 . dense_4 (the tanh of dense_4 stays in the epilogue, model.py:401-424):  [2*pilot] -> [SK2]
  {
    DVec W5 = to_d(h->g5.W), W6 = to_d(h->g6.W);
    DVec W = matmul_d(matmul_d(to_d(h->g4.W), P2, SK2, W5, SK2), P2, SK2, W6, SK2);
    DVec b = matmul_d(to_d(h->g4.bias), 1, SK2, W5, SK2);
    for (int j = 0; j < SK2; ++j) b[j] += h->g5.bias[j];
    b = matmul_d(b, 1, SK2, W6, SK2);
    for (int j = 0; j < SK2; ++j) b[j] += h->g6.bias[j];
    set_folded(&h->f4, P2, SK2, W, b);
  }
  // f9 = [conv3d_3 (eq) ; conv3d_2 (corr, real rows)] . dense_5 . (CP slice) . fft_like  (model.py:437-462, 1246-1264):
  //      per symbol [eq (2K) | corr (K)] -> [2F]
  {
    // dense_5 restricted to the columns the receiver consumes
    DVec W10((size_t)4 * K * 2 * Tin), b10(2 * Tin);
    for (int r = 0; r < 4 * K; ++r)
      for (int c = 0; c < 2 * Tin; ++c) W10[(size_t)r * 2 * Tin + c] = h->g10.W[(size_t)r * 2 * T + cp_off + c];
    for (int c = 0; c < 2 * Tin; ++c) b10[c] = h->g10.bias[cp_off + c];
    DVec top(W10.begin(), W10.begin() + (size_t)2 * K * 2 * Tin), bot(W10.begin() + (size_t)2 * K * 2 * Tin, W10.end());
    DVec A9 = matmul_d(to_d(h->g9.W), 2 * K, 2 * K, top, 2 * Tin);      // [2K, 2Tin]
    DVec A8 = matmul_d(to_d(h->g8.W), K, 2 * K, bot, 2 * Tin);          // [K, 2Tin]
    DVec bm = matmul_d(to_d(h->g9.bias), 1, 2 * K, top, 2 * Tin);
    DVec bm8 = matmul_d(to_d(h->g8.bias), 1, 2 * K, bot, 2 * Tin);
    for (int c = 0; c < 2 * Tin; ++c) bm[c] += bm8[c] + b10[c];
    DVec mid((size_t)3 * K * 2 * Tin);
    std::copy(A9.begin(), A9.end(), mid.begin());
    std::copy(A8.begin(), A8.end(), mid.begin() + A9.size());
    DVec R = to_d(h->r1.W);                                             // [2Tin, 2F]
    DVec W = matmul_d(mid, 3 * K, 2 * Tin, R, 2 * F);
    DVec b = matmul_d(bm, 1, 2 * Tin, R, 2 * F);
    for (int j = 0; j < 2 * F; ++j) b[j] += h->r1.bias[j];
    set_folded(&h->f9, 3 * K, 2 * F, W, b);
  }
  int rc;
  if ((rc = upload_layer(h, &h->f1, s))) return rc;
  if ((rc = upload_layer(h, &h->f4, s))) return rc;
  if ((rc = upload_layer(h, &h->f9, s))) return rc;
  h->fold_built = true;
  return 0;
}

// ---------------------------------------------------------------------------------------
// GEMM dispatch
// ---------------------------------------------------------------------------------------
template <class Epi>
static int run_gemm(dccn_handle* h, int slot, const GemmLayer& L, const Act& A, int a_col_off, int64_t M,
                    const Epi& epi_in, cudaStream_t s, KSched ks = KSched()) {
  const int prec = h->cfg.precision;
#ifdef DCCN_TRACE
  {   // debug build: only the GEMM of profile slot $DCCN_TRACE_SLOT writes the timeline buffer
    static int want = getenv("DCCN_TRACE_SLOT") ? atoi(getenv("DCCN_TRACE_SLOT")) : -1;
    long long* p = (slot == want) ? g_trace_host_ptr : nullptr;
    cudaMemcpyToSymbolAsync(g_trace_buf, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s);
    static int abl_slot = getenv("DCCN_ABL_SLOT") ? atoi(getenv("DCCN_ABL_SLOT")) : -1;
    static int abl_copy[2];
    abl_copy[0] = 0;
    abl_copy[1] = g_abl_host;
    cudaMemcpyToSymbolAsync(g_abl_dev, &abl_copy[(abl_slot < 0 || abl_slot == slot) ? 1 : 0], sizeof(int), 0,
                            cudaMemcpyHostToDevice, s);
  }
#endif
  LaunchScope ls(h, slot, s);
  if (prec == DCCN_PREC_EXACT)
    return launch_gemm_simt<Epi>(A.p0 + a_col_off, A.p1 ? A.p1 + a_col_off : nullptr, A.ld, L.dW, (int)M, L.N, L.K,
                                 epi_in, s);
  TcOperands op;
  int rc = make_tmap(&op.a0, A.p0 + a_col_off, M, L.K, A.ld, 128);
  if (rc) return rc;
  // destinations of the epilogue's bulk tensor stores (32 x 32 boxes)
  Epi epi = epi_in;
  if constexpr (std::is_same<Epi, EpiStore>::value) {
    if ((rc = make_tmap(&epi.tm_out, epi.out.p0 + epi.out.col_off, epi.M, epi.N, epi.out.ld, 32))) return rc;
    if (epi.aux && (rc = make_tmap(&epi.tm_aux, epi.aux, epi.M, epi.N, epi.aux_ld, 32))) return rc;
  } else if constexpr (std::is_same<Epi, EpiPhaseEq>::value || std::is_same<Epi, EpiPhaseEqSym>::value) {
    if ((rc = make_tmap(&epi.tm_eq, epi.eq.p0 + epi.eq.col_off, epi.M, epi.eq.ld - epi.eq.col_off, epi.eq.ld, 32)))
      return rc;
    if (epi.chest_out && (rc = make_tmap(&epi.tm_chest, epi.chest_out, epi.M, epi.N, epi.N, 32))) return rc;
    epi.pf = h->epi_prefetch;
    if (epi.corr.p0 && (rc = make_tmap(&epi.tm_corr, epi.corr.p0 + epi.corr.col_off, epi.M, epi.corr.ld - epi.corr.col_off,
                                       epi.corr.ld, 32)))
      return rc;
  }
  op.b0 = L.tmB0;
  const bool split = prec == DCCN_PREC_PARITY;
  if (split) op.b1 = L.tmB1;
  // MMA order inside a k-block (gemm_tc.cuh `small_first`): cross terms first for every forward / dgrad GEMM; the
  // split-K weight-gradient GEMMs (contraction over the batch, heavy cancellation) measured better interleaved
  const int small_first = (h->small_first && ks.ksplit <= 1) ? 1 : 0;
  // DCCN_F16X3 (staged, inference only: the training step re-derives only the tf32 planes on the device)
  if constexpr (std::is_same<Epi, EpiStore>::value || std::is_same<Epi, EpiPhaseEq>::value ||
                std::is_same<Epi, EpiPhaseEqSym>::value) {
    if (split && h->f16x3 && L.f16_ok && !h->tr && !h->train_fwd && ks.ksplit <= 1) {
      op.b0 = L.tmH0;
      op.b1 = L.tmH1;
      const int kc = h->kc;
      if constexpr (std::is_same<Epi, EpiStore>::value) {
        if (L.BN == 32)
          return launch_gemm_tc<32, true, 1, true, false, Epi, true>(op, (int)M, L.N, L.K, kc, epi, s, h->num_sms, ks, 1.f,
                                                                     L.w_scale_inv, A.amax, small_first);
      }
      if (L.BN == 128)
        return launch_gemm_tc<128, true, 2, true, false, Epi, true>(op, (int)M, L.N, L.K, kc, epi, s, h->num_sms, ks, 1.f,
                                                                    L.w_scale_inv, A.amax, small_first);
      op.b0 = L.tmB0;
      op.b1 = L.tmB1;
    }
  }
  // parity mode on 128-wide tiles: A staged in TMEM (TS-form MMA) + cta_group::2 CTA pairs (half a weight tile per SM)
#define DCCN_TC_PAR(BNV, CGV)                                                                                \
  do {                                                                                                       \
    if (L.mc) return launch_gemm_tc<BNV, true, CGV, true, true, Epi>(op, (int)M, L.N, L.K, h->kc, epi, s, h->num_sms, ks);  \
    if (h->a_tmem) return launch_gemm_tc<BNV, true, CGV, true, false, Epi>(op, (int)M, L.N, L.K, h->kc, epi, s, h->num_sms, ks, 1.f, 1.f, nullptr, small_first); \
    return launch_gemm_tc<BNV, true, CGV, false, false, Epi>(op, (int)M, L.N, L.K, h->kc, epi, s, h->num_sms, ks);  \
  } while (0)
#define DCCN_TC_SS(BNV, CGV)                                                                                 \
  return split ? launch_gemm_tc<BNV, true, CGV, false, false, Epi>(op, (int)M, L.N, L.K, h->kc, epi, s, h->num_sms, ks) \
               : launch_gemm_tc<BNV, false, CGV, false, false, Epi>(op, (int)M, L.N, L.K, 0, epi, s, h->num_sms, ks)
  if constexpr (std::is_same<Epi, EpiStore>::value) {
    switch (L.BN) {
      case 32: if (split) DCCN_TC_PAR(32, 1); DCCN_TC_SS(32, 1);
      case 192: DCCN_TC_SS(192, 2);
      case 256: DCCN_TC_SS(256, 2);
      default: if (split) DCCN_TC_PAR(128, 2); DCCN_TC_SS(128, 2);
    }
  } else {   // fused phase-equaliser / demod-head epilogues only exist for 128-wide tiles
    DCCN_CHECK(L.BN == 128, "fused epilogue expects BN=128 (N=%d)", L.N);
    if (split) DCCN_TC_PAR(128, 2);
    DCCN_TC_SS(128, 2);
  }
#undef DCCN_TC_PAR
#undef DCCN_TC_SS
}

int run_gemm_store(dccn_handle* h, int slot, const GemmLayer& L, const Act& A, int a_col_off, int64_t M,
                   const EpiStore& epi, cudaStream_t s, int ksplit) {
  KSched ks;
  ks.ksplit = ksplit;
  return run_gemm<EpiStore>(h, slot, L, A, a_col_off, M, epi, s, ks);
}

template <int NB, bool V1>
static int run_head(dccn_handle* h, int64_t Bc, const uint8_t* bits, float* soft, uint8_t* hard,
                    unsigned long long* conf, double* ce, cudaStream_t s) {
  EpiHead<NB, V1> e;
  e.bias = h->fused_head ? h->r2.dBias : nullptr;
  e.hw = h->hw;
  e.bits = bits;
  e.soft = soft;
  e.hard = hard;
  e.conf = conf;
  e.ce_sum = ce;
  e.M = (int)Bc;
  e.N = h->r2.N;
  if (h->fused_head) return run_gemm(h, SLOT_R2_HEAD, h->r2, h->r1o, 0, Bc, e, s);
  // default: plain-store GEMM (bias added, one fp32 plane) + the full-occupancy head kernel
  Act oiq = h->out_iq;
  int rc = run_gemm(h, SLOT_R2_GEMM, h->r2, h->r1o, 0, Bc, store_epi(h->r2, oiq, 0, Bc), s);
  if (rc) return rc;
  LaunchScope ls(h, SLOT_R2_HEAD, s);
  // one block per frame at a time (thread <-> data subcarrier), as many resident blocks as fit
  const int D = h->r2.N >> 1;
  const int subs = h->head_subs == 2 ? 2 : 1;
  int threads = ((D + subs - 1) / subs + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  long long blocks = (long long)h->num_sms * (h->head_blocks > 0 ? h->head_blocks : 2048 / threads);
  if (blocks > Bc) blocks = Bc;
  if (subs == 2) head_kernel<NB, V1, 2><<<(unsigned)blocks, threads, 0, s>>>(oiq.p0, e);
  else head_kernel<NB, V1, 1><<<(unsigned)blocks, threads, 0, s>>>(oiq.p0, e);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_head_dispatch(dccn_handle* h, int64_t Bc, const uint8_t* bits, float* soft, uint8_t* hard,
                             unsigned long long* conf, double* ce, cudaStream_t s) {
  const bool v1 = h->cfg.head == DCCN_HEAD_V1;
  switch (h->NB) {
    case 1: return v1 ? run_head<1, true>(h, Bc, bits, soft, hard, conf, ce, s) : run_head<1, false>(h, Bc, bits, soft, hard, conf, ce, s);
    case 2: return v1 ? run_head<2, true>(h, Bc, bits, soft, hard, conf, ce, s) : run_head<2, false>(h, Bc, bits, soft, hard, conf, ce, s);
    case 3: return v1 ? run_head<3, true>(h, Bc, bits, soft, hard, conf, ce, s) : run_head<3, false>(h, Bc, bits, soft, hard, conf, ce, s);
    default: return v1 ? run_head<4, true>(h, Bc, bits, soft, hard, conf, ce, s) : run_head<4, false>(h, Bc, bits, soft, hard, conf, ce, s);
  }
}

int run_moments(dccn_handle* h, const float* x, int64_t B, float* mean, float* rstd, cudaStream_t s) {
  const int P = h->P;
  DCCN_CUDA_OK(cudaMemsetAsync(h->d_sums, 0, (size_t)2 * P * sizeof(double), s));
  const int gx = (P / 4 + 127) / 128;
  int gy = (int)((B + 15) / 16);
  const int gy_max = (h->num_sms * 8) / gx;
  if (gy > gy_max) gy = gy_max;
  if (gy < 1) gy = 1;
  LaunchScope ls(h, SLOT_MOMENTS, s, 2);
  moments_partial_kernel<<<dim3(gx, gy), 128, 0, s>>>(x, (long long)B, P, h->d_sums);
  moments_final_kernel<<<(P + 255) / 256, 256, 0, s>>>(h->d_sums, (long long)B, P, mean, rstd);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// chained per-symbol runs of equalizer_ofdm (chain.cu): same layers, same fp32 rounding per layer, intermediates in TMEM
// ---------------------------------------------------------------------------------------
static bool chain_layer_ok(const GemmLayer& L) { return L.f16_ok && L.BN == 128; }

static bool chain_enabled(const dccn_handle* h) {
  return h->chain && h->cfg.precision == DCCN_PREC_PARITY && h->f16x3 && h->a_tmem && h->kc == 1 && !h->tr &&
         !h->train_fwd && (h->eq_opt == 0 || h->eq_opt == 7);
}

static long long* g_chain_trace[2] = {nullptr, nullptr};   // tools/chain_trace.py: timeline buffers of the front / tail chain

static int chain_finish(dccn_handle* h, int slot, ChainParams& p, cudaStream_t s) {
  int rc;
  p.trace = g_chain_trace[slot == SLOT_CH_TAIL ? 1 : 0];
  EpiStore& e = p.epi;
  if ((rc = make_tmap(&e.tm_out, e.out.p0 + e.out.col_off, e.M, e.N, e.out.ld, 32))) return rc;
  if (e.aux && (rc = make_tmap(&e.tm_aux, e.aux, e.M, e.N, e.aux_ld, 32))) return rc;
  p.small_first = h->small_first ? 1 : 0;
  LaunchScope ls(h, slot, s);
  return launch_chain(p, s, h->num_sms);
}

// dense (2 Tin -> 2K) -> learned DFT (1,K) complex conv, per symbol                       model.py:370-379
static bool chain_front_ok(const dccn_handle* h) {
  return chain_enabled(h) && chain_layer_ok(h->g1) && chain_layer_ok(h->g2) && h->g1.N == 128 && h->g1.K <= 192 &&
         h->g2.K == 128 && h->g2.N <= 128;
}
static int run_chain_front(dccn_handle* h, const Act& a0v, int cp_off, int64_t MS, const Act& fv, cudaStream_t s) {
  ChainParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap(&p.tmA[0], a0v.p0 + cp_off, MS, h->g1.K, a0v.ld, 128))) return rc;
  p.tmA[1] = p.tmA[0];
  p.tmW[0][0] = h->g1.tmH0;  p.tmW[0][1] = h->g1.tmH1;
  p.tmW[1][0] = h->g2.tmH0;  p.tmW[1][1] = h->g2.tmH1;
  p.st[0] = ChainStage{0, (h->g1.K + 63) / 64, 0, 1, 0, h->g1.w_scale_inv, h->g1.dBias, a0v.amax};
  p.st[1] = ChainStage{-1, 2, 0, 1, -1, h->g2.w_scale_inv, nullptr, nullptr};
  p.nst = 2;
  p.M = (int)MS;
  p.epi = store_epi(h->g2, fv, 0, MS);
  return chain_finish(h, SLOT_CH_FRONT, p, s);
}

// (1,K) conv(eq) | (1,K) conv(corr) -> concat -> dense_5 (4K -> 2T), per symbol             model.py:437-462
static bool chain_tail_ok(const dccn_handle* h) {
  return chain_enabled(h) && chain_layer_ok(h->g8) && chain_layer_ok(h->g9) && chain_layer_ok(h->g10) &&
         h->g9.K == 128 && h->g9.N == 128 && h->g8.K == 64 && h->g8.N == 128 && h->g10.K == 256 && h->g10.N <= 256;
}
static int run_chain_tail(dccn_handle* h, const Act& eqv, const Act& corrv, int64_t MS, const Act& oeqv, float* eq_out,
                          int eq_out_ld, cudaStream_t s) {
  ChainParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_tmap(&p.tmA[0], eqv.p0, MS, h->g9.K, eqv.ld, 128))) return rc;
  if ((rc = make_tmap(&p.tmA[1], corrv.p0, MS, h->g8.K, corrv.ld, 128))) return rc;
  p.tmW[0][0] = h->g9.tmH0;   p.tmW[0][1] = h->g9.tmH1;
  p.tmW[1][0] = h->g8.tmH0;   p.tmW[1][1] = h->g8.tmH1;
  p.tmW[2][0] = h->g10.tmH0;  p.tmW[2][1] = h->g10.tmH1;
  // slots: eq staged in 0,1 -> conv3d_3 output (cat columns 0..2K) back into 0,1; corr staged in 2 -> conv3d_2 output
  // (cat columns 2K..4K) into 2,3; dense_5 contracts slots 0..3
  p.st[0] = ChainStage{0, 2, 0, 1, 0, h->g9.w_scale_inv, h->g9.dBias, eqv.amax};
  p.st[1] = ChainStage{1, 1, 2, 1, 2, h->g8.w_scale_inv, h->g8.dBias, corrv.amax};
  p.st[2] = ChainStage{-1, 4, 0, (h->g10.N + 127) / 128, -1, h->g10.w_scale_inv, nullptr, nullptr};
  p.nst = 3;
  p.M = (int)MS;
  p.epi = store_epi(h->g10, oeqv, 0, MS, 0, eq_out, eq_out_ld);
  return chain_finish(h, SLOT_CH_TAIL, p, s);
}

// one chunk of Bc frames through [equalizer ->] receiver
int run_chunk(dccn_handle* h, const float* x, int64_t Bc, const uint8_t* bits, float* soft, uint8_t* hard,
                     float* eq_out, float* chest_out, unsigned long long* conf, double* ce, int flags,
                     cudaStream_t s) {
  const bool use_eq = h->cfg.equalizer && !(flags & DCCN_FWD_SKIP_EQ);
  const int S = h->S, K = h->K, T = h->T, Tin = h->Tin, P = h->P;
  const int cp_off = (T - Tin) * 2;            // receiver / equalizer skip the CP when !use_cp
  int rc;
  DCCN_CUDA_OK(cudaMemsetAsync(h->d_amax, 0, kAmaxSlots * sizeof(unsigned), s));
  // ---- a2 (+ layer norm) ---------------------------------------------------------
  {
    const int warps_per_block = 8;
    const int grid = (int)((Bc + warps_per_block - 1) / warps_per_block);
    DCCN_CHECK(P <= 10 * 128, "frame of %d floats exceeds the prep kernel's register tile", P);
    LaunchScope ls(h, SLOT_PREP, s);
    prep_kernel<10><<<grid, 256, 0, s>>>(x, (long long)Bc, P, h->d_mean, h->d_rstd,
                                         (flags & DCCN_FWD_NO_NORM) ? 0 : 1, use_eq ? 1 : 0, out_of(h->a0), h->a0.amax);
    DCCN_CUDA_OK(cudaGetLastError());
  }
  const Act* rx_in = &h->a0;
  const bool folded = use_eq && h->eq_opt == 0 && (flags & DCCN_FWD_FOLDED) && !h->train_fwd && !h->tr && !eq_out &&
                      !(flags & DCCN_FWD_EQ_ONLY);
  if (folded) {
    // ---- same function, 5 GEMMs: consecutive linear layers were pre-multiplied (build_folded) ----------
    if (!h->fold_built && (rc = build_folded(h, s))) return rc;
    const int64_t MS = Bc * S;
    Act a0v = h->a0;  a0v.ld = 2 * T;
    Act fv = h->f;    fv.ld = 2 * K;
    Act eqcv = h->eqc;                                      // [MS, 3K]
    Act r1v = h->r1o; r1v.ld = 2 * h->F;
    if ((rc = run_gemm(h, SLOT_F1, h->f1, a0v, cp_off, MS, store_epi(h->f1, fv, 0, MS), s))) return rc;
    if ((rc = run_gemm(h, SLOT_G3, h->g3, h->f, 0, Bc, store_epi(h->g3, h->p32, 0, Bc), s))) return rc;
    if ((rc = run_gemm(h, SLOT_F4, h->f4, h->p32, 0, Bc, store_epi(h->f4, h->u1, 0, Bc, /*tanh*/ 1), s))) return rc;
    {
      EpiPhaseEqSym e;
      e.bias = h->g7.dBias;
      e.f0 = h->f.p0;
      e.f1 = h->f.p1;
      e.ld_f = h->f.ld;
      e.eq = ActOut{h->eqc.p0, nullptr, S * 3 * K, 0};      // row = frame; per symbol [eq | corr]
      e.corr = ActOut{h->eqc.p0, nullptr, S * 3 * K, 2 * K};
      e.amax_eq = e.amax_corr = h->eqc.amax;
      e.sym_cols = 2 * K;
      e.sym_stride = 3 * K;
      e.chest_out = chest_out;
      e.M = (int)Bc;
      e.N = h->g7.N;
      KSched ks;
      if (h->band_skip && h->g7.BN == 2 * K && (2 * K) % 32 == 0) ks.band = ((S - 1) / 2) * (2 * K / 32);
      if ((rc = run_gemm(h, SLOT_G7_PHASEEQ, h->g7, h->u1, 0, Bc, e, s, ks))) return rc;
    }
    if ((rc = run_gemm(h, SLOT_F9, h->f9, eqcv, 0, MS, store_epi(h->f9, r1v, 0, MS), s))) return rc;
    return run_head_dispatch(h, Bc, bits, soft, hard, conf, ce, s);
  }
  if (use_eq && h->eqs.generic) {
    // ---- ablation equalizers: front (2 per-symbol layers) -> pilot -> dense chain [-> Toeplitz] + phase equaliser -> tail
    const EqSpec& sp = h->eqs;
    const int64_t MS = Bc * S;
    Act a0v = h->a0;  a0v.ld = 2 * T;
    Act t1v = h->t1;  t1v.ld = 2 * K;
    Act fv = h->f;    fv.ld = 2 * K;
    Act eqv = h->eq;  eqv.ld = 2 * K;
    Act midv = h->cat; midv.ld = 2 * K;                 // tail intermediate [MS, 2K] (the cat buffer is [MS, 4K])
    Act oeqv = h->oeq; oeqv.ld = 2 * T;
    if ((rc = run_gemm(h, SLOT_G1, h->g1, a0v, cp_off, MS, store_epi(h->g1, t1v, 0, MS), s))) return rc;
    if ((rc = run_gemm(h, SLOT_G2, h->g2, t1v, 0, MS, store_epi(h->g2, fv, 0, MS), s))) return rc;
    if ((rc = run_gemm(h, SLOT_G3, h->g3, h->f, 0, Bc, store_epi(h->g3, h->p32, 0, Bc), s))) return rc;
    GemmLayer* chain[4] = {&h->g4, &h->g5, &h->g6, &h->gx0};
    const int chain_slot[4] = {SLOT_G4, SLOT_G5, SLOT_G6, SLOT_G6};
    const Act* src = &h->p32;
    // a training forward keeps every chain output (u1, u2, u3) and the channel estimate (chest_buf) for the backward pass
    const Act* keep[3] = {&h->u1, &h->u2, &h->u3};
    if (h->train_fwd && !chest_out) chest_out = h->chest_buf;
    auto phase_eq = [&](const GemmLayer& L, const Act& in, int act, bool band) -> int {
      EpiPhaseEq e;
      e.bias = L.dBias;
      e.f0 = h->f.p0;
      e.f1 = h->f.p1;
      e.ld_f = h->f.ld;
      e.eq = out_of(h->eq);
      e.amax_eq = h->eq.amax;
      e.corr = ActOut{nullptr, nullptr, 0, 0};          // no correlation branch in these graphs
      e.chest_out = chest_out;
      e.act = act;
      e.M = (int)Bc;
      e.N = L.N;
      KSched ks;
      if (band && h->band_skip && L.BN == 2 * K && (2 * K) % 32 == 0) ks.band = ((S - 1) / 2) * (2 * K / 32);
      return run_gemm(h, SLOT_G7_PHASEEQ, L, in, 0, Bc, e, s, ks);
    };
    for (int i = 0; i < sp.n_chain; ++i) {
      const bool last = (i == sp.n_chain - 1) && !sp.toeplitz;
      if (last) {
        if ((rc = phase_eq(*chain[i], *src, sp.chain_act[i], false))) return rc;
      } else {
        const Act* dst = h->train_fwd ? keep[i] : ((i & 1) ? &h->u2 : &h->u1);
        if ((rc = run_gemm(h, chain_slot[i], *chain[i], *src, 0, Bc, store_epi(*chain[i], *dst, 0, Bc, sp.chain_act[i]), s)))
          return rc;
        src = dst;
      }
    }
    if (sp.toeplitz && (rc = phase_eq(h->g7, *src, 0, true))) return rc;
    // tail: dense(2K) or the constant inverse DFT per symbol, then dense(2T) per symbol
    if ((rc = run_gemm(h, SLOT_G9, h->g9, eqv, 0, MS, store_epi(h->g9, midv, 0, MS), s))) return rc;
    if ((rc = run_gemm(h, SLOT_G10, h->g10, midv, 0, MS, store_epi(h->g10, oeqv, 0, MS, 0, eq_out, 2 * T), s))) return rc;
    rx_in = &h->oeq;
    if (flags & DCCN_FWD_EQ_ONLY) return 0;
  } else if (use_eq) {
    const int64_t MS = Bc * S;
    // views: [Bc, S*X] buffers are addressed as [Bc*S, X] by the per-symbol layers
    Act a0v = h->a0;  a0v.ld = 2 * T;
    Act t1v = h->t1;  t1v.ld = 2 * K;
    Act fv = h->f;    fv.ld = 2 * K;
    Act eqv = h->eq;  eqv.ld = 2 * K;
    Act corrv = h->corr; corrv.ld = K;
    Act catv = h->cat;   // [Bc*S, 4K]
    Act oeqv = h->oeq; oeqv.ld = 2 * T;
    // dense: per-symbol 2*Tin -> 2K                               model.py:370
    if (chain_front_ok(h)) {
      // both layers in one kernel, the [MS, 2K] intermediate stays in tensor memory (chain.cu)
      if ((rc = run_chain_front(h, a0v, cp_off, MS, fv, s))) return rc;
    } else {
      if ((rc = run_gemm(h, SLOT_G1, h->g1, a0v, cp_off, MS, store_epi(h->g1, t1v, 0, MS), s))) return rc;
      // learned DFT (1,K) 'valid' complex conv                       model.py:377-379
      if ((rc = run_gemm(h, SLOT_G2, h->g2, t1v, 0, MS, store_epi(h->g2, fv, 0, MS), s))) return rc;
    }
    // pilot bottleneck and channel-estimate MLP                    model.py:393-424
    if ((rc = run_gemm(h, SLOT_G3, h->g3, h->f, 0, Bc, store_epi(h->g3, h->p32, 0, Bc), s))) return rc;
    // (chain activations: linear, linear, tanh for equalizer_ofdm; tanh x3 for equalizer_separateIQ, model.py:1140-1162)
    if ((rc = run_gemm(h, SLOT_G4, h->g4, h->p32, 0, Bc, store_epi(h->g4, h->u1, 0, Bc, h->eqs.chain_act[0]), s))) return rc;
    if ((rc = run_gemm(h, SLOT_G5, h->g5, h->u1, 0, Bc, store_epi(h->g5, h->u2, 0, Bc, h->eqs.chain_act[1]), s))) return rc;
    // (a training forward keeps dense_2's output in u1 and writes the tanh output to u3)
    const Act& c4 = h->train_fwd ? h->u3 : h->u1;
    if (h->train_fwd && !chest_out) chest_out = h->chest_buf;
    if ((rc = run_gemm(h, SLOT_G6, h->g6, h->u2, 0, Bc, store_epi(h->g6, c4, 0, Bc, /*tanh*/ 1), s))) return rc;
    // (S,K) 'same' complex conv as Toeplitz GEMM + fused phase equaliser   model.py:426-437
    {
      EpiPhaseEq e;
      e.bias = h->g7.dBias;
      e.f0 = h->f.p0;
      e.f1 = h->f.p1;
      e.ld_f = h->f.ld;
      e.eq = out_of(h->eq);
      e.corr = out_of(h->corr);
      e.amax_eq = h->eq.amax;
      e.amax_corr = h->corr.amax;
      e.chest_out = chest_out;
      e.M = (int)Bc;
      e.N = h->g7.N;
      // block-banded Toeplitz operand: symbols more than (S-1)/2 apart share no tap -> skip those k-blocks
      KSched ks;
      if (h->band_skip && h->g7.BN == 2 * K && (2 * K) % 32 == 0) ks.band = ((S - 1) / 2) * (2 * K / 32);
      if ((rc = run_gemm(h, SLOT_G7_PHASEEQ, h->g7, c4, 0, Bc, e, s, ks))) return rc;
    }
    if (h->mon_snr_db) {      // snr_db monitor of this pass (reads the phase-equalised frame the epilogue just wrote)
      g_launches += 1;
      snr_monitor_kernel<<<(unsigned)((Bc + 7) / 8), 256, 0, s>>>((const float2*)h->eq.p0, (long long)Bc, S, K,
                                                                 h->mon_pilot_carriers, h->mon_n_pilot,
                                                                 h->mon_snr_db + h->mon_frame0);
      DCCN_CUDA_OK(cudaGetLastError());
    }
    // corr / eq (1,K) 'valid' complex convs -> [eq_out | corr_out]  model.py:437-448
    if (chain_tail_ok(h)) {
      // the three layers in one kernel: the concatenated [MS, 4K] tile goes to dense_5 through tensor memory (chain.cu)
      if ((rc = run_chain_tail(h, eqv, corrv, MS, oeqv, eq_out, 2 * T, s))) return rc;
    } else {
      if ((rc = run_gemm(h, SLOT_G8, h->g8, corrv, 0, MS, store_epi(h->g8, catv, 2 * K, MS), s))) return rc;
      if ((rc = run_gemm(h, SLOT_G9, h->g9, eqv, 0, MS, store_epi(h->g9, catv, 0, MS), s))) return rc;
      // dense_5: 4K -> 2T per symbol                                  model.py:457-462
      if ((rc = run_gemm(h, SLOT_G10, h->g10, catv, 0, MS, store_epi(h->g10, oeqv, 0, MS, 0, eq_out, 2 * T), s))) return rc;
    }
    rx_in = &h->oeq;
    if (flags & DCCN_FWD_EQ_ONLY) return 0;
  }
  // ---- ofdm_dense_rx -----------------------------------------------------------------
  {
    const int64_t MS = Bc * S;
    Act inv = *rx_in;  inv.ld = 2 * T;
    Act r1v = h->r1o;  r1v.ld = 2 * h->F;
    if ((rc = run_gemm(h, SLOT_R1, h->r1, inv, cp_off, MS, store_epi(h->r1, r1v, 0, MS), s))) return rc;
    if ((rc = run_head_dispatch(h, Bc, bits, soft, hard, conf, ce, s))) return rc;
  }
  return 0;
}

// ---- measurement aid: raw TMA -> shared-memory delivery rate of one SM (no MMA) ----------
// Each CTA's producer thread keeps `stages` slots of `boxes` [128 x 32 fp32] boxes in flight; a
// consumer thread frees a slot as soon as its bytes have landed.  Reports SM cycles per CTA.
__global__ void __launch_bounds__(64, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int rows, int kcols, int stages, int boxes, int iters,
                long long* clks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * boxes * 16384);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int row_tiles = rows / 128, k_tiles = kcols / 32;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int rt = blockIdx.x % row_tiles, kt = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], boxes * 16384);
      for (int b = 0; b < boxes; ++b)
        tma_load_2d(smem + (stage * boxes + b) * 16384, &tm, &full[stage], kt * 32, ((rt + b * 7) % row_tiles) * 128);
      if (++kt == k_tiles) {
        kt = 0;
        rt = (rt + gridDim.x) % row_tiles;
      }
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    clks[blockIdx.x] = clock64() - t0;
  }
}

// ---- measurement aid: tcgen05.mma issue / execution rate of one SM (no TMA, operands = whatever is in smem) ----
// mode 0: SS form, mode 1: TS form (A from TMEM).  `per_commit` MMAs between tcgen05.commit's (0 = one at the end).
template <int BN>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int n_mma, int per_commit, int mode, int dep, long long* clks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (mode >= 2 ? 131072 : 16384 + BN * 128));
  uint64_t* bar2 = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (mode >= 2 ? 131072 : 16384 + BN * 128) / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem)[i] = 0.f;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) {
      mbar_init(bar, 1);
      mbar_init(bar2, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tptr;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(BN);
    const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 16384));
    long long t0 = clock64();
    if (elect_one()) {
      const int batch = per_commit > 0 ? per_commit : n_mma;
      for (int i = 0; i < n_mma; i += batch) {
        const uint32_t d = tb + (dep ? 0u : (uint32_t)(((i / batch) & 1) * BN));   // dep=0: alternate two accumulators per batch
        if (mode >= 2) {
          // the parity GEMM's real operand pattern: batch i uses smem stage i % 4 (B_hi at +0, B_lo at +16 KB) and
          // TMEM staging columns 256 + 64 * (i % 4); k-step k advances 32 B / 8 columns.  mode 2: order lo*hi, hi*lo,
          // hi*hi (as shipped); mode 3: lo*hi, hi*hi, hi*lo (consecutive MMAs share an operand)
          const int st = (i / batch) & 3;
          const uint32_t bh = smem_u32(smem) + st * 32768, bl = bh + 16384, ta = tb + 256 + st * 64;
          for (int k = 0; k < batch / 3; ++k) {
            const uint64_t dbh = umma_desc_sw128(bh + (k & 3) * 32), dbl = umma_desc_sw128(bl + (k & 3) * 32);
            const uint32_t ka = ta + (k & 3) * 8;
            umma_tf32_ts(d, ka + 32, dbh, idesc, k ? 1u : 0u);
            if (mode == 2) {
              umma_tf32_ts(d, ka, dbl, idesc, 1u);
              umma_tf32_ts(d, ka, dbh, idesc, 1u);
            } else {
              umma_tf32_ts(d, ka, dbh, idesc, 1u);
              umma_tf32_ts(d, ka, dbl, idesc, 1u);
            }
          }
        } else
        for (int j = 0; j < batch; ++j) {
          if (mode == 0) umma_tf32(d, da, db, idesc, 1u);
          else umma_tf32_ts(d, tb + 256, db, idesc, 1u);
        }
        if (per_commit > 0) umma_commit(bar2);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clks[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

__global__ void conf_copy_kernel(const unsigned long long* src, long long* dst) {
  if (threadIdx.x < 4) dst[threadIdx.x] += (long long)src[threadIdx.x];
}

void conf_accumulate(const unsigned long long* src, int64_t* dst, cudaStream_t s) {
  g_launches += 1;
  conf_copy_kernel<<<1, 32, 0, s>>>(src, (long long*)dst);
}

}  // namespace dccn

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int dccn_abi_version(void) { return DCCN_ABI_VERSION; }
const char* dccn_last_error(void) { return g_last_error.c_str(); }

int dccn_create(const dccn_cfg* cfg, dccn_handle** out) {
  DCCN_CHECK(cfg && out, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(-3, "no CUDA device: libdccn has no CPU fallback");
  DCCN_CHECK(cfg->nbits >= 1 && cfg->nbits <= 4, "nbits must be 1..4");
  DCCN_CHECK(cfg->nfft > 0 && cfg->nfft % 16 == 0 && cfg->nsymbol > 0 && cfg->nfilter > 0 && cfg->nfilter % 16 == 0,
             "bad geometry");
  DCCN_CHECK(cfg->precision >= 0 && cfg->precision <= 2, "bad precision mode");
  dccn_handle* h = new dccn_handle();
  h->cfg = *cfg;
  DCCN_CUDA_OK(cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  DCCN_CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
  h->num_sms = prop.multiProcessorCount;
  if (cfg->precision != DCCN_PREC_EXACT && prop.major != 10) {
    delete h;
    return set_error(-3, "tcgen05 precision modes need an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
  }
  h->S = cfg->nsymbol;
  h->K = cfg->nfft;
  h->T = cfg->nfft + cfg->cp_len;
  h->Tin = cfg->use_cp ? h->T : h->K;
  h->F = cfg->nfilter;
  h->D = cfg->n_data;
  h->NB = cfg->nbits;
  h->P = h->S * h->T * 2;
  if (cfg->equalizer && cfg->eq_opt != 0) {
    // wiring of the ablation graphs (dev/py/ofdmreceiver_np_mp.py:292-311 -> dev/py/model.py:482-1084)
    EqSpec sp;
    switch (cfg->eq_opt) {
      case 1: sp.front2_cconv = 0; sp.n_chain = 3; sp.chain_act[0] = 0; sp.chain_act[1] = 0; sp.chain_act[2] = 1; sp.toeplitz = 1; sp.tail = 1; break;
      case 2: sp.front2_cconv = 1; sp.n_chain = 1; sp.chain_act[0] = 0; sp.toeplitz = 0; sp.tail = 2; break;
      case 4: sp.front2_cconv = 1; sp.n_chain = 2; sp.chain_act[0] = 0; sp.chain_act[1] = 1; sp.toeplitz = 0; sp.tail = 2; break;
      case 5: sp.front2_cconv = 1; sp.n_chain = 4; sp.chain_act[0] = 0; sp.chain_act[1] = 1; sp.chain_act[2] = 1; sp.chain_act[3] = 1; sp.toeplitz = 0; sp.tail = 2; break;
      case 3: sp.front2_cconv = 0; sp.n_chain = 4; sp.chain_act[0] = 1; sp.chain_act[1] = 1; sp.chain_act[2] = 1; sp.chain_act[3] = 1; sp.toeplitz = 0; sp.tail = 1; break;
      case 7: sp.front2_cconv = 1; sp.n_chain = 3; sp.chain_act[0] = 1; sp.chain_act[1] = 1; sp.chain_act[2] = 1; sp.toeplitz = 1; sp.tail = 0; sp.vector = 1; break;
      default:
        delete h;
        return set_error(-2, "eq_opt=%d: implemented equalizer graphs are --opt 0,1,2,3,4,5,7", (int)cfg->eq_opt);
    }
    sp.generic = cfg->eq_opt != 7;
    h->eqs = sp;
    h->eq_opt = cfg->eq_opt;
  }
  h->chunk = cfg->chunk_frames > 0 ? cfg->chunk_frames : 65536;   // per-launch overheads (~10 us x 15 kernels) amortise over the pass
  if (const char* e = getenv("DCCN_KC")) h->kc = atoi(e);
  if (const char* e = getenv("DCCN_SMALL_FIRST")) h->small_first = atoi(e);
  if (const char* e = getenv("DCCN_EPI_PREFETCH")) h->epi_prefetch = atoi(e);
  if (const char* e = getenv("DCCN_HEAD_SUBS")) h->head_subs = atoi(e);       // experiment knobs of the head kernel
  if (const char* e = getenv("DCCN_HEAD_BLOCKS")) h->head_blocks = atoi(e);
  if (const char* e = getenv("DCCN_BN_WIDE")) h->bn_wide = atoi(e);
  if (const char* e = getenv("DCCN_FUSED_HEAD")) h->fused_head = atoi(e);
  if (const char* e = getenv("DCCN_A_TMEM")) h->a_tmem = atoi(e);
  if (const char* e = getenv("DCCN_SMS")) h->num_sms = atoi(e);   // experiment knob: restrict the persistent grids
  if (const char* e = getenv("DCCN_PAIR")) h->multicast = atoi(e);
  if (const char* e = getenv("DCCN_MC_MIN_K")) h->mc_min_k = atoi(e);
  if (const char* e = getenv("DCCN_BAND")) h->band_skip = atoi(e);
  if (const char* e = getenv("DCCN_F16X3")) h->f16x3 = atoi(e);   // 0: tf32 hi/lo pairs (the round-1 form)
  if (const char* e = getenv("DCCN_CHAIN")) h->chain = atoi(e);   // 0: every per-symbol layer through HBM
  if (const char* e = getenv("DCCN_TX_V2")) h->tx_v2 = atoi(e);
  if (const char* e = getenv("DCCN_BN192")) h->bn192 = atoi(e);
  if (const char* e = getenv("DCCN_FOLD")) if (atoi(e)) h->default_flags |= DCCN_FWD_FOLDED;
  if (h->P % 4 != 0 || (2 * h->T) % 4 != 0) {
    delete h;
    return set_error(-2, "frame size must be a multiple of 4 floats");
  }
  if (cfg->equalizer && cfg->nfilter != cfg->nfft) {
    delete h;
    return set_error(-2, "equalizer_ofdm requires nfilter == nfft");
  }
  // ---- workspace -----------------------------------------------------------------
  // (the inter-layer buffers are allocated by ensure_workspace on first use)
  int rc = 0;
  rc |= dev_alloc(h, (void**)&h->d_sums, (size_t)2 * h->P * sizeof(double));
  rc |= dev_alloc(h, (void**)&h->d_mean, (size_t)h->P * 4);
  rc |= dev_alloc(h, (void**)&h->d_rstd, (size_t)h->P * 4);
  rc |= dev_alloc(h, (void**)&h->d_power, sizeof(double));
  rc |= dev_alloc(h, (void**)&h->d_conf, 4 * sizeof(unsigned long long));
  rc |= dev_alloc(h, (void**)&h->d_ce, sizeof(double));
  rc |= dev_alloc(h, (void**)&h->d_amax, kAmaxSlots * sizeof(unsigned));
  rc |= dev_alloc(h, (void**)&h->d_txmap, (size_t)h->S * h->K * sizeof(int32_t));
  for (int i = 0; i < 2 && !rc; ++i) {
    rc |= dev_alloc(h, (void**)&h->slot[i].d_conf, 4 * sizeof(int64_t));
    rc |= dev_alloc(h, (void**)&h->slot[i].d_ce, sizeof(double));
    if (cudaMallocHost((void**)&h->slot[i].h_conf, 4 * sizeof(int64_t)) != cudaSuccess ||
        cudaMallocHost((void**)&h->slot[i].h_ce, sizeof(double)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->slot[i].copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->slot[i].computed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->slot[i].done, cudaEventDisableTiming) != cudaSuccess)
      rc = set_error(-1, "host-slot allocation failed");
  }
  if (!rc && (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
              cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) != cudaSuccess))
    rc = set_error(-1, "cudaStreamCreate failed");
  if (rc) {
    dccn_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

void dccn_destroy(dccn_handle* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  train_free(h);
  for (int i = 0; i < 2; ++i) {
    if (h->slot[i].h_conf) cudaFreeHost(h->slot[i].h_conf);
    if (h->slot[i].h_ce) cudaFreeHost(h->slot[i].h_ce);
    if (h->slot[i].copied) cudaEventDestroy(h->slot[i].copied);
    if (h->slot[i].computed) cudaEventDestroy(h->slot[i].computed);
    if (h->slot[i].done) cudaEventDestroy(h->slot[i].done);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  for (void* p : h->allocs) cudaFree(p);
  for (void* p : h->ws_allocs) cudaFree(p);
  delete h;
}

size_t dccn_workspace_bytes(const dccn_handle* h) { return h ? h->ws_bytes : 0; }

int dccn_set_weight(dccn_handle* h, const char* tf_name, const float* host, const int64_t* shape, int rank) {
  DCCN_CHECK(h && tf_name && host && shape && rank >= 0 && rank <= 8, "bad argument");
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < rank; ++i) {
    t.shape.push_back(shape[i]);
    n *= shape[i];
  }
  t.data.assign(host, host + n);
  h->raw[tf_name] = std::move(t);
  h->committed = false;
  return 0;
}

int64_t dccn_get_weight(dccn_handle* h, const char* tf_name, float* host, int64_t capacity) {
  if (!h || !tf_name) return set_error(-2, "bad argument");
  auto it = h->raw.find(tf_name);
  if (it == h->raw.end()) return set_error(-2, "weight '%s' was not set", tf_name);
  if (h->tr && train_fetch_weight(h, tf_name, &it->second) < 0) return -1;   // trained value lives on the device
  const HostTensor* t = &it->second;
  const int64_t n = (int64_t)t->data.size();
  if (host) {
    if (capacity < n) return set_error(-2, "buffer too small for '%s'", tf_name);
    memcpy(host, t->data.data(), (size_t)n * 4);
  }
  return n;
}

int dccn_commit_weights(dccn_handle* h, void* stream) {
  DCCN_CHECK(h, "null handle");
  int rc = build_layers(h, (cudaStream_t)stream);
  if (rc) return rc;
  h->committed = true;
  h->fold_built = false;
  if (h->tr) return train_on_commit(h, (cudaStream_t)stream);   // re-seed the device master copies
  return 0;
}

int dccn_batch_moments(dccn_handle* h, const float* x_dev, int64_t B, float* mean_dev, float* rstd_dev,
                       void* stream) {
  DCCN_CHECK(h && x_dev && mean_dev && rstd_dev && B > 0, "bad argument");
  return run_moments(h, x_dev, B, mean_dev, rstd_dev, (cudaStream_t)stream);
}

int dccn_forward(dccn_handle* h, const float* x_dev, int64_t B, const uint8_t* bits_dev, float* soft_dev,
                 uint8_t* hard_dev, float* eq_dev, float* chest_dev, int64_t* conf_dev, double* ce_sum_dev,
                 int flags, void* stream) {
  DCCN_CHECK(h && x_dev, "null argument");
  flags |= h->default_flags;
  DCCN_CHECK(!(flags & DCCN_FWD_EQ_ONLY) || h->cfg.equalizer, "DCCN_FWD_EQ_ONLY needs cfg.equalizer");
  DCCN_CHECK(h->committed, "weights not committed (dccn_commit_weights)");
  DCCN_CHECK(B > 0, "empty batch");
  DCCN_CHECK(!(eq_dev || chest_dev) || h->cfg.equalizer, "eq/chest outputs need cfg.equalizer");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = 0;
  if (!(flags & DCCN_FWD_NO_NORM)) rc = run_moments(h, x_dev, B, h->d_mean, h->d_rstd, s);
  if (rc) return rc;
  const bool want_conf = bits_dev && conf_dev;
  if (want_conf) DCCN_CUDA_OK(cudaMemsetAsync(h->d_conf, 0, 4 * sizeof(unsigned long long), s));
  const int D = h->D, NB = h->NB;
  // equal passes of at most h->chunk frames (a 70 000-frame batch runs as 2 x 35 000, not 65 536 + 4 464)
  const int64_t n_pass = (B + h->chunk - 1) / h->chunk;
  const int64_t per = ((B + n_pass - 1) / n_pass + 127) / 128 * 128;
  if ((flags & DCCN_FWD_FOLDED) && h->cfg.equalizer && !h->ws_eqc) h->ws_eqc = h->ws_dirty = true;
  if ((rc = ensure_workspace(h, per))) return rc;
  DCCN_CHECK(!h->mon_snr_db || (h->cfg.equalizer && h->eq_opt == 0 && !(flags & (DCCN_FWD_SKIP_EQ | DCCN_FWD_FOLDED))),
             "the snr_db monitor belongs to equalizer_ofdm's layer-by-layer schedule");
  for (int64_t b0 = 0; b0 < B; b0 += per) {
    const int64_t Bc = (B - b0) < per ? (B - b0) : per;
    h->mon_frame0 = b0;
    rc = run_chunk(h, x_dev + (size_t)b0 * h->P, Bc, bits_dev ? bits_dev + (size_t)b0 * D * NB : nullptr,
                   soft_dev ? soft_dev + (size_t)b0 * D * NB * 2 : nullptr,
                   hard_dev ? hard_dev + (size_t)b0 * D * NB : nullptr,
                   eq_dev ? eq_dev + (size_t)b0 * h->P : nullptr,
                   chest_dev ? chest_dev + (size_t)b0 * h->S * h->K * 2 : nullptr,
                   want_conf ? h->d_conf : nullptr, (bits_dev && ce_sum_dev) ? ce_sum_dev : nullptr, flags, s);
    if (rc) return rc;
  }
  h->mon_snr_db = nullptr;       // one-shot request
  if (want_conf) {
    g_launches += 1;
    conf_copy_kernel<<<1, 32, 0, s>>>(h->d_conf, (long long*)conf_dev);
    DCCN_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int dccn_forward_monitors(dccn_handle* h, float* snr_db_dev, const int32_t* pilot_carriers_dev, int n_pilot_carriers) {
  DCCN_CHECK(h, "null handle");
  DCCN_CHECK(!snr_db_dev || (pilot_carriers_dev && n_pilot_carriers > 0), "pilot carrier list missing");
  h->mon_snr_db = snr_db_dev;
  h->mon_pilot_carriers = pilot_carriers_dev;
  h->mon_n_pilot = n_pilot_carriers;
  return 0;
}

int dccn_monitors(dccn_handle* h, const float* x_dev, int64_t B, const float* snr_db_dev, uint64_t seed, double* sums_dev,
                  float* input_dev, void* iq_tx_dev, void* iq_rx_dev, void* stream) {
  DCCN_CHECK(h && x_dev && sums_dev && B > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = run_moments(h, x_dev, B, h->d_mean, h->d_rstd, s);
  if (rc) return rc;
  DCCN_CUDA_OK(cudaMemsetAsync(sums_dev, 0, 2 * sizeof(double), s));
  long long blocks = (B + 7) / 8;
  if (blocks > (long long)h->num_sms * 8) blocks = (long long)h->num_sms * 8;
  g_launches += 1;
  monitor_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float2*)x_dev, (long long)B, h->S * h->T, (const float2*)h->d_mean,
                                                  (const float2*)h->d_rstd, snr_db_dev, seed, sums_dev, (float2*)input_dev,
                                                  (__half2*)iq_tx_dev, (__half2*)iq_rx_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

int dccn_forward_host_begin(dccn_handle* h, int slot, const float* x_host, int64_t B, const uint8_t* bits_host,
                            uint8_t* hard_host, void* stream) {
  DCCN_CHECK(h && x_host && B > 0 && (slot == 0 || slot == 1), "bad argument");
  dccn_handle::HostSlot& S = h->slot[slot];
  DCCN_CHECK(!S.busy, "slot %d still in flight: call dccn_forward_host_end first", slot);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nb = (size_t)h->D * h->NB;
  if (S.frames < B) {
    // (re)allocate staging; old buffers stay in the alloc list until destroy
    int rc = dev_alloc(h, (void**)&S.d_x, (size_t)B * h->P * 4);
    rc |= dev_alloc(h, (void**)&S.d_bits, (size_t)B * nb);
    rc |= dev_alloc(h, (void**)&S.d_hard, (size_t)B * nb);
    if (rc) return rc;
    S.frames = B;
  }
  // H2D on the copy stream (the previous pass over this slot was waited for in ..._end)
  DCCN_CUDA_OK(cudaMemcpyAsync(S.d_x, x_host, (size_t)B * h->P * 4, cudaMemcpyHostToDevice, h->copy_stream));
  if (bits_host) DCCN_CUDA_OK(cudaMemcpyAsync(S.d_bits, bits_host, (size_t)B * nb, cudaMemcpyHostToDevice, h->copy_stream));
  DCCN_CUDA_OK(cudaEventRecord(S.copied, h->copy_stream));
  // the pass, on the caller's stream, after the copy
  DCCN_CUDA_OK(cudaStreamWaitEvent(s, S.copied, 0));
  DCCN_CUDA_OK(cudaMemsetAsync(S.d_conf, 0, 4 * sizeof(int64_t), s));
  DCCN_CUDA_OK(cudaMemsetAsync(S.d_ce, 0, sizeof(double), s));
  int rc = dccn_forward(h, S.d_x, B, bits_host ? S.d_bits : nullptr, nullptr, hard_host ? S.d_hard : nullptr, nullptr,
                        nullptr, S.d_conf, S.d_ce, 0, s);
  if (rc) return rc;
  // results leave on their own stream: the D2H copy of this batch's decisions (the other copy engine) overlaps the
  // next batch's pass on the caller's stream and its H2D copy on the copy stream
  DCCN_CUDA_OK(cudaEventRecord(S.computed, s));
  DCCN_CUDA_OK(cudaStreamWaitEvent(h->d2h_stream, S.computed, 0));
  if (hard_host) DCCN_CUDA_OK(cudaMemcpyAsync(hard_host, S.d_hard, (size_t)B * nb, cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaMemcpyAsync(S.h_conf, S.d_conf, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaMemcpyAsync(S.h_ce, S.d_ce, sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaEventRecord(S.done, h->d2h_stream));
  S.busy = true;
  S.B = B;
  return 0;
}

// Same as dccn_forward_host_begin with the labels packed 8 per byte.
int dccn_forward_host_begin_packed(dccn_handle* h, int slot, const float* x_host, int64_t B,
                                   const uint8_t* bits_packed_host, uint8_t* hard_host, void* stream) {
  DCCN_CHECK(h && x_host && bits_packed_host && B > 0 && (slot == 0 || slot == 1), "bad argument");
  dccn_handle::HostSlot& S = h->slot[slot];
  DCCN_CHECK(!S.busy, "slot %d still in flight: call dccn_forward_host_end first", slot);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nb = (size_t)h->D * h->NB;
  DCCN_CHECK(((size_t)B * nb) % 8 == 0, "packed labels need B * n_data * nbits to be a multiple of 8");
  const size_t n_bytes = (size_t)B * nb / 8;
  if (S.frames < B) {
    int rc = dev_alloc(h, (void**)&S.d_x, (size_t)B * h->P * 4);
    rc |= dev_alloc(h, (void**)&S.d_bits, (size_t)B * nb);
    rc |= dev_alloc(h, (void**)&S.d_hard, (size_t)B * nb);
    if (rc) return rc;
    S.frames = B;
  }
  if (S.pack_frames < B) {
    int rc = dev_alloc(h, (void**)&S.d_pack, n_bytes);
    if (rc) return rc;
    S.pack_frames = B;
  }
  DCCN_CUDA_OK(cudaMemcpyAsync(S.d_x, x_host, (size_t)B * h->P * 4, cudaMemcpyHostToDevice, h->copy_stream));
  DCCN_CUDA_OK(cudaMemcpyAsync(S.d_pack, bits_packed_host, n_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  DCCN_CUDA_OK(cudaEventRecord(S.copied, h->copy_stream));
  DCCN_CUDA_OK(cudaStreamWaitEvent(s, S.copied, 0));
  {
    g_launches += 1;
    long long blocks = ((long long)n_bytes + 255) / 256;
    const long long cap = (long long)h->num_sms * 8;
    if (blocks > cap) blocks = cap;
    unpack_bits_kernel<<<(unsigned)blocks, 256, 0, s>>>(S.d_pack, (long long)n_bytes, S.d_bits);
    DCCN_CUDA_OK(cudaGetLastError());
  }
  DCCN_CUDA_OK(cudaMemsetAsync(S.d_conf, 0, 4 * sizeof(int64_t), s));
  DCCN_CUDA_OK(cudaMemsetAsync(S.d_ce, 0, sizeof(double), s));
  int rc = dccn_forward(h, S.d_x, B, S.d_bits, nullptr, hard_host ? S.d_hard : nullptr, nullptr, nullptr, S.d_conf, S.d_ce,
                        0, s);
  if (rc) return rc;
  DCCN_CUDA_OK(cudaEventRecord(S.computed, s));
  DCCN_CUDA_OK(cudaStreamWaitEvent(h->d2h_stream, S.computed, 0));
  if (hard_host) DCCN_CUDA_OK(cudaMemcpyAsync(hard_host, S.d_hard, (size_t)B * nb, cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaMemcpyAsync(S.h_conf, S.d_conf, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaMemcpyAsync(S.h_ce, S.d_ce, sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
  DCCN_CUDA_OK(cudaEventRecord(S.done, h->d2h_stream));
  S.busy = true;
  S.B = B;
  return 0;
}

int dccn_forward_host_end(dccn_handle* h, int slot, int64_t* conf_host, double* ce_sum_host) {
  DCCN_CHECK(h && (slot == 0 || slot == 1), "bad argument");
  dccn_handle::HostSlot& S = h->slot[slot];
  DCCN_CHECK(S.busy, "slot %d has no batch in flight", slot);
  DCCN_CUDA_OK(cudaEventSynchronize(S.done));
  S.busy = false;
  if (conf_host) memcpy(conf_host, S.h_conf, 4 * sizeof(int64_t));
  if (ce_sum_host) *ce_sum_host = *S.h_ce;
  return 0;
}

int dccn_forward_host(dccn_handle* h, const float* x_host, int64_t B, const uint8_t* bits_host, uint8_t* hard_host,
                      int64_t* conf_host, double* ce_sum_host, void* stream) {
  int rc = dccn_forward_host_begin(h, 0, x_host, B, bits_host, hard_host, stream);
  if (rc) return rc;
  return dccn_forward_host_end(h, 0, conf_host, ce_sum_host);
}

int dccn_cconv2d(const float* x_dev, int64_t B, int L, int W, int C, const float* kernel_dev, const float* bias_dev,
                 int filters, int kl, int kw, int padding, float* y_dev, void* stream) {
  DCCN_CHECK(x_dev && kernel_dev && bias_dev && y_dev, "null argument");
  DCCN_CHECK(B > 0 && L > 0 && W > 0 && C > 0 && filters > 0 && kl > 0 && kw > 0, "bad shape");
  DCCN_CHECK(padding == 0 || padding == 1, "padding must be 0 ('valid') or 1 ('same')");
  int Lo, Wo, pl = 0, pw = 0;
  if (padding == 1) {
    Lo = L; Wo = W; pl = (kl - 1) / 2; pw = (kw - 1) / 2;
  } else {
    Lo = L - kl + 1; Wo = W - kw + 1;
    DCCN_CHECK(Lo > 0 && Wo > 0, "'valid' kernel larger than the input");
  }
  const long long total = (long long)B * Lo * Wo * filters;
  cconv2d_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const float2*)x_dev, (long long)B, L, W, C, kernel_dev, bias_dev, filters, kl, kw, pl, pw, Lo, Wo,
      (float2*)y_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

int dccn_vconv2d(const float* x_dev, int64_t B, int L, int W, int C, const float* kernel_dev, const float* bias_dev,
                 int filters, int kl, int kw, int padding, float* y_dev, void* stream) {
  DCCN_CHECK(x_dev && kernel_dev && bias_dev && y_dev, "null argument");
  DCCN_CHECK(B > 0 && L > 0 && W > 0 && C > 0 && filters > 0 && kl > 0 && kw > 0, "bad shape");
  DCCN_CHECK(padding == 0 || padding == 1, "padding must be 0 ('valid') or 1 ('same')");
  int Lo, Wo, pl = 0, pw = 0;
  if (padding == 1) {
    Lo = L; Wo = W; pl = (kl - 1) / 2; pw = (kw - 1) / 2;
  } else {
    Lo = L - kl + 1; Wo = W - kw + 1;
    DCCN_CHECK(Lo > 0 && Wo > 0, "'valid' kernel larger than the input");
  }
  const long long total = (long long)B * Lo * Wo * filters;
  g_launches += 1;
  vconv2d_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const float2*)x_dev, (long long)B, L, W, C, kernel_dev, bias_dev, filters, kl, kw, pl, pw, Lo, Wo,
      (float2*)y_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

int dccn_chan_fir_awgn(dccn_handle* h, const float* tx_dev, int64_t B, int n_samp, const double* alpha_dev,
                       const double* coeff_dev, int n_taps, int n_fir, const double* z_dev, const float* snr_db_dev,
                       const double* normals_dev, uint64_t seed, float* rx_dev, float* fir_only_dev, void* stream) {
  DCCN_CHECK(h && tx_dev && rx_dev && snr_db_dev && B > 0 && n_samp > 0, "bad argument");
  DCCN_CHECK(n_taps >= 0 && n_taps <= 32 && n_fir <= kMaxFir, "at most 32 paths / FIR taps");
  DCCN_CHECK(n_taps == 0 || (coeff_dev != nullptr && n_fir >= 1), "coeff_dev / n_fir missing");
  cudaStream_t s = (cudaStream_t)stream;
  DCCN_CUDA_OK(cudaMemsetAsync(h->d_power, 0, sizeof(double), s));
  float* faded = fir_only_dev ? fir_only_dev : rx_dev;
  {
    LaunchScope ls(h, SLOT_CHAN_FIR, s);
    chan_fir_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>((const float2*)tx_dev, (long long)B, n_samp, alpha_dev,
                                                           coeff_dev, n_taps, n_fir, z_dev, seed, 0, 1, (float2*)faded,
                                                           h->d_power);
  }
  LaunchScope ls2(h, SLOT_AWGN, s);
  long long blocks = (B + 7) / 8;                    // one warp per frame, grid-stride
  if (blocks > (long long)h->num_sms * 8) blocks = (long long)h->num_sms * 8;
  awgn_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float2*)faded, (long long)B, n_samp, h->d_power, snr_db_dev,
                                               normals_dev, seed ^ 0x9E3779B97F4A7C15ull, (float2*)rx_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

/* fading only (no AWGN), for a strided subset of the frames: frame0, frame0+fstride, ... */
int dccn_chan_fading(dccn_handle* h, const float* tx_dev, int64_t B, int n_sym, int n_sc, const double* alpha_dev,
                     const double* coeff_dev, int n_taps, int n_fir, double doppler_hz, double sample_rate,
                     const double* z_or_theta_dev, uint64_t seed, int64_t frame0, int64_t fstride, int reset_power,
                     float* faded_dev, void* stream) {
  DCCN_CHECK(h && tx_dev && faded_dev && B > 0 && n_sym > 0 && n_sc > 0 && fstride >= 1 && frame0 >= 0, "bad argument");
  DCCN_CHECK(n_taps >= 0 && n_taps <= kMaxPaths && n_fir >= 1 && n_fir <= kMaxFir, "at most %d paths / %d FIR taps", kMaxPaths, kMaxFir);
  DCCN_CHECK(n_taps == 0 || coeff_dev != nullptr, "coeff_dev missing");
  DCCN_CHECK(doppler_hz <= 0.0 || n_taps > 0, "Doppler fading needs a tap profile");
  cudaStream_t s = (cudaStream_t)stream;
  if (reset_power) DCCN_CUDA_OK(cudaMemsetAsync(h->d_power, 0, sizeof(double), s));
  if (frame0 >= B) return 0;
  const long long nf = (B - frame0 + fstride - 1) / fstride;
  LaunchScope ls(h, SLOT_CHAN_FIR, s);
  if (doppler_hz > 0.0)
    chan_doppler_kernel<<<(unsigned)((nf + 7) / 8), 256, 0, s>>>((const float2*)tx_dev, (long long)B, n_sym, n_sc,
                                                               alpha_dev, coeff_dev, n_taps, n_fir, doppler_hz,
                                                               (double)n_sc / sample_rate, z_or_theta_dev, seed,
                                                               frame0, fstride, (float2*)faded_dev, h->d_power);
  else
    chan_fir_kernel<<<(unsigned)((nf + 7) / 8), 256, 0, s>>>((const float2*)tx_dev, (long long)B, n_sym * n_sc, alpha_dev,
                                                           coeff_dev, n_taps, n_fir, z_or_theta_dev, seed, frame0,
                                                           fstride, (float2*)faded_dev, h->d_power);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

/* AWGN_channel_np on an already faded batch; uses the power accumulated by dccn_chan_fading */
int dccn_chan_awgn(dccn_handle* h, const float* faded_dev, int64_t B, int n_samp, const float* snr_db_dev,
                   const double* normals_dev, uint64_t seed, float* rx_dev, void* stream) {
  DCCN_CHECK(h && faded_dev && rx_dev && snr_db_dev && B > 0 && n_samp > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  LaunchScope ls(h, SLOT_AWGN, s);
  long long blocks = (B + 7) / 8;                    // one warp per frame, grid-stride
  if (blocks > (long long)h->num_sms * 8) blocks = (long long)h->num_sms * 8;
  awgn_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float2*)faded_dev, (long long)B, n_samp, h->d_power, snr_db_dev,
                                               normals_dev, seed ^ 0x9E3779B97F4A7C15ull, (float2*)rx_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

int dccn_ber_accum(const uint8_t* hard_dev, const uint8_t* bits_dev, int64_t n, int64_t* conf_dev, void* stream) {
  DCCN_CHECK(hard_dev && bits_dev && conf_dev && n >= 0, "bad argument");
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ber_accum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hard_dev, bits_dev, (long long)n,
                                                                      (unsigned long long*)conf_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

int dccn_tx_frames(dccn_handle* h, const uint8_t* bits_dev, int64_t B, const int32_t* data_sc_dev, int n_data,
                   const int32_t* pilot_sc_dev, int n_pilot, const float* constellation_dev, float pilot_re,
                   float pilot_im, float* tx_dev, void* stream) {
  DCCN_CHECK(h && bits_dev && data_sc_dev && constellation_dev && tx_dev && B > 0, "bad argument");
  DCCN_CHECK(n_data == h->D, "n_data (%d) != cfg.n_data (%d)", n_data, h->D);
  cudaStream_t s = (cudaStream_t)stream;
  const int S = h->S, K = h->K;
  DCCN_CHECK(n_pilot >= 0 && (n_pilot == 0 || pilot_sc_dev), "pilot_sc_dev missing");
  // subcarrier role map (-1 guard, -2 pilot, >= 0 data index), rebuilt on the device from the caller's index arrays by
  // one tiny kernel: the call stays fully asynchronous (no D2H, no allocation, no stream synchronisation) and nothing
  // is cached across calls, so the index arrays may change between calls
  g_launches += 2;
  txmap_kernel<<<1, 512, 0, s>>>(data_sc_dev, n_data, pilot_sc_dev, n_pilot, S * K, h->d_txmap);
  const long long syms = (long long)B * S;
  if (h->tx_v2 && K == 64 && h->cfg.cp_len > 0 && h->cfg.cp_len < 64) {
    long long blocks = (syms + 7) / 8;
    const long long cap = (long long)h->num_sms * 8;
    if (blocks > cap) blocks = cap;
    tx64_kernel<<<(unsigned)blocks, 256, 0, s>>>(bits_dev, (long long)B, S, h->cfg.cp_len, h->NB, h->D, h->d_txmap,
                                                 (const float2*)constellation_dev, make_float2(pilot_re, pilot_im),
                                                 (float2*)tx_dev);
  } else {
    const size_t smem = (size_t)9 * K * sizeof(double2);
    tx_kernel<<<(unsigned)((syms + 7) / 8), 256, smem, s>>>(bits_dev, (long long)B, S, K, h->cfg.cp_len, h->NB, h->D,
                                                            h->d_txmap, (const float2*)constellation_dev,
                                                            make_float2(pilot_re, pilot_im), (float2*)tx_dev);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

/* bits -> OFDM frames -> static Rayleigh FIR in one kernel (nfft = 64): the transmitted frame stays in shared memory */
int dccn_tx_fade(dccn_handle* h, const uint8_t* bits_dev, int64_t B, const int32_t* data_sc_dev, int n_data,
                 const int32_t* pilot_sc_dev, int n_pilot, const float* constellation_dev, float pilot_re, float pilot_im,
                 const double* alpha_dev, const double* coeff_dev, int n_taps, int n_fir, const double* z_dev,
                 uint64_t seed, int reset_power, float* tx_dev, float* faded_dev, void* stream) {
  DCCN_CHECK(h && bits_dev && data_sc_dev && constellation_dev && faded_dev && B > 0, "bad argument");
  DCCN_CHECK(n_data == h->D, "n_data (%d) != cfg.n_data (%d)", n_data, h->D);
  DCCN_CHECK(n_pilot >= 0 && (n_pilot == 0 || pilot_sc_dev), "pilot_sc_dev missing");
  DCCN_CHECK(h->K == 64 && h->cfg.cp_len > 0 && h->cfg.cp_len <= 16 && h->S <= 8,
             "dccn_tx_fade is the nfft = 64 feeder (cp <= 16, <= 8 symbols); use dccn_tx_frames + dccn_chan_fading");
  DCCN_CHECK(n_taps >= 0 && n_taps <= 32 && n_fir >= 1 && n_fir <= kMaxFir, "at most 32 paths / FIR taps");
  DCCN_CHECK(n_taps == 0 || coeff_dev != nullptr, "coeff_dev missing");
  cudaStream_t s = (cudaStream_t)stream;
  if (reset_power) DCCN_CUDA_OK(cudaMemsetAsync(h->d_power, 0, sizeof(double), s));
  g_launches += 1;
  txmap_kernel<<<1, 512, 0, s>>>(data_sc_dev, n_data, pilot_sc_dev, n_pilot, h->S * h->K, h->d_txmap);
  LaunchScope ls(h, SLOT_CHAN_FIR, s);
  long long blocks = (B + kGenWarps - 1) / kGenWarps;
  const long long cap = (long long)h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int n_samp = h->S * h->T;
  auto launch = [&](auto kern, size_t smem) -> int {
    static bool attr_set[2] = {false, false};
    const int which = smem == tx_fade_smem<18>() ? 0 : 1;
    if (!attr_set[which]) {
      DCCN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[which] = true;
    }
    kern<<<(unsigned)blocks, 32 * kGenWarps, smem, s>>>(bits_dev, (long long)B, h->S, h->cfg.cp_len, h->NB, h->D, h->d_txmap,
                                                       (const float2*)constellation_dev, make_float2(pilot_re, pilot_im),
                                                       alpha_dev, coeff_dev, n_taps, n_fir, z_dev, seed, (float2*)tx_dev,
                                                       (float2*)faded_dev, h->d_power);
    return 0;
  };
  // L = output samples per lane of the FIR's blocked mapping: 18 covers the 560-sample LTE frame, 20 the 640-sample v1 frame
  int rc = n_samp <= 32 * 18 ? launch(tx_fade_kernel<18>, tx_fade_smem<18>()) : launch(tx_fade_kernel<20>, tx_fade_smem<20>());
  if (rc) return rc;
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

#ifdef DCCN_TRACE
/* debug build only (tools/trace_gemm.py): buffer of 8 x 4096 int64 clock samples written by CTA 0 of the next GEMMs */
int dccn_debug_trace(long long* buf_dev) {
  dccn::g_trace_host_ptr = buf_dev;
  return 0;
}
/* debug build only (tools/ablate_gemm.py): switch pieces of the GEMM pipeline off to time the rest */
int dccn_debug_abl(int mask) {
  dccn::g_abl_host = mask;
  return 0;
}
#endif

/* measurement aid (tools/chain_trace.py, not part of include/dccn.h): CTA 0 of the next chained kernels (which = 0 front,
   1 tail) writes 5 x 1024 clock64() samples into buf_dev; nullptr switches it off */
int dccn_debug_chain_trace(int which, long long* buf_dev) {
  if (which < 0 || which > 1) return -2;
  g_chain_trace[which] = buf_dev;
  return 0;
}

int64_t dccn_launch_count(void) { return (int64_t)g_launches.load(); }

int dccn_profile_enable(dccn_handle* h, int on) {
  DCCN_CHECK(h, "null handle");
  h->prof = on != 0;
  return 0;
}

int dccn_profile_collect(dccn_handle* h, double* ms_out, int64_t* count_out, int max_slots) {
  DCCN_CHECK(h && ms_out && count_out && max_slots >= SLOT_COUNT, "need room for %d slots", (int)SLOT_COUNT);
  DCCN_CUDA_OK(cudaDeviceSynchronize());
  for (int i = 0; i < max_slots; ++i) {
    ms_out[i] = 0.0;
    count_out[i] = 0;
  }
  for (auto& r : h->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_out[r.slot] += ms;
      count_out[r.slot] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof_recs.clear();
  return SLOT_COUNT;
}

const char* dccn_profile_slot_name(int slot) { return (slot >= 0 && slot < SLOT_COUNT) ? kSlotNames[slot] : ""; }

int dccn_debug_tma_rate(const float* mat_dev, int rows, int cols, int ld, int stages, int boxes, int iters,
                        int grid, long long* clks_dev) {
  DCCN_CHECK(mat_dev && clks_dev && rows % 128 == 0 && cols % 32 == 0 && stages >= 1 && boxes >= 1, "bad argument");
  CUtensorMap tm;
  int rc = make_tmap(&tm, mat_dev, rows, cols, ld, 128);
  if (rc) return rc;
  const int smem = stages * boxes * 16384 + 1024 + 256;
  DCCN_CHECK(smem <= 227 * 1024, "too much shared memory");
  DCCN_CUDA_OK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_rate_kernel<<<grid, 64, smem>>>(tm, rows, cols, stages, boxes, iters, clks_dev);
  DCCN_CUDA_OK(cudaGetLastError());
  DCCN_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

int dccn_debug_mma_rate(int bn, int n_mma, int per_commit, int mode, int dep, int grid, long long* clks_dev) {
  DCCN_CHECK(clks_dev && (bn == 128 || bn == 256), "bn must be 128 or 256");
  const int smem = (mode >= 2 ? 131072 : 16384 + bn * 128) + 1024 + 64;
  DCCN_CHECK(mode < 2 || (bn == 128 && per_commit > 0 && per_commit % 3 == 0), "mode 2/3: bn 128, per_commit = 3 * k-steps");
  if (bn == 128) {
    DCCN_CUDA_OK(cudaFuncSetAttribute(mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<128><<<grid, 64, smem>>>(n_mma, per_commit, mode, dep, clks_dev);
  } else {
    DCCN_CUDA_OK(cudaFuncSetAttribute(mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<256><<<grid, 64, smem>>>(n_mma, per_commit, mode, dep, clks_dev);
  }
  DCCN_CUDA_OK(cudaGetLastError());
  DCCN_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

uint32_t dccn_crc32c(const void* data_host, size_t n, uint32_t crc) {
  // slicing-by-8 over tables built on first use (reflected polynomial 0x82F63B78)
  static uint32_t tbl[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tbl[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tbl[t][i] = (tbl[t - 1][i] >> 8) ^ tbl[0][tbl[t - 1][i] & 0xFFu];
    ready = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data_host);
  crc = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = tbl[7][lo & 0xFFu] ^ tbl[6][(lo >> 8) & 0xFFu] ^ tbl[5][(lo >> 16) & 0xFFu] ^ tbl[4][lo >> 24] ^
          tbl[3][hi & 0xFFu] ^ tbl[2][(hi >> 8) & 0xFFu] ^ tbl[1][(hi >> 16) & 0xFFu] ^ tbl[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = tbl[0][(crc ^ *p++) & 0xFFu] ^ (crc >> 8);
  return ~crc;
}

int dccn_bit_source(uint8_t* bits_dev, int64_t n, uint64_t seed, void* stream) {
  DCCN_CHECK(bits_dev && n >= 0, "bad argument");
  if (n == 0) return 0;
  const long long threads = (n + 127) / 128;
  g_launches += 1;
  bit_source_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits_dev, (long long)n, seed);
  DCCN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
