// handle.cuh -- the library handle and the helpers shared by the translation units of libdccn.so
// (dccn.cu: inference path + C ABI; train.cu: the equalizer transfer-learning step).
#pragma once
#include "../../include/dccn.h"

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "epilogue.cuh"

namespace dccn {

// ---------------------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------------------
struct Act {          // activation matrix, 1 or 2 planes
  float* p0 = nullptr;
  float* p1 = nullptr;
  int ld = 0;
  unsigned* amax = nullptr;   // device word: max |v| written by the producer(s) of this buffer in the current pass
                              // (bit pattern of a float >= 0); feeds the operand scale of the fp16 hi/lo GEMM form
};

struct GemmLayer {
  int K = 0, N = 0;
  int BN = 128;                 // tcgen05 tile width
  std::vector<float> W;         // host [K, N]
  std::vector<float> bias;      // host [N]
  float* dW = nullptr;          // [K, N] fp32 (SIMT path)
  float* dWt0 = nullptr;        // [N, K] tf32-hi (or full fp32 for FAST)  -- B operand, K-major
  float* dWt1 = nullptr;        // [N, K] tf32-lo (PARITY only)
  float* dBias = nullptr;
  CUtensorMap tmB0, tmB1;
  // DCCN_F16X3 (staged): [N, K] fp16 hi / lo planes of W * w_scale (w_scale = a power of two that puts max|W| in
  // [2^13, 2^14), so that the lo plane stays a normal fp16), their [BN x 64] tensor maps, and 1 / w_scale
  void* dWh0 = nullptr;
  void* dWh1 = nullptr;
  CUtensorMap tmH0, tmH1;
  float w_scale_inv = 1.f;
  bool f16_ok = false;
  bool built = false;
  bool fused = false;           // consumed by a fused (phase-eq / demod-head) epilogue
  bool mc = false;              // run as cta_group::2 CTA pairs (each CTA holds half of the weight tile)
};

// wiring of an equalizer graph (--opt), see pack_layers_host / run_chunk
struct EqSpec {
  int front2_cconv = 1;       // 1: (1,K) 'valid' complex conv, 0: dense 2K -> 2K
  int n_chain = 3;            // frame-level dense layers after the pilot bottleneck
  int chain_act[4] = {0, 0, 1, 0};   // 0 linear, 1 tanh
  int toeplitz = 1;           // (S,K) 'same' complex conv behind the chain
  int tail = 0;               // 0: cconv(eq) | cconv(corr) -> dense, 1: dense(2K) -> dense(2T), 2: tf.ifft -> dense(2T)
  int vector = 0;             // 1: the complex convs are layers_conv2d_vector (complex.py:199-255), --opt 7
  int generic = 0;            // 1: run through the generic wiring of run_chunk (opts 1-5); 0: equalizer_ofdm's own path
};

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

extern std::atomic<long long> g_launches;   // kernels launched by this library (dccn_launch_count)
enum { SLOT_MOMENTS = 0, SLOT_PREP, SLOT_G1, SLOT_G2, SLOT_G3, SLOT_G4, SLOT_G5, SLOT_G6, SLOT_G7_PHASEEQ,
       SLOT_G8, SLOT_G9, SLOT_G10, SLOT_R1, SLOT_R2_HEAD, SLOT_CHAN_FIR, SLOT_AWGN, SLOT_R2_GEMM, SLOT_T_HEAD, SLOT_T_DGRAD, SLOT_T_WGRAD, SLOT_T_POINT, SLOT_T_ADAM, SLOT_T_REPACK, SLOT_F1, SLOT_F4, SLOT_F9, SLOT_CH_FRONT, SLOT_CH_TAIL, SLOT_COUNT };
extern const char* kSlotNames[SLOT_COUNT];
struct ProfRec { int slot; cudaEvent_t a, b; };
struct TrainState;
}  // namespace dccn

struct dccn_handle {
  dccn_cfg cfg;
  bool prof = false;
  std::vector<dccn::ProfRec> prof_recs;
  int device = 0;
  int num_sms = 148;
  std::map<std::string, dccn::HostTensor> raw;
  bool committed = false;
  // geometry
  int S, K, T, Tin, F, D, NB, P;      // T = samples/symbol incl. CP, Tin = samples the receiver consumes
  int chunk;           // most frames one internal pass may take (cfg.chunk_frames, default 65536)
  int64_t ws_frames = 0;   // frames the inter-layer buffers are currently sized for (grown on demand, ensure_workspace)
  std::vector<void*> ws_allocs;
  bool ws_eqc = false, ws_train = false;   // optional buffers: folded schedule / training forward
  bool ws_dirty = false;                   // an optional buffer was requested after the last build
  size_t ws_act_bytes = 0;
  int kc = 1;          // k-blocks accumulated inside TMEM before the fp32 register add (parity mode)
  int small_first = 1; // parity GEMMs: per k-block the 8 cross-term MMAs first, then the 4 hi*hi MMAs (DCCN_SMALL_FIRST=0: interleaved)
  int epi_prefetch = 1; // phase-equaliser epilogue: L2-prefetch the tile's f rows at tile start (DCCN_EPI_PREFETCH=0 disables)
  int head_subs = 2;   // data subcarriers per thread of the demod head kernel (1 or 2)
  int head_blocks = 0; // resident head blocks per SM (0 = 2048 / threads)
  int bn_wide = 0;     // use 256-wide tiles for the 896-wide layers
  int a_tmem = 1;      // parity mode: A operand hi/lo staged in TMEM (TS-form MMA) instead of shared memory
  int multicast = 0;   // cta_group::2 CTA pairs (DCCN_PAIR=1 enables; measured slower than single-CTA tiles)
  int mc_min_k = 128;
  int fused_head = 0;  // 1: demod head inside the GEMM epilogue; 0: separate full-occupancy kernel
  int bn192 = 0;       // 1: 192-wide tiles for 128 < N <= 192 (2-stage smem-split form); 0: two 128-wide A-in-TMEM tiles
  int band_skip = 1;   // skip the structurally-zero k-blocks of the Toeplitz ((S,K) 'same' conv) operand
  int tx_v2 = 1;       // 8 x 8 IDFT transmitter kernel for nfft = 64 (DCCN_TX_V2=0: the generic K-point DFT kernel)
  int32_t* d_txmap = nullptr;           // [S*K] subcarrier role map, rebuilt on the device by every dccn_tx_frames call
  int chain = 1;       // per-symbol layer runs of equalizer_ofdm as chained kernels (chain.cu; DCCN_CHAIN=0: layer by layer through HBM)
  int f16x3 = 1;       // inference GEMMs of the parity mode through the fp16 hi/lo kind::f16 form (DCCN_F16X3=0: tf32 pairs)
  // monitor outputs requested for the NEXT forward (dccn_forward_monitors), consumed and cleared by it
  float* mon_snr_db = nullptr;          // [B] equalizer snr_db (model.py:464-475)
  const int32_t* mon_pilot_carriers = nullptr;
  int mon_n_pilot = 0;
  int64_t mon_frame0 = 0;               // first frame of the pass being run (offset into mon_snr_db)
  unsigned* d_amax = nullptr;           // [kAmaxSlots] per-buffer max |activation| of the current pass (Act::amax)
  // layers
  dccn::GemmLayer r1, r2;                               // receiver: learned DFT, demod dense
  dccn::GemmLayer g1, g2, g3, g4, g5, g6, g7, g8, g9, g10;   // equalizer
  dccn::GemmLayer gx0;              // 4th chain layer of the ablation graphs (--opt 3, 5)
  int eq_opt = 0;                   // cfg.eq_opt; != 0: the roles of g1..g10 follow `eqs` (pack_layers_host)
  dccn::EqSpec eqs;
  dccn::HeadWeights hw;
  // DCCN_FWD_FOLDED: consecutive linear layers pre-multiplied (built on first use, from the committed weights)
  dccn::GemmLayer f1, f4, f9;       // dense.conv3d | dense_2.dense_3.dense_4(tanh) | (conv3d_3,conv3d_2).dense_5.fft_like
  bool fold_built = false;
  int default_flags = 0;            // OR-ed into the flags of every forward (DCCN_FOLD=1 sets DCCN_FWD_FOLDED)
  dccn::Act eqc;                    // folded schedule: per symbol [eq (2K) | corr (K)]
  // workspace
  std::vector<void*> allocs;
  double* d_sums = nullptr;       // [2P] moments accumulators
  float* d_mean = nullptr;
  float* d_rstd = nullptr;
  double* d_power = nullptr;      // channel power accumulator
  unsigned long long* d_conf = nullptr;   // internal [4]
  double* d_ce = nullptr;
  dccn::Act a0, t1, f, p32, u1, u2, eq, corr, cat, oeq, r1o, out_iq;
  // training (train.cu): state, and the two extra forward buffers a backward pass needs
  dccn::TrainState* tr = nullptr;
  bool train_fwd = false;         // forward pass keeps every activation (tanh output -> u3, chest -> chest_buf)
  dccn::Act u3;
  float* chest_buf = nullptr;
  // staging for the host-buffer entry points: two slots so that the H2D copy of one batch
  // overlaps the pass over the previous one (copy stream + the caller's compute stream)
  struct HostSlot {
    float* d_x = nullptr;
    uint8_t* d_bits = nullptr;
    uint8_t* d_hard = nullptr;
    uint8_t* d_pack = nullptr;     // labels packed 8 per byte (dccn_forward_host_begin_packed)
    int64_t pack_frames = 0;
    int64_t frames = 0;            // capacity
    int64_t B = 0;                 // batch in flight
    int64_t* d_conf = nullptr;     // device results
    double* d_ce = nullptr;
    int64_t* h_conf = nullptr;     // pinned host results
    double* h_ce = nullptr;
    uint8_t* hard_host = nullptr;  // caller's destination for hard bits (may be null)
    cudaEvent_t copied = nullptr, computed = nullptr, done = nullptr;
    bool busy = false;
  } slot[2];
  cudaStream_t copy_stream = nullptr;      // H2D of the host-buffer entry points
  cudaStream_t d2h_stream = nullptr;       // their D2H (decisions, confusion matrix, loss)
  size_t ws_bytes = 0;
};

namespace dccn {

// counts launches and (when profiling is on) brackets them with CUDA events on the launch stream
struct LaunchScope {
  dccn_handle* h;
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  int slot;
  LaunchScope(dccn_handle* h_, int slot_, cudaStream_t s_, int n_kernels = 1) : h(h_), s(s_), slot(slot_) {
    g_launches += n_kernels;
    if (h && h->prof && h->prof_recs.size() < 65536) {
      cudaEvent_t a;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
      h->prof_recs.push_back(ProfRec{slot, a, b});
    }
  }
  ~LaunchScope() {
    if (b) cudaEventRecord(b, s);
  }
};


// ---- shared between dccn.cu and train.cu ------------------------------------------------------
int make_tmap(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
int dev_alloc(dccn_handle* h, void** p, size_t bytes);
int alloc_act(dccn_handle* h, Act* a, int64_t rows, int ld, bool split);
int ensure_workspace(dccn_handle* h, int64_t frames);            // inter-layer buffers for passes of up to `frames` frames
const HostTensor* find(const dccn_handle* h, const std::string& n);
int pack_layers_host(dccn_handle* h);                            // h->raw -> GemmLayer::W / bias (host only)
int upload_layer(dccn_handle* h, GemmLayer* L, cudaStream_t s);  // GemmLayer::W -> device operands (+ TMA maps)
int run_moments(dccn_handle* h, const float* x, int64_t B, float* mean, float* rstd, cudaStream_t s);
int run_chunk(dccn_handle* h, const float* x, int64_t Bc, const uint8_t* bits, float* soft, uint8_t* hard,
              float* eq_out, float* chest_out, unsigned long long* conf, double* ce, int flags, cudaStream_t s);
int run_gemm_store(dccn_handle* h, int slot, const GemmLayer& L, const Act& A, int a_col_off, int64_t M,
                   const EpiStore& epi, cudaStream_t s, int ksplit = 1);
void conf_accumulate(const unsigned long long* src, int64_t* dst, cudaStream_t s);
// train.cu hooks used by the C ABI in dccn.cu
void train_free(dccn_handle* h);
int train_on_commit(dccn_handle* h, cudaStream_t s);
int train_fetch_weight(dccn_handle* h, const char* tf_name, HostTensor* t);

inline ActOut out_of(const Act& a, int col_off = 0) { return ActOut{a.p0, a.p1, a.ld, col_off}; }

constexpr int kAmaxSlots = 32;

inline EpiStore store_epi(const GemmLayer& L, const Act& dst, int col_off, int64_t M, int act = 0, float* aux = nullptr,
                          int aux_ld = 0) {
  EpiStore e;
  e.amax = dst.amax;
  e.bias = L.dBias;
  e.out = out_of(dst, col_off);
  e.aux = aux;
  e.aux_ld = aux_ld;
  e.act = act;
  e.M = (int)M;
  e.N = L.N;
  return e;
}

}  // namespace dccn
