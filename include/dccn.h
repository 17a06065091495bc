/*
 * dccn.h -- C ABI of libdccn.so: the B200-native (sm_100a) implementation of the
 * DCCN OFDM receiver hot path of zhongyuanzhao/dl_ofdm.
 *
 * The reference has no FFI: its boundary is Python calling a TF-1 graph
 * (`session.run(fetches, feed)` at dev/py/ofdmreceiver_np.py:80,234,256 and
 * dev/py/ofdmreceiver_np_mp.py:89,419,445).  Each entry point below names the
 * reference interface it replaces.  The Python host (dl_ofdm_b200/) binds this
 * header with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a device pointer owned by
 *     the caller (e.g. torch tensor storage), *_host is host memory;
 *   - every call is asynchronous on the caller-supplied stream (a cudaStream_t
 *     passed as void*; NULL = legacy default stream) unless it says otherwise;
 *   - return value 0 = success, negative = error (dccn_last_error() gives the
 *     thread-local message).  The library never falls back to a CPU path.
 *   - tensors use the reference layouts: IQ is the fastest axis everywhere.
 */
#ifndef DCCN_H_
#define DCCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCCN_ABI_VERSION 3

/* arithmetic of the GEMM layers */
enum {
  DCCN_PREC_EXACT  = 0, /* fp32 FFMA on CUDA cores (bit-near the reference's fp32)        */
  DCCN_PREC_PARITY = 1, /* tcgen05 kind::tf32, 3-pass hi/lo split, fp32 accumulate in TMEM */
  DCCN_PREC_FAST   = 2  /* tcgen05 kind::tf32, single pass (reduced precision; not parity) */
};
enum { DCCN_HEAD_DEV = 0, DCCN_HEAD_V1 = 1 };
/* dccn_forward flags: entry points of the sub-graphs the reference exposes as functions */
enum {
  DCCN_FWD_NO_NORM = 1, /* x is already the normalised 'input:0' tensor (skip a2)              */
  DCCN_FWD_EQ_ONLY = 2, /* stop after equalizer_ofdm (eq_dev / chest_dev are the outputs)      */
  DCCN_FWD_SKIP_EQ = 4, /* run ofdm_dense_rx directly on x even if the handle has an equalizer */
  DCCN_FWD_FOLDED = 8   /* inference-graph optimisation (opt-in): consecutive LINEAR layers of the reference graph are
                           pre-multiplied in fp64 when first used -- dense.conv3d, dense_2.dense_3.dense_4 (before its
                           tanh), (conv3d_3 | conv3d_2).dense_5.fft_like -- so 12 GEMMs become 5.  Same function of
                           the inputs (different fp32 rounding); ignored when eq_dev is requested, for
                           DCCN_FWD_EQ_ONLY, and once dccn_train_init has been called (weights then change). */
};

/* Geometry + model selection; field names follow the reference FLAGS
 * (dev/py/ofdmreceiver_np_mp.py:33-58) and ofdm_tx attributes (dev/py/ofdm.py:198-273). */
typedef struct dccn_cfg {
  int32_t nfft;        /* FLAGS.nfft      K, 64                                        */
  int32_t cp_len;      /* ofdm_tx.CP      16 (longcp) or 4                             */
  int32_t nsymbol;     /* FLAGS.nsymbol   7 (dev / 'lte' pilots) or 8 (v1)             */
  int32_t nfilter;     /* FLAGS.nfilter   F, 64                                        */
  int32_t nbits;       /* FLAGS.nbits     1..4                                         */
  int32_t use_cp;      /* FLAGS.cp        receiver consumes the cyclic prefix          */
  int32_t n_data;      /* ofdm_tx.frame_size  data subcarriers per frame (320 / 368)   */
  int32_t pilot_size;  /* ofdm_tx.pilot_size  16 for 'lte'                             */
  int32_t head;        /* DCCN_HEAD_*     demodulation head variant                    */
  int32_t equalizer;   /* 0 = basic receiver only, 1 = equalizer_ofdm (--opt=0) in front */
  int32_t precision;   /* DCCN_PREC_*                                                  */
  int32_t chunk_frames;/* most frames per internal pass (0 = 65536); buffers grow on demand */
  int32_t eq_opt;      /* FLAGS.opt       which equalizer graph when equalizer != 0 (ofdmreceiver_np_mp.py:292-311):
                          0 equalizer_ofdm, 1 equalizer_nocconv, 2 equalizer_noresdl, 3 equalizer_dnnE,
                          4 equalizer_noresdl2, 5 equalizer_noresdl4, 7 equalizer_separateIQ (dev/py/model.py:349-1218);
                          inference only for != 0 */
} dccn_cfg;

typedef struct dccn_handle dccn_handle;

/* -- lifecycle ------------------------------------------------------------- */
int         dccn_abi_version(void);
const char* dccn_last_error(void);
/* Replaces graph construction: ofdm_dense_rx(...) dev/py/model.py:1222 and
 * equalizer_ofdm(...) dev/py/model.py:349 (built at ofdmreceiver_np.py:144,
 * ofdmreceiver_np_mp.py:294).  Binds to the current CUDA device. */
int  dccn_create(const dccn_cfg* cfg, dccn_handle** out);
void dccn_destroy(dccn_handle* h);
size_t dccn_workspace_bytes(const dccn_handle* h);

/* -- weights ---------------------------------------------------------------
 * Replaces tf.train.Saver.restore (dev/py/model.py:51-56).  `tf_name` is the TF
 * variable name ("fft_like/conv3d/kernel", "demodulation/dense/bias",
 * "Equalizer/dense_4/kernel", ...), `host` the tensor in the reference layout
 * (conv3d [kl,kw,1,Cin,2F], dense [in,out], conv2d [1,1,Cin,Cout]).  The library
 * repacks: dead-tap removal of the 'same' (1,K) layer, [[a,b],[-b,-a]] tiling of
 * complex layers with the reference sign (dev/py/complex.py:187-188), Toeplitz
 * expansion of the (S,K) 'same' conv, tf32 hi/lo split.  Synchronous. */
int dccn_set_weight(dccn_handle* h, const char* tf_name, const float* host,
                    const int64_t* shape, int rank);
/* copies the stored (reference-layout) tensor back (the trained value once dccn_train_init was called);
 * returns element count or <0 */
int64_t dccn_get_weight(dccn_handle* h, const char* tf_name, float* host, int64_t capacity);
int dccn_commit_weights(dccn_handle* h, void* stream);

/* -- a2: tf.nn.moments(x,[0]) + batch_normalization (ofdmreceiver_np.py:128-129)
 * x_dev float32 [B,S,T,2]; mean_dev/rstd_dev float32 [S*T*2] (rstd = rsqrt(var+1e-9)). */
int dccn_batch_moments(dccn_handle* h, const float* x_dev, int64_t B,
                       float* mean_dev, float* rstd_dev, void* stream);

/* -- the receiver pass: replaces session.run([conf_matrix, linear_ber, ce_mean,
 * output, ...], {tx_ofdm: x, bits_in: y}) (ofdmreceiver_np.py:80, _mp.py:89).
 *   x_dev      float32 [B,S,T,2]   'tx_ofdm' feed
 *   bits_dev   uint8   [B,D,nbits] 'bits_in' feed (0/1), or NULL (no BER/loss)
 *   soft_dev   float32 [B,D,nbits,2] 'output' (softmax), or NULL
 *   hard_dev   uint8   [B,D,nbits]  argmax of 'output' (first index on ties), or NULL
 *   eq_dev     float32 [B,S,T,2]    equalizer output (cfg.equalizer only), or NULL
 *   chest_dev  float32 [B,S,K,2]    channel estimate 'chest' (cfg.equalizer only), or NULL
 *   conf_dev   int64   [2,2]        'conf_matrix' (rows = truth); ACCUMULATED into
 *   ce_sum_dev double  [1]          sum of the softmax-xent terms (ce_mean * count); accumulated
 * The batch-moment norm is part of the pass (moments over all B frames) unless
 * DCCN_FWD_NO_NORM; with it and DCCN_FWD_SKIP_EQ the call is ofdm_dense_rx(x)
 * (model.py:1222), with DCCN_FWD_NO_NORM|DCCN_FWD_EQ_ONLY it is equalizer_ofdm(x)
 * (model.py:349). */
int dccn_forward(dccn_handle* h, const float* x_dev, int64_t B, const uint8_t* bits_dev,
                 float* soft_dev, uint8_t* hard_dev, float* eq_dev, float* chest_dev,
                 int64_t* conf_dev, double* ce_sum_dev, int flags, void* stream);

/* Same pass with HOST buffers (pinned recommended): copies x/bits H2D, runs,
 * copies conf/ce (and hard bits if hard_host != NULL) D2H; synchronises the
 * stream before returning.  conf_host/ce_sum_host are overwritten. */
int dccn_forward_host(dccn_handle* h, const float* x_host, int64_t B, const uint8_t* bits_host,
                      uint8_t* hard_host, int64_t* conf_host, double* ce_sum_host, void* stream);

/* Pipelined form of the same call: two slots; `begin` queues H2D (internal copy stream) + the pass +
 * the D2H of the results and returns immediately, `end` blocks until that slot's results are on the
 * host.  Issuing begin(slot^1) before end(slot) overlaps the next batch's PCIe copy with the pass. */
int dccn_forward_host_begin(dccn_handle* h, int slot, const float* x_host, int64_t B,
                            const uint8_t* bits_host, uint8_t* hard_host, void* stream);
int dccn_forward_host_end(dccn_handle* h, int slot, int64_t* conf_host, double* ce_sum_host);

/* `begin` with the labels packed 8 per byte: bits_packed_host holds B * n_data * nbits / 8 bytes, bit j of byte i
 * = label 8 i + j of the flattened [B, n_data, nbits] array (numpy.packbits(bitorder='little')); the device unpacks
 * them.  The reference feeds `bits_in` as int32 [B, D, nbits] (dev/py/ofdmreceiver_np.py:123) -- 5 120 B per 16-QAM
 * frame; this form moves 160 B, so the host-buffer call is bound by the fp32 IQ alone (4 480 B per frame).
 * hard_host (optional, pinned): the hard decisions [B, n_data, nbits] uint8, copied back on the library's own D2H
 * stream so that they overlap the next batch's pass and H2D copy. */
int dccn_forward_host_begin_packed(dccn_handle* h, int slot, const float* x_host, int64_t B,
                                   const uint8_t* bits_packed_host, uint8_t* hard_host, void* stream);

/* -- a1: layers_conv2d_complex(inputs, filters, kernal, strides=1, padding)
 * (dev/py/complex.py:140-196), op-level.  x_dev [B,L,W,C,2], kernel_dev
 * [kl,kw,1,C,2*filters], bias_dev [2*filters], y_dev [B,L',W',filters,2];
 * padding 0 = 'valid', 1 = 'same'. */
int dccn_cconv2d(const float* x_dev, int64_t B, int L, int W, int C,
                 const float* kernel_dev, const float* bias_dev, int filters, int kl, int kw,
                 int padding, float* y_dev, void* stream);

/* -- layers_conv2d_vector(inputs, filters, kernal, strides=1, padding) (dev/py/complex.py:199-255), op-level: the
 * layer of equalizer_separateIQ (--opt 7).  x_dev [B,L,W,C,2], kernel_dev [kl,kw,2,C,2*filters] (conv3d kernel of depth
 * 2 across IQ), bias_dev [2*filters], y_dev [B,L',W',filters,2]; padding 0 = 'valid', 1 = 'same'. */
int dccn_vconv2d(const float* x_dev, int64_t B, int L, int W, int C,
                 const float* kernel_dev, const float* bias_dev, int filters, int kl, int kw,
                 int padding, float* y_dev, void* stream);

/* -- a6 + a7: rayleigh_chan_lte static branch (dev/py/radio.py:432-437,491-506)
 * followed by AWGN_channel_np (dev/py/radio.py:513-526).
 *   tx_dev      float32 [B, n_samp, 2]  complex IQ of the transmitted frames (n_samp = S*T)
 *   alpha_dev   float64 [n_taps, n_fir] interpolation matrix (3gpp/AM_*.csv); NULL => identity 1x1
 *   coeff_dev   float64 [n_taps]        path amplitudes ch_coeff (radio.py:367-371)
 *   z_dev       float64 [B, n_taps, 2]  pre-drawn N(0,1/2) path gains, or NULL => Philox(seed)
 *   snr_db_dev  float32 [B]             per-frame SNR in dB
 *   normals_dev float64 [B, n_samp, 2]  pre-drawn N(0,1) noise, or NULL => Philox(seed)
 *   rx_dev      float32 [B, n_samp, 2]  output (normalised by the batch mean power + noise)
 *   fir_only_dev float32 [B, n_samp, 2] optional: the faded signal before AWGN (complex64 like radio.py:492)
 * n_taps == 0 skips the fading (AWGN channel). */
int dccn_chan_fir_awgn(dccn_handle* h, const float* tx_dev, int64_t B, int n_samp,
                       const double* alpha_dev, const double* coeff_dev, int n_taps, int n_fir,
                       const double* z_dev, const float* snr_db_dev, const double* normals_dev,
                       uint64_t seed, float* rx_dev, float* fir_only_dev, void* stream);

/* -- f-2: the two halves separately, plus the mobile (Doppler) branch and per-frame profile cycling.
 * dccn_chan_fading: rayleigh_chan_lte.channel for frames frame0, frame0+fstride, ... (mixRayleigh deals
 * frames i%4 to flat/etu/eva/epa, dev/py/radio.py:450-467).  doppler_hz > 0 selects doppler_channel
 * (dev/py/radio.py:399-422); then z_or_theta_dev is float64 [B,2,48,n_taps] uniform(0,2pi) phases, else
 * float64 [B,n_taps,2] N(0,1/2) path gains (NULL -> Philox(seed)).  The batch power sum |rx|^2 is
 * accumulated in the handle (reset_power != 0 clears it first); dccn_chan_awgn normalises by it. */
int dccn_chan_fading(dccn_handle* h, const float* tx_dev, int64_t B, int n_sym, int n_sc,
                     const double* alpha_dev, const double* coeff_dev, int n_taps, int n_fir,
                     double doppler_hz, double sample_rate, const double* z_or_theta_dev, uint64_t seed,
                     int64_t frame0, int64_t fstride, int reset_power, float* faded_dev, void* stream);
int dccn_chan_awgn(dccn_handle* h, const float* faded_dev, int64_t B, int n_samp,
                   const float* snr_db_dev, const double* normals_dev, uint64_t seed, float* rx_dev,
                   void* stream);

/* -- a5: tf.confusion_matrix(bits, argmax) (ofdmreceiver_np.py:165-169); conf accumulated */
int dccn_ber_accum(const uint8_t* hard_dev, const uint8_t* bits_dev, int64_t n,
                   int64_t* conf_dev, void* stream);

/* -- next row f-1: OFDM transmitter on the GPU (dev/py/ofdm.py:328-380).
 *   bits_dev  uint8 [B, D, nbits]; data_sc/pilot_sc: frame-level subcarrier indices (s*K+k)
 *   constellation float32 [2^nbits, 2]; tx_dev float32 [B, S, K+CP, 2] */
int dccn_tx_frames(dccn_handle* h, const uint8_t* bits_dev, int64_t B,
                   const int32_t* data_sc_dev, int n_data, const int32_t* pilot_sc_dev, int n_pilot,
                   const float* constellation_dev, float pilot_re, float pilot_im,
                   float* tx_dev, void* stream);

/* Fused feeder for nfft = 64: dccn_tx_frames followed by the static branch of dccn_chan_fading on every frame, in ONE
 * kernel -- the transmitted frame never touches HBM.  Replaces the call pair
 *   iq_tx_cmpx, test_xs, _ = ofdmobj.ofdm_tx_frame_np(test_ys); test_xs, _ = fading.run(iq_tx_cmpx)
 * (dev/py/ofdmreceiver_np.py:227-228, dev/py/ofdmreceiver_np_mp.py:84-87) for a single static profile.  Bit-identical to
 * the two separate calls.  Accumulates the batch power for a following dccn_chan_awgn (reset_power != 0 zeroes it first).
 * tx_dev: optional copy of the transmitted frames [B,S,T,2] (NULL = not wanted).  z_dev: optional injected path gains
 * [B, n_taps, 2] float64 (NULL = Philox draws from `seed`, the same stream dccn_chan_fading uses). */
int dccn_tx_fade(dccn_handle* h, const uint8_t* bits_dev, int64_t B, const int32_t* data_sc_dev, int n_data,
                 const int32_t* pilot_sc_dev, int n_pilot, const float* constellation_dev, float pilot_re, float pilot_im,
                 const double* alpha_dev, const double* coeff_dev, int n_taps, int n_fir, const double* z_dev,
                 uint64_t seed, int reset_power, float* tx_dev, float* faded_dev, void* stream);

/* -- BASELINE config 4: transfer learning of the equalizer in front of the frozen receiver.
 * Replaces `session.run([train_op, ce_mean, ...], {x, y})` (dev/py/ofdmreceiver_np_mp.py:419) with the graph of
 * dev/py/ofdmreceiver_np_mp.py:335-347:  total_loss = ce_mean + reg_coeff * sum(l2 * sum(w^2)) over kernel+bias of
 * the six tf.layers.dense of equalizer_ofdm, AdamOptimizer.minimize(total_loss, var_list = Equalizer/ *).
 * The receiver variables (fft_like/ *, demodulation/ *) stay frozen. */
typedef struct dccn_train_cfg {
  float reg_coeff;    /* REG_COEFF, 0.001 (ofdmreceiver_np_mp.py:337)                        */
  float l2;           /* tf.keras.regularizers.l2(l=0.01) (model.py:372-373)                  */
  float beta1, beta2; /* tf.train.AdamOptimizer defaults 0.9, 0.999                           */
  float eps;          /* 1e-8                                                                 */
  int64_t max_batch;  /* largest B of dccn_train_step (<= chunk_frames)                       */
  int32_t mode;       /* DCCN_TRAIN_EQ (above) or DCCN_TRAIN_RX (below)                       */
} dccn_train_cfg;
/* DCCN_TRAIN_RX: training of the basic receiver itself, `session.run([train_op, ...])` of
 * dev/py/ofdmreceiver_np.py:234 with the graph of :154-189 -- every ofdm_dense_rx variable is trainable
 * (fft_like/conv3d, demodulation/dense, demodulation/conv2d, demodulation/dense_1; kernel + bias),
 * total_loss = ce_mean + berlin * REG_COEFF * sum(l2 * sum(w^2) over demodulation/dense and demodulation/dense_1)
 * [+ BER_COEFF * ber, which has no gradient], berlin = BER of the minibatch, REG_COEFF = 1e-4 (pass it as reg_coeff).
 * The handle must have been created without equalizer.  A DCCN_TRAIN_RX step synchronises the stream once (the head
 * kernels take their weights by value). */
#define DCCN_TRAIN_EQ 0
#define DCCN_TRAIN_RX 1
/* Allocates optimiser slots / gradient buffers and derives the backward operands; weights must be committed. */
int dccn_train_init(dccn_handle* h, const dccn_train_cfg* cfg, void* stream);
/* One minibatch: forward (batch-moment norm included unless DCCN_FWD_NO_NORM), backward of total_loss w.r.t. every
 * Equalizer/ * variable and, if apply_update != 0, one Adam update with `learning_rate` = the decayed rate of this
 * step (exponential_decay(init, global_step, 500, 0.98, staircase), computed by the caller) and the bias
 * corrections of TF's Adam.  conf_dev (int64 [2,2]) / ce_sum_dev (double) are ACCUMULATED like dccn_forward
 * (training monitors `conf_matrix`, `ce_mean`), either may be NULL. */
int dccn_train_step(dccn_handle* h, const float* x_dev, int64_t B, const uint8_t* bits_dev, float learning_rate,
                    int apply_update, int64_t* conf_dev, double* ce_sum_dev, int flags, void* stream);
/* d total_loss / d var of the last dccn_train_step, reference layout; returns the element count or <0. Synchronous. */
int64_t dccn_train_get_grad(dccn_handle* h, const char* tf_name, float* host, int64_t capacity);
/* number of Adam updates applied so far (TF's global_step); settable when resuming from a checkpoint */
int64_t dccn_train_global_step(const dccn_handle* h);
int dccn_train_set_global_step(dccn_handle* h, int64_t step);

/* -- measurement hooks (bench.py): kernels launched by the library so far; per-kernel CUDA-event
 * timing on the launch stream.  dccn_profile_collect synchronises the device and returns the number
 * of slots; ms_out/count_out[i] = summed duration / launches of slot i since the last collect. */
int64_t dccn_launch_count(void);
int dccn_profile_enable(dccn_handle* h, int on);
int dccn_profile_collect(dccn_handle* h, double* ms_out, int64_t* count_out, int max_slots);
const char* dccn_profile_slot_name(int slot);

/* measurement aid (tools/tma_rate.py): TMA -> shared memory delivery rate per SM, no MMA */
int dccn_debug_tma_rate(const float* mat_dev, int rows, int cols, int ld, int stages, int boxes, int iters,
                        int grid, long long* clks_dev);

/* measurement aid (tools/mma_rate.py): tcgen05.mma kind::tf32 issue / execution rate of one SM */
int dccn_debug_mma_rate(int bn, int n_mma, int per_commit, int mode, int dep, int grid, long long* clks_dev);

/* uniform random bits (util.bit_source, dev/py/util.py:25-34) from Philox */
int dccn_bit_source(uint8_t* bits_dev, int64_t n, uint64_t seed, void* stream);

/* Monitor tensors of the reference graph that are NOT on the receiver's data path but are fetched by its drivers
 * (`session.run([conf_matrix, berlin, power_tx, noise_pwr, ce_mean, iq_tx, iq_rx], ...)`, dev/py/ofdmreceiver_np.py:80;
 * named identities :172-183).  sums_dev[0] = sum over all B * S * T samples of |clip_by_norm(input, 8)|^2 (tx_power =
 * sums[0] / (B S T)), sums_dev[1] = sum of |noise|^2 of the (bypassed) in-graph AWGN channel at snr_db_dev[B]
 * (noise_power; Philox draws from `seed`, so equal to the reference in distribution only; NULL snr_db_dev skips it).
 * Optional outputs (NULL = not wanted): input_dev float [B,S,T,2] = 'input:0'; iq_tx_dev / iq_rx_dev fp16 [B*S*T, 2] =
 * 'iq_tx:0' / 'iq_rx:0' (constellation dumps).  Computes the batch moments of x itself. */
int dccn_monitors(dccn_handle* h, const float* x_dev, int64_t B, const float* snr_db_dev, uint64_t seed, double* sums_dev,
                  float* input_dev, void* iq_tx_dev, void* iq_rx_dev, void* stream);

/* One-shot request for the NEXT dccn_forward on this handle: write equalizer_ofdm's `snr_db` monitor
 * (dev/py/model.py:464-475: log10(clip(mean / variance of |equalized_freq|^2 over the S * P pilot-carrier points, 1e-3, 1e4)))
 * to snr_db_dev[B].  pilot_carriers_dev: ofdm_tx.pilotCarriers (P subcarrier indices, the same in every symbol).
 * snr_db_dev = NULL cancels.  Only with cfg.equalizer, eq_opt 0, layer-by-layer schedule. */
int dccn_forward_monitors(dccn_handle* h, float* snr_db_dev, const int32_t* pilot_carriers_dev, int n_pilot_carriers);

/* Host helper (no GPU involved): CRC-32C (Castagnoli) of `n` bytes continuing from `crc` (0 to start) -- the checksum of
 * every tensor and table block in a TF-bundle checkpoint (`tf.train.Saver`, dev/py/ofdmreceiver_np.py:192,268-272).
 * The Python writer (dl_ofdm_b200/tfbundle.py) saves a 12 MB equalizer checkpoint on every improving epoch; a
 * byte-at-a-time Python loop made that save cost more than the epoch it follows. */
uint32_t dccn_crc32c(const void* data_host, size_t n, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif /* DCCN_H_ */
