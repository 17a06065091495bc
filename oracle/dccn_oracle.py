"""CPU oracle for the DCCN OFDM receiver hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy *restatement* of the reference algorithm
(zhongyuanzhao/dl_ofdm @ 5665b50).  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg; the product (``dl_ofdm_b200``) never imports it.

Why a restatement: the reference's path is a TensorFlow-1.x graph and TensorFlow
cannot be installed in this image (no wheel for Python 3.12, no network), so the
graph itself cannot run here.  Every function below cites the reference lines it
follows, plus the TF-1.15 op semantics it encodes (cross-correlation convs with
SAME pad_before=(k-1)//2, ``tf.layers.dense`` contracting the last axis,
leaky_relu alpha 0.2, softmax over the last axis, argmax first-index ties).

How it is pinned (tests/test_oracle_golden.py):
  * v1 graph: the 8 trained checkpoints the reference ships under
    ``test_v1/model`` decode to the known-answer BER curves of BASELINE.md
    (BER=0 at high SNR, 4.3e-2 at 10 dB for 16-QAM, ~0.24 with the textbook
    complex sign as the negative control).
  * channel / AWGN / transmitter: bit-for-bit against outputs of the reference's
    own NumPy code (``dev/py/radio.py``, ``dev/py/ofdm.py``) imported in the
    build container by ``oracle/make_golden.py`` (fixtures in tests/golden/).
  * dev-architecture receiver and ``equalizer_ofdm`` with *trained* weights:
    **parity unpinned** -- the reference ships no dev checkpoints.  For those
    the oracle is checked against an independent op-for-op mirror of the TF
    graph (``oracle/tf_mirror.py``: padded conv3d + reshape/sub recombination)
    on seeded glorot weights.

All functions take ``dtype``: ``np.float64`` is the "truth" used for tolerances,
``np.float32`` mimics the reference's fp32 arithmetic (different summation order
than Eigen, so only statistically equal to TF's own fp32).
"""
from __future__ import annotations

import numpy as np

LEAKY_ALPHA = 0.2          # tf.nn.leaky_relu default (const 0.2 in the v1 graph)
BN_EPS = 1e-9              # dev/py/ofdmreceiver_np.py:129
LN_EPS = 1e-12             # tf.contrib.layers.layer_norm variance_epsilon (TF 1.15)
SQRT2 = np.sqrt(2.0)


# ---------------------------------------------------------------------------
# a1  layers_conv2d_complex  (dev/py/complex.py:140-196)
# ---------------------------------------------------------------------------
def conv2d_complex(x, kernel, bias, padding='valid', dtype=np.float64):
    """Reference "complex" 2-D convolution.

    x      [B, L, W, C, 2]           (IQ last)
    kernel [kl, kw, 1, C, 2F]        (tf.layers.conv3d layout, depth 1 on IQ)
    bias   [2F]
    returns [B, L', W', F, 2]

    Follows complex.py:168-192: IQ is moved in front of C, ONE real conv3d with
    2F filters is applied to both I and Q, the result [...,2,2F] is re-viewed as
    [...,4,F] and recombined as  re = c0 - c3,  im = c1 - c2, i.e.
        re = xr*Wa - xi*Wb + (ba - bb)
        im = xr*Wb - xi*Wa + (bb - ba)           (NOT the textbook product)
    TF convs are cross-correlations; SAME pads (k-1)//2 before, k//2 after.
    """
    x = np.asarray(x, dtype=dtype)
    kernel = np.asarray(kernel, dtype=dtype)
    bias = np.asarray(bias, dtype=dtype)
    B, L, W, C, two = x.shape
    assert two == 2
    kl, kw, kd, Cin, F2 = kernel.shape
    assert kd == 1 and Cin == C and F2 % 2 == 0
    F = F2 // 2
    if padding.lower() == 'same':
        pl, pw = (kl - 1) // 2, (kw - 1) // 2
        xp = np.zeros((B, L + kl - 1, W + kw - 1, C, 2), dtype=dtype)
        xp[:, pl:pl + L, pw:pw + W] = x
        Lo, Wo = L, W
    elif padding.lower() == 'valid':
        xp = x
        Lo, Wo = L - kl + 1, W - kw + 1
    else:
        raise ValueError(padding)
    # conv[b,l,w,iq,f2] = sum_{i,j,c} xp[b,l+i,w+j,c,iq] * kernel[i,j,0,c,f2]
    conv = np.zeros((B, Lo, Wo, 2, F2), dtype=dtype)
    for i in range(kl):
        for j in range(kw):
            patch = xp[:, i:i + Lo, j:j + Wo]                 # [B,Lo,Wo,C,2]
            if not patch.any():
                continue        # tap sees only SAME-padding zeros (adds exactly 0)
            conv += np.einsum('blwcq,cf->blwqf', patch, kernel[i, j, 0])
    conv += bias
    c4 = conv.reshape(B, Lo, Wo, 4, F)                         # complex.py:185
    re = c4[:, :, :, 0] - c4[:, :, :, 3]                        # complex.py:187
    im = c4[:, :, :, 1] - c4[:, :, :, 2]                        # complex.py:188
    return np.stack([re, im], axis=-1)                          # [B,Lo,Wo,F,2]


def pack_complex_kernel(Wa, Wb, ba, bb, dtype=np.float64):
    """[K,F] (Wa,Wb) + biases -> interleaved real GEMM operands.

    Bp[2k,2f]=Wa  Bp[2k,2f+1]=Wb  Bp[2k+1,2f]=-Wb  Bp[2k+1,2f+1]=-Wa
    bias_p[2f]=ba-bb  bias_p[2f+1]=bb-ba
    """
    K, F = Wa.shape
    Bp = np.zeros((2 * K, 2 * F), dtype=dtype)
    Bp[0::2, 0::2] = Wa
    Bp[0::2, 1::2] = Wb
    Bp[1::2, 0::2] = -Wb
    Bp[1::2, 1::2] = -Wa
    bp = np.zeros(2 * F, dtype=dtype)
    bp[0::2] = ba - bb
    bp[1::2] = bb - ba
    return Bp, bp


# ---------------------------------------------------------------------------
# a2  transmitter normalisation  (dev/py/ofdmreceiver_np.py:128-129)
# ---------------------------------------------------------------------------
def batch_moment_norm(x, dtype=np.float64):
    """tf.nn.moments(x,[0]) + tf.nn.batch_normalization(eps=1e-9) / sqrt(2).

    Moments are over the BATCH axis per (symbol, sample, iq) position.  TF
    evaluates  x*inv + (-mean*inv)  with inv = rsqrt(var+eps).
    """
    x = np.asarray(x, dtype=dtype)
    mean = x.mean(axis=0)
    var = np.mean((x - mean) ** 2, axis=0)
    inv = 1.0 / np.sqrt(var + dtype(BN_EPS))
    z = x * inv + (-mean * inv)
    return (z / dtype(SQRT2)).astype(dtype), mean, inv


def layer_norm(x, dtype=np.float64):
    """tf.contrib.layers.layer_norm(center=False, scale=False, begin_norm_axis=1).

    Used at dev/py/model.py:363: moments over all non-batch axes, eps 1e-12.
    """
    x = np.asarray(x, dtype=dtype)
    ax = tuple(range(1, x.ndim))
    mean = x.mean(axis=ax, keepdims=True)
    var = np.mean((x - mean) ** 2, axis=ax, keepdims=True)
    inv = 1.0 / np.sqrt(var + dtype(LN_EPS))
    return x * inv + (-mean * inv)


def _leaky(x):
    return np.maximum(LEAKY_ALPHA * x, x)         # v1 graph: Maximum(alpha*x, x)


def _softmax2(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


# ---------------------------------------------------------------------------
# a3  ofdm_dense_rx  (dev/py/model.py:1222-1292)  + the v1 variant of the head
# ---------------------------------------------------------------------------
def ofdm_dense_rx(z, w, nbits, cp_len, use_cp=True, head='dev', nfilter=64,
                  dtype=np.float64, return_intermediate=False):
    """Basic receiver on normalised IQ ``z`` [B,S,T,2] -> softmax [B,D,nbits,2].

    ``w`` maps TF variable names to arrays:
      fft_like/conv3d/{kernel [1,K,1,K,2F], bias [2F]}
      demodulation/dense/{kernel [S*F*2, D*2], bias}
      demodulation/conv2d/{kernel [1,1,2,2^nb], bias}
      demodulation/conv2d_1/{kernel [1,1,2^nb,2^nb], bias}    (head == 'v1' only)
      demodulation/dense_1/{kernel [2^nb+2, 2*nb], bias}
    head 'dev' follows model.py:1275-1288; head 'v1' follows the graph stored in
    test_v1/model/*.meta (conv2d -> conv2d_1 -> leaky, per-symbol reshape).
    """
    z = np.asarray(z, dtype=dtype)
    B, S, T, _ = z.shape
    out = z if use_cp else z[:, :, cp_len:, :]                  # model.py:1236-1238
    K = out.shape[2]
    F = nfilter
    conv_in = out.reshape(B, S, 1, K, 2)                        # model.py:1248
    fft = conv2d_complex(conv_in, w['fft_like/conv3d/kernel'],
                         w['fft_like/conv3d/bias'], 'same', dtype)   # [B,S,1,F,2]
    fft = fft.reshape(B, S, F, 2)                               # model.py:1262
    flat = fft.reshape(B, S * F * 2)                            # model.py:1268
    Wd = np.asarray(w['demodulation/dense/kernel'], dtype=dtype)
    bd = np.asarray(w['demodulation/dense/bias'], dtype=dtype)
    out_iq = (flat @ Wd + bd).reshape(B, -1, 2)                 # [B,D,2]  model.py:1275
    Wc = np.asarray(w['demodulation/conv2d/kernel'], dtype=dtype).reshape(2, -1)
    bc = np.asarray(w['demodulation/conv2d/bias'], dtype=dtype)
    h = out_iq @ Wc + bc                                        # 1x1 conv, model.py:1278
    if head == 'v1':
        Wc1 = np.asarray(w['demodulation/conv2d_1/kernel'], dtype=dtype)
        Wc1 = Wc1.reshape(Wc1.shape[2], Wc1.shape[3])
        h = h @ Wc1 + np.asarray(w['demodulation/conv2d_1/bias'], dtype=dtype)
    h = _leaky(h)                                               # model.py:1280
    cat = np.concatenate([h, out_iq], axis=-1)                  # model.py:1282
    W1 = np.asarray(w['demodulation/dense_1/kernel'], dtype=dtype)
    b1 = np.asarray(w['demodulation/dense_1/bias'], dtype=dtype)
    logits = _leaky(cat @ W1 + b1)                              # model.py:1283-1288
    logits = logits.reshape(B, -1, nbits, 2)                    # model.py:1290
    soft = _softmax2(logits)                                    # model.py:1291
    if return_intermediate:
        return soft, dict(fft_out=fft, out_iq=out_iq, logits=logits)
    return soft


# ---------------------------------------------------------------------------
# a4  equalizer_ofdm  (dev/py/model.py:349-478), variables under 'Equalizer/'
# ---------------------------------------------------------------------------
def equalizer_ofdm(z, w, nfft, cp_len, use_cp=True, dtype=np.float64,
                   return_intermediate=False, prefix='Equalizer/'):
    """Channel-equalisation block: normalised IQ [B,S,T,2] -> equalised [B,S,T,2].

    Returns (equalized, chest_complex[B,S,K]).  The snr_db side output
    (model.py:464-475) feeds only monitors and is not restated.
    """
    def g(name):
        return np.asarray(w[prefix + name], dtype=dtype)
    z = np.asarray(z, dtype=dtype)
    B, S, T, _ = z.shape
    K = nfft
    chest = layer_norm(z, dtype)                                # model.py:363
    if not use_cp:
        chest = chest[:, :, cp_len:cp_len + K, :].reshape(B, S, K * 2)   # :365-367
    else:
        chest = chest.reshape(B, S, T * 2)                      # :369
    t1 = chest @ g('dense/kernel') + g('dense/bias')            # :370  [B,S,2K]
    t1 = t1.reshape(B, S, K, 1, 2)                              # :377
    f = conv2d_complex(t1, g('conv3d/kernel'), g('conv3d/bias'), 'valid', dtype)  # [B,S,1,K,2]
    f = np.transpose(f, (0, 1, 3, 2, 4))                        # :379  [B,S,K,1,2]
    inputs_c = f[..., 0] + 1j * f[..., 1]                       # :382  [B,S,K,1]
    flat = f.reshape(B, S * K * 2)                              # :391
    pilot = flat @ g('dense_1/kernel') + g('dense_1/bias')      # :393
    c = pilot @ g('dense_2/kernel') + g('dense_2/bias')         # :401
    c = c @ g('dense_3/kernel') + g('dense_3/bias')             # :407
    c = np.tanh(c @ g('dense_4/kernel') + g('dense_4/bias'))    # :419
    c5 = c.reshape(B, S, K, 1, 2)                               # :425
    c5 = conv2d_complex(c5, g('conv3d_1/kernel'), g('conv3d_1/bias'), 'same', dtype)  # :426 [B,S,K,1,2]
    chest_c = c5[..., 0] + 1j * c5[..., 1]                      # :428  [B,S,K,1]
    ab = np.abs(chest_c)                                        # :431
    conj_n = np.real(chest_c) / ab - 1j * (np.imag(chest_c) / ab)   # :432-433 (no eps)
    eq = inputs_c * conj_n                                      # :434
    corr = eq * np.conj(eq)                                     # :437
    corr5 = np.stack([corr.real, corr.imag], axis=-1)           # [B,S,K,1,2]
    corr_o = conv2d_complex(corr5, g('conv3d_2/kernel'), g('conv3d_2/bias'), 'valid', dtype)  # [B,S,1,K,2]
    corr_o = np.transpose(corr_o, (0, 1, 3, 2, 4))[:, :, :, 0, :]   # :439-440 [B,S,K,2]
    eq5 = np.stack([eq.real, eq.imag], axis=-1)
    eq_o = conv2d_complex(eq5, g('conv3d_3/kernel'), g('conv3d_3/bias'), 'valid', dtype)
    eq_o = np.transpose(eq_o, (0, 1, 3, 2, 4))[:, :, :, 0, :]       # :442-448 [B,S,K,2]
    cat = np.concatenate([eq_o, corr_o], axis=-1)               # :455  [B,S,K,4]
    cat = cat.reshape(B, S, K * 4)                              # :456
    out = cat @ g('dense_5/kernel') + g('dense_5/bias')         # :457
    out = out.reshape(B, S, T, 2)                               # :462
    chest_out = chest_c.reshape(B, S, K)                        # :477
    if return_intermediate:
        return out, chest_out, dict(inputs_complex=inputs_c[..., 0], eq=eq[..., 0], t1=t1,
                                    pilot=pilot, tanh=c)
    return out, chest_out


# ---------------------------------------------------------------------------
# ablation equalizers, --opt != 0  (dev/py/ofdmreceiver_np_mp.py:292-311)
# ---------------------------------------------------------------------------
# All of them are rewirings of equalizer_ofdm's template:
#   layer_norm -> [CP slice] -> dense(2K) per symbol -> FRONT2 -> inputs_complex
#   -> flatten -> dense(pilot) -> CHAIN of frame-level dense layers (linear / tanh) -> [(S,K) 'same' complex conv]
#   -> chest -> phase-only equalise -> TAIL -> [B,S,T,2]
# opt  function (dev/py/model.py)        FRONT2        CHAIN (0 lin, 1 tanh)   conv   TAIL
#  0   equalizer_ofdm      :349-478      cconv (1,K)   0 0 1                   yes    cconv(eq) | cconv(corr) -> concat -> dense
#  1   equalizer_nocconv   :482-609      dense(2K)     0 0 1                   yes    dense(2K) -> dense(2T)
#  2   equalizer_noresdl   :612-714      cconv (1,K)   0                       no     tf.ifft -> dense(2T)
#  4   equalizer_noresdl2  :718-826      cconv (1,K)   0 1                     no     tf.ifft -> dense(2T)
#  5   equalizer_noresdl4  :829-950      cconv (1,K)   0 1 1 1                 no     tf.ifft -> dense(2T)
#  3   equalizer_dnnE      :953-1084     dense(2K)     1 1 1 1                 no     dense(2K) -> dense(2T)
#  7   equalizer_separateIQ:1088-1218    vconv (1,K)   1 1 1                   yes(v) vconv(eq) | vconv(corr) -> concat -> dense   [layers_conv2d_vector]
# (opt 6 names equalizer_doppler, which does not exist in dev/py/model.py -- the reference itself raises NameError; opt 9 / 10 are equalizer_ofdm.)
# Variable names follow TF-1's per-scope auto-numbering in creation order (dense, dense_1, ...; conv3d, conv3d_1, ...).
def conv2d_vector(x, kernel, bias, padding='valid', dtype=np.float64):
    """layers_conv2d_vector (dev/py/complex.py:199-255): one real conv3d over (length, width, IQ) with kernel depth 2
    across IQ and 2*filters channels, NO complex recombination.
    x [B,L,W,1,2] (one input channel), kernel [kl,kw,2,1,2F], bias [2F] -> [B,L',W',F,2].
    'valid': the IQ axis collapses to one position, channels [0,F) are the real parts, [F,2F) the imaginary parts
    (reshape [..,1,2F] -> [..,2,F], :245-249).  'same' (used with F = 1): TF pads the size-2 IQ axis (0 before, 1 after);
    the reshape merges (IQ position, channel) and keeps merged indices 0 and 1 = IQ position 0, channels 0 and 1."""
    x = np.asarray(x, dtype=dtype)
    k = np.asarray(kernel, dtype=dtype)
    b = np.asarray(bias, dtype=dtype)
    B, L, W, C, _ = x.shape
    assert C == 1 and k.shape[2] == 2 and k.shape[3] == 1
    kl, kw = k.shape[0], k.shape[1]
    F = k.shape[4] // 2
    xs = x[:, :, :, 0, :]                                            # [B,L,W,2]
    if padding == 'valid':
        Lo, Wo = L - kl + 1, W - kw + 1
        xp = xs
    else:
        assert F == 1, "'same' vector conv is only used with one filter"
        Lo, Wo = L, W
        pl, pw = (kl - 1) // 2, (kw - 1) // 2
        xp = np.zeros((B, L + kl - 1, W + kw - 1, 2), dtype=dtype)
        xp[:, pl:pl + L, pw:pw + W, :] = xs
    out = np.zeros((B, Lo, Wo, 2 * F), dtype=dtype)
    for i in range(kl):
        for j in range(kw):
            out += xp[:, i:i + Lo, j:j + Wo, :] @ k[i, j, :, 0, :]   # [.,2] @ [2,2F]
    out = out + b
    return np.stack([out[..., :F], out[..., F:]], axis=-1)           # [B,Lo,Wo,F,2]


EQ_SPECS = {
    7: dict(front2='vconv', chain=(1, 1, 1), toeplitz=True, tail='corr', vector=True),   # equalizer_separateIQ :1088-1218
    0: dict(front2='cconv', chain=(0, 0, 1), toeplitz=True, tail='corr'),
    1: dict(front2='dense', chain=(0, 0, 1), toeplitz=True, tail='dense2'),
    2: dict(front2='cconv', chain=(0,), toeplitz=False, tail='ifft'),
    4: dict(front2='cconv', chain=(0, 1), toeplitz=False, tail='ifft'),
    5: dict(front2='cconv', chain=(0, 1, 1, 1), toeplitz=False, tail='ifft'),
    3: dict(front2='dense', chain=(1, 1, 1, 1), toeplitz=False, tail='dense2'),
}


def eq_layer_names(opt):
    """Creation-ordered layer list of ``--opt``: [(role, tf layer name)] with roles front1, front2, pilot, chain<i>,
    toeplitz, tail_corr, tail_eq, tail1, tail2."""
    sp = EQ_SPECS[opt]
    nd = nc = 0
    out = []

    def dense(role):
        nonlocal nd
        out.append((role, 'dense' if nd == 0 else 'dense_%d' % nd))
        nd += 1

    def conv(role):
        nonlocal nc
        out.append((role, 'conv3d' if nc == 0 else 'conv3d_%d' % nc))
        nc += 1

    dense('front1')
    conv('front2') if sp['front2'] in ('cconv', 'vconv') else dense('front2')
    dense('pilot')
    for i in range(len(sp['chain'])):
        dense('chain%d' % i)
    if sp['toeplitz']:
        conv('toeplitz')
    if sp['tail'] == 'corr':
        conv('tail_corr')
        conv('tail_eq')
        dense('tail2')
    elif sp['tail'] == 'dense2':
        dense('tail1')
        dense('tail2')
    else:
        dense('tail2')
    return out


def equalizer_variant(z, w, opt, nfft, cp_len, use_cp=True, dtype=np.float64, prefix='Equalizer/'):
    """Ablation equalizers (and opt 0) op by op in complex arithmetic: normalised IQ [B,S,T,2] -> (equalised
    [B,S,T,2], chest complex [B,S,K]).  Line numbers: see the table above."""
    sp = EQ_SPECS[opt]
    names = dict(eq_layer_names(opt))

    def g(role, what):
        return np.asarray(w[prefix + names[role] + '/' + what], dtype=dtype)

    def dense(x, role, act=0):
        y = x @ g(role, 'kernel') + g(role, 'bias')
        return np.tanh(y) if act else y

    z = np.asarray(z, dtype=dtype)
    B, S, T, _ = z.shape
    K = nfft
    c = layer_norm(z, dtype)
    c = c[:, :, cp_len:cp_len + K, :].reshape(B, S, K * 2) if not use_cp else c.reshape(B, S, T * 2)
    c = dense(c, 'front1')                                              # [B,S,2K]
    vec = sp.get('vector', False)
    cconv = conv2d_vector if vec else conv2d_complex
    if sp['front2'] in ('cconv', 'vconv'):
        f = cconv(c.reshape(B, S, K, 1, 2), g('front2', 'kernel'), g('front2', 'bias'), 'valid', dtype)
        f = np.transpose(f, (0, 1, 3, 2, 4))[:, :, :, 0, :]             # [B,S,K,2]
    else:
        f = dense(c, 'front2').reshape(B, S, K, 2)
    inputs_c = f[..., 0] + 1j * f[..., 1]                               # [B,S,K]
    c = dense(f.reshape(B, S * K * 2), 'pilot')
    for i, act in enumerate(sp['chain']):
        c = dense(c, 'chain%d' % i, act)
    c5 = c.reshape(B, S, K, 1, 2)
    if sp['toeplitz']:
        c5 = cconv(c5, g('toeplitz', 'kernel'), g('toeplitz', 'bias'), 'same', dtype)
    chest_c = (c5[..., 0] + 1j * c5[..., 1])[:, :, :, 0]                # [B,S,K]
    ab = np.abs(chest_c)
    eq = inputs_c * (np.real(chest_c) / ab - 1j * (np.imag(chest_c) / ab))     # phase-only equalise, no eps
    if sp['tail'] == 'corr':
        corr = eq * np.conj(eq)
        def cc(v, role):
            v5 = np.stack([v.real, v.imag], -1).reshape(B, S, K, 1, 2)
            o = cconv(v5, g(role, 'kernel'), g(role, 'bias'), 'valid', dtype)
            return np.transpose(o, (0, 1, 3, 2, 4))[:, :, :, 0, :]
        t = np.concatenate([cc(eq, 'tail_eq'), cc(corr, 'tail_corr')], -1).reshape(B, S, K * 4)
    elif sp['tail'] == 'dense2':
        t = dense(np.stack([eq.real, eq.imag], -1).reshape(B, S, K * 2), 'tail1')
    else:
        e = np.fft.ifft(eq, axis=-1)                                    # tf.ifft over the K subcarriers of a symbol
        t = np.stack([e.real, e.imag], -1).reshape(B, S, K * 2).astype(dtype)
    out = dense(t, 'tail2').reshape(B, S, T, 2)
    return out, chest_c


# ---------------------------------------------------------------------------
# a5  BER head / loss  (dev/py/ofdmreceiver_np.py:154-169, dev/py/util.py:44-48)
# ---------------------------------------------------------------------------
def ber_head(soft, bits):
    """argmax (first index on ties) + 2x2 confusion matrix + double-softmax CE.

    Returns (hard uint8 [B,D,nb], conf int64 [2,2] (rows = truth), ber, ce_mean).
    """
    soft = np.asarray(soft)
    hard = (soft[..., 1] > soft[..., 0]).astype(np.uint8)       # argmax, ties -> 0
    y = np.asarray(bits).reshape(-1).astype(np.int64)
    o = hard.reshape(-1).astype(np.int64)
    conf = np.zeros((2, 2), dtype=np.int64)
    np.add.at(conf, (y, o), 1)                                  # tf.confusion_matrix(labels, pred)
    ber = float(conf[0, 1] + conf[1, 0]) / float(conf.sum())
    p = soft.reshape(-1, 2).astype(np.float64)
    lse = np.log(np.exp(p[:, 0]) + np.exp(p[:, 1]))             # softmax_xent on softmax outputs
    ce = float(np.mean(lse - p[np.arange(p.shape[0]), y]))
    return hard, conf, ber, ce


# ---------------------------------------------------------------------------
# a6  rayleigh_chan_lte static branch  (dev/py/radio.py:339-372, 432-437, 491-506)
# ---------------------------------------------------------------------------
TAP_PROFILES = {   # radio.py:340-366  (delay ns, power dB)
    'etu': ([0, 50, 120, 200, 230, 500, 1600, 2300, 5000],
            [-1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -3.0, -5.0, -7.0]),
    'epa': ([0, 30, 70, 90, 110, 190, 410],
            [0.0, -1.0, -2.0, -3.0, -8.0, -17.2, -20.8]),
    'eva': ([0, 30, 150, 310, 370, 710, 1090, 1730, 2510],
            [0.0, -1.5, -1.4, -3.6, -0.6, -9.1, -7.0, -12.0, -16.9]),
    'custom': ([0, 70, 200, 230, 500, 1600, 2700, 3000],
               [0.0, -1.4, -1.4, -1.0, -3.0, -9.1, -15.0, -19.0]),
    'flat': ([0], [0.0]),
}


def channel_coeff(chan):
    """radio.py:367-371: linear POWER / sqrt(sum power) used as the path amplitude."""
    pw = 10.0 ** (np.asarray(TAP_PROFILES[chan.lower()][1], dtype=np.float64) / 10.0)
    return pw * (1.0 / np.sqrt(np.sum(pw)))


def rayleigh_static(tx, z, ch_coeff, alpha):
    """Static (non-Doppler) Rayleigh FIR for a batch of frames.

    tx [B, S*T] complex, z [B, n_taps] complex CN(0,1) path gains,
    ch_coeff [n_taps], alpha [n_taps, N_fir].  Per frame (radio.py:432-437):
        g  = (z * ch_coeff) @ alpha
        rx = np.convolve(tx, g, 'same')          (centred, zero history)
    Stored as complex64 like radio.py:492, returned as float [B, S*T, 2].
    """
    tx = np.asarray(tx)
    B, N = tx.shape
    out = np.zeros((B, N), dtype=np.complex64)
    gs = (np.asarray(z) * ch_coeff) @ np.asarray(alpha, dtype=np.float64)
    for b in range(B):
        out[b] = np.convolve(tx[b], gs[b], mode='same')
    return np.stack([out.real, out.imag], axis=-1), gs


def rayleigh_doppler(tx, theta, Fd, ch_coeff, alpha, n_sym, n_sc, sample_rate, ss=48):
    """Mobile (Doppler) branch, dev/py/radio.py:387-422: sum-of-sinusoids path gains re-drawn per
    OFDM symbol, per-symbol 'same' convolution over the symbol plus n_taps samples of history.

    tx [B, n_sym*n_sc] complex; theta [B, 2, ss, n_taps] uniform(0, 2pi) phases (re / im branch).
    """
    tx = np.asarray(tx)
    B = tx.shape[0]
    n_taps = len(ch_coeff)
    alpha = np.asarray(alpha, dtype=np.float64)
    k_vec = np.arange(1, n_taps + 1)
    n_vec = (np.arange(1, ss + 1).reshape(ss, 1) - 0.5) * np.pi / (4 * ss)        # :389
    a0 = k_vec * np.pi / (4 * ss)
    f_re = Fd * np.cos(n_vec + a0)                                                 # :392
    f_im = Fd * np.cos(n_vec - a0)
    const1 = np.sqrt(1.0 / ss)
    t_sym = n_sc / sample_rate                                                     # :407
    out = np.zeros((B, n_sym * n_sc), dtype=np.complex64)
    for b in range(B):
        pre = np.zeros(n_taps + n_sym * n_sc, dtype=np.complex64)                  # :402-403
        pre[n_taps:] = tx[b]
        for i in range(n_sym):
            t = i * t_sym
            mu_re = const1 * np.sum(np.cos(2 * np.pi * t * f_re + theta[b, 0]), 0)  # :411-414
            mu_im = const1 * np.sum(np.cos(2 * np.pi * t * f_im + theta[b, 1]), 0)
            g = ((mu_re + 1j * mu_im) * ch_coeff) @ alpha                          # :417-418
            roll = pre[n_sc * i: n_taps + n_sc * (i + 1)]                          # :419
            out[b, n_sc * i:n_sc * (i + 1)] = np.convolve(roll, g, mode='same')[n_taps:]   # :420-421
    return np.stack([out.real, out.imag], axis=-1)


# ---------------------------------------------------------------------------
# a7  AWGN_channel_np  (dev/py/radio.py:513-526)
# ---------------------------------------------------------------------------
def awgn(x, snr_db, normals):
    """x [B,S,T,2] -> x/sqrt(mean over the WHOLE batch of I^2+Q^2) + n.

    ``normals`` are the standard-normal draws ([B,S,T,2]); std per component is
    sqrt(0.5)*10^(-SNR/20) with SNR per frame ([B,1]).  float64 like the reference.
    """
    x = np.asarray(x, dtype=np.float64)
    pw = np.square(x[..., 0:1]) + np.square(x[..., 1:])
    savg = np.nanmean(pw)
    xn = x / np.sqrt(savg)
    std = np.sqrt(0.5) * np.power(10.0, -np.asarray(snr_db, dtype=np.float64) / 20.0)
    noise = np.asarray(normals, dtype=np.float64) * std.reshape(-1, 1, 1, 1)
    out = xn + noise
    npw = np.mean(np.square(noise[..., 0:1]) + np.square(noise[..., 1:]))
    return out, float(npw), float(savg)


# ---------------------------------------------------------------------------
# full pipelines as the reference wires them
# ---------------------------------------------------------------------------
def basic_receiver(x, w, nbits, cp_len, use_cp=True, head='dev', nfilter=64,
                   dtype=np.float64):
    """tx_ofdm -> a2 -> a3   (dev/py/ofdmreceiver_np.py:121-146)."""
    z, _, _ = batch_moment_norm(x, dtype)
    return ofdm_dense_rx(z, w, nbits, cp_len, use_cp, head, nfilter, dtype)


def equalized_receiver(x, w, nbits, nfft, cp_len, use_cp=True, nfilter=64,
                       dtype=np.float64, opt=0):
    """tx_ofdm -> a2 -> a4 -> (+0) -> a3   (dev/py/ofdmreceiver_np_mp.py:292-320)."""
    z, _, _ = batch_moment_norm(x, dtype)
    if opt == 0:
        eq, chest = equalizer_ofdm(z, w, nfft, cp_len, use_cp, dtype)
    else:
        eq, chest = equalizer_variant(z, w, opt, nfft, cp_len, use_cp, dtype)
    soft = ofdm_dense_rx(eq, w, nbits, cp_len, use_cp, 'dev', nfilter, dtype)
    return soft, eq, chest


# ---------------------------------------------------------------------------
# weights: TF default initialisers (glorot_uniform kernels, zero biases)
# ---------------------------------------------------------------------------
def glorot_weights(rng, nbits, nfft=64, cp_len=16, nsymbol=7, nfilter=64, n_data=320,
                   pilot_size=16, use_cp=True, head='dev', equalizer=True,
                   bias_scale=0.0, chest_bias=None, eq_opt=0):
    """Seeded weights with the reference's variable names, layouts and init.

    ``bias_scale`` > 0 draws small non-zero biases so parity tests exercise the
    bias paths (TF initialises biases to zero; trained models have them non-zero).
    ``chest_bias`` = (b0, b1) sets Equalizer/conv3d_1/bias so that the channel
    estimate stays away from 0: the phase-only equaliser divides by |chest| with
    no epsilon (model.py:430-433), and an untrained estimate that crosses 0 makes
    the frame ill-conditioned in every arithmetic.
    """
    T = nfft + cp_len if use_cp else nfft
    S, F, K = nsymbol, nfilter, nfft
    M = 2 ** nbits

    def glorot(shape, fan_in, fan_out):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    def bias(n):
        if bias_scale == 0.0:
            return np.zeros(n, dtype=np.float32)
        return (bias_scale * rng.standard_normal(n)).astype(np.float32)

    w = {}

    def conv3d(name, kl, kw, cin, cout):
        rf = kl * kw
        w[name + '/kernel'] = glorot((kl, kw, 1, cin, cout), rf * cin, rf * cout)
        w[name + '/bias'] = bias(cout)

    def dense(name, i, o):
        w[name + '/kernel'] = glorot((i, o), i, o)
        w[name + '/bias'] = bias(o)

    conv3d('fft_like/conv3d', 1, T, T, 2 * F)
    dense('demodulation/dense', S * F * 2, n_data * 2)
    w['demodulation/conv2d/kernel'] = glorot((1, 1, 2, M), 2, M)
    w['demodulation/conv2d/bias'] = bias(M)
    if head == 'v1':
        w['demodulation/conv2d_1/kernel'] = glorot((1, 1, M, M), M, M)
        w['demodulation/conv2d_1/bias'] = bias(M)
    dense('demodulation/dense_1', M + 2, 2 * nbits)
    if equalizer and eq_opt != 0:
        e = 'Equalizer/'
        sp = EQ_SPECS[eq_opt]
        SK2 = S * K * 2
        shapes = {'front1': ('d', T * 2, K * 2), 'front2': ('c', 1, K, 2 * K) if sp['front2'] in ('cconv', 'vconv') else ('d', K * 2, K * 2),
                  'pilot': ('d', SK2, pilot_size * 2), 'toeplitz': ('c', S, K, 2), 'tail_corr': ('c', 1, K, 2 * K),
                  'tail_eq': ('c', 1, K, 2 * K), 'tail1': ('d', K * 2, K * 2),
                  'tail2': ('d', K * 4 if sp['tail'] == 'corr' else K * 2, (nfft + cp_len) * 2)}
        for i in range(len(sp['chain'])):
            shapes['chain%d' % i] = ('d', pilot_size * 2 if i == 0 else SK2, SK2)
        layers = eq_layer_names(eq_opt)
        for role, name in layers:
            sh = shapes[role]
            if sh[0] == 'd':
                dense(e + name, sh[1], sh[2])
            elif sp.get('vector', False):      # layers_conv2d_vector: conv3d kernel [kl, kw, 2, 1, 2F]
                rf = sh[1] * sh[2] * 2
                w[e + name + '/kernel'] = glorot((sh[1], sh[2], 2, 1, sh[3]), rf, rf * sh[3])
                w[e + name + '/bias'] = bias(sh[3])
            else:
                conv3d(e + name, sh[1], sh[2], 1, sh[3])
        if chest_bias is not None:
            # keep the channel estimate away from 0 (no epsilon in the phase-only equaliser)
            names = dict(layers)
            if sp['toeplitz']:
                w[e + names['toeplitz'] + '/bias'] = np.asarray(chest_bias, dtype=np.float32)
            else:
                last = e + names['chain%d' % (len(sp['chain']) - 1)]
                w[last + '/kernel'] = (0.3 * w[last + '/kernel']).astype(np.float32)
                b = np.empty(SK2, dtype=np.float32)
                b[0::2], b[1::2] = chest_bias[0], chest_bias[1]
                w[last + '/bias'] = b
    elif equalizer:
        e = 'Equalizer/'
        dense(e + 'dense', T * 2, K * 2)
        conv3d(e + 'conv3d', 1, K, 1, 2 * K)
        dense(e + 'dense_1', S * K * 2, pilot_size * 2)
        dense(e + 'dense_2', pilot_size * 2, S * K * 2)
        dense(e + 'dense_3', S * K * 2, S * K * 2)
        dense(e + 'dense_4', S * K * 2, S * K * 2)
        conv3d(e + 'conv3d_1', S, K, 1, 2)
        conv3d(e + 'conv3d_2', 1, K, 1, 2 * K)
        conv3d(e + 'conv3d_3', 1, K, 1, 2 * K)
        dense(e + 'dense_5', K * 4, (nfft + cp_len) * 2)
        if chest_bias is not None:
            w[e + 'conv3d_1/bias'] = np.asarray(chest_bias, dtype=np.float32)
    return w
