"""CPU oracle for the equalizer transfer-learning step (BASELINE.json config 4)
--  TEST INFRASTRUCTURE ONLY (same import rules as dccn_oracle.py).

What the reference does (dev/py/ofdmreceiver_np_mp.py:292-347):

  * graph: tx_ofdm -> batch-moment norm -> equalizer_ofdm (scope 'Equalizer') -> (+0) ->
    frozen ofdm_dense_rx -> softmax -> ce_mean  (dev/py/ofdmreceiver_np.py:154-162: softmax-xent
    applied ON the softmax outputs, mean over every bit of the batch);
  * total_loss = ce_mean + REG_COEFF * sum(REGULARIZATION_LOSSES), REG_COEFF = 0.001 (:336-341); the
    regularisers are tf.keras.regularizers.l2(l=0.01) = 0.01 * sum(w^2) on kernel AND bias of the six
    tf.layers.dense of equalizer_ofdm (dev/py/model.py:370-461); the conv3d layers have none;
  * optimizer.minimize(total_loss, var_list=Equalizer vars) with tf.train.AdamOptimizer(lr),
    lr = exponential_decay(init_learning, global_step, 500, 0.98, staircase=True) (:343-347).

TensorFlow's autodiff itself is not in the tree, so the backward pass below is a hand-derived NumPy
restatement in GEMM form (complex layers packed per SURVEY App. D).  It is pinned by
tests/test_train_oracle.py against torch autograd run through the independent op-for-op mirror of
the TF graph (oracle/tf_mirror.py, padded conv3d formulation) in float64: **parity unpinned** against
TF itself (no trained dev checkpoint or gradient dump is shipped by the reference).

TF-1.15 semantics encoded here: AdamOptimizer update  lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
m = b1*m + (1-b1)*g, v = b2*v + (1-b2)*g^2, w -= lr_t*m/(sqrt(v)+eps)  with b1=.9, b2=.999, eps=1e-8;
LeakyRelu gradient = g where x > 0 else alpha*g; gradients of complex ops for a real loss equal the
real-variable gradients w.r.t. (re, im).
"""
from __future__ import annotations

import numpy as np

from . import dccn_oracle as orc

REG_COEFF = 0.001            # dev/py/ofdmreceiver_np_mp.py:337
L2_L = 0.01                  # tf.keras.regularizers.l2(l=0.01), dev/py/model.py:372-373
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8
DENSE_NAMES = ['dense', 'dense_1', 'dense_2', 'dense_3', 'dense_4', 'dense_5']
CONV_NAMES = ['conv3d', 'conv3d_1', 'conv3d_2', 'conv3d_3']


def trainable_names(prefix='Equalizer/'):
    return [prefix + n + s for n in DENSE_NAMES + CONV_NAMES for s in ('/kernel', '/bias')]


def learning_rate(init_learning, global_step):
    """tf.train.exponential_decay(init, step, 500, 0.98, staircase=True)  (_mp.py:343-344)."""
    return init_learning * 0.98 ** (global_step // 500)


# ---------------------------------------------------------------------------------------------
# packing of the reference layouts into GEMM operands and the adjoint maps (gradients back)
# ---------------------------------------------------------------------------------------------
def _pack_1xk(kernel, bias):
    k = kernel[0, :, 0, 0, :]
    F = k.shape[1] // 2
    return orc.pack_complex_kernel(k[:, :F], k[:, F:], bias[:F], bias[F:], dtype=kernel.dtype)


def _unpack_1xk(dBp, dbp, shape):
    """adjoint of _pack_1xk: d(kernel [1,K,1,1,2F]), d(bias [2F])."""
    gWa = dBp[0::2, 0::2] - dBp[1::2, 1::2]
    gWb = dBp[0::2, 1::2] - dBp[1::2, 0::2]
    gk = np.concatenate([gWa, gWb], axis=1).reshape(shape)
    gba = dbp[0::2] - dbp[1::2]
    return gk, np.concatenate([gba, -gba])


def _pack_toeplitz(kernel, bias, S, K):
    """(S,K) 'same' complex conv with one filter as a dense [2SK, 2SK] matrix."""
    n = S * K * 2
    Bp = np.zeros((n, n), dtype=kernel.dtype)
    pl, pw = (S - 1) // 2, (K - 1) // 2
    for i in range(S):
        for j in range(K):
            wa, wb = kernel[i, j, 0, 0, 0], kernel[i, j, 0, 0, 1]
            d = np.arange(max(0, pl - i), min(S, S + pl - i))
            h = np.arange(max(0, pw - j), min(K, K + pw - j))
            co = ((d[:, None] * K + h[None, :]) * 2).ravel()
            ri = (((d[:, None] + i - pl) * K + (h[None, :] + j - pw)) * 2).ravel()
            Bp[ri, co] = wa
            Bp[ri, co + 1] = wb
            Bp[ri + 1, co] = -wb
            Bp[ri + 1, co + 1] = -wa
    bp = np.zeros(n, dtype=kernel.dtype)
    bp[0::2] = bias[0] - bias[1]
    bp[1::2] = bias[1] - bias[0]
    return Bp, bp


def _unpack_toeplitz(dBp, dbp, S, K):
    gk = np.zeros((S, K, 1, 1, 2), dtype=dBp.dtype)
    pl, pw = (S - 1) // 2, (K - 1) // 2
    for i in range(S):
        for j in range(K):
            d = np.arange(max(0, pl - i), min(S, S + pl - i))
            h = np.arange(max(0, pw - j), min(K, K + pw - j))
            co = ((d[:, None] * K + h[None, :]) * 2).ravel()
            ri = (((d[:, None] + i - pl) * K + (h[None, :] + j - pw)) * 2).ravel()
            gk[i, j, 0, 0, 0] = np.sum(dBp[ri, co] - dBp[ri + 1, co + 1])
            gk[i, j, 0, 0, 1] = np.sum(dBp[ri, co + 1] - dBp[ri + 1, co])
    g0 = np.sum(dbp[0::2] - dbp[1::2])
    return gk, np.array([g0, -g0], dtype=dBp.dtype)


# layers_conv2d_vector (--opt 7, dev/py/complex.py:199-255): plain real maps, no complex recombination
def _pack_vec_1xk(kernel, bias):
    """(1,K) 'valid', kernel [1,K,2,1,2F]: rows (k, iq), columns (f, part) with part 0 = channels [0,F), 1 = [F,2F)."""
    k = kernel[0, :, :, 0, :]                                 # [K, 2, 2F]
    K, _, F2 = k.shape
    F = F2 // 2
    Bp = k.reshape(K, 2, 2, F).transpose(0, 1, 3, 2).reshape(2 * K, 2 * F)
    bp = bias.reshape(2, F).T.reshape(2 * F)
    return np.ascontiguousarray(Bp), np.ascontiguousarray(bp)


def _unpack_vec_1xk(dBp, dbp, shape):
    K, F = shape[1], shape[4] // 2
    gk = dBp.reshape(K, 2, F, 2).transpose(0, 1, 3, 2).reshape(shape)
    return gk, dbp.reshape(F, 2).T.reshape(2 * F)


def _pack_vec_toeplitz(kernel, bias, S, K):
    """(S,K) 'same', one filter, kernel [S,K,2,1,2]: re / im = channels 0 / 1 at IQ position 0."""
    n = S * K * 2
    Bp = np.zeros((n, n), dtype=kernel.dtype)
    pl, pw = (S - 1) // 2, (K - 1) // 2
    for i in range(S):
        for j in range(K):
            d = np.arange(max(0, pl - i), min(S, S + pl - i))
            h = np.arange(max(0, pw - j), min(K, K + pw - j))
            co = ((d[:, None] * K + h[None, :]) * 2).ravel()
            ri = (((d[:, None] + i - pl) * K + (h[None, :] + j - pw)) * 2).ravel()
            for iq in range(2):
                for ch in range(2):
                    Bp[ri + iq, co + ch] = kernel[i, j, iq, 0, ch]
    bp = np.zeros(n, dtype=kernel.dtype)
    bp[0::2], bp[1::2] = bias[0], bias[1]
    return Bp, bp


def _unpack_vec_toeplitz(dBp, dbp, S, K):
    gk = np.zeros((S, K, 2, 1, 2), dtype=dBp.dtype)
    pl, pw = (S - 1) // 2, (K - 1) // 2
    for i in range(S):
        for j in range(K):
            d = np.arange(max(0, pl - i), min(S, S + pl - i))
            h = np.arange(max(0, pw - j), min(K, K + pw - j))
            co = ((d[:, None] * K + h[None, :]) * 2).ravel()
            ri = (((d[:, None] + i - pl) * K + (h[None, :] + j - pw)) * 2).ravel()
            for iq in range(2):
                for ch in range(2):
                    gk[i, j, iq, 0, ch] = np.sum(dBp[ri + iq, co + ch])
    return gk, np.array([dbp[0::2].sum(), dbp[1::2].sum()], dtype=dBp.dtype)


# ---------------------------------------------------------------------------------------------
# forward + backward of  ce_mean + REG_COEFF * sum(l2)  w.r.t. the Equalizer variables
# ---------------------------------------------------------------------------------------------
def loss_and_grads(x, bits, w, nbits, nfft=64, cp_len=16, use_cp=True, nfilter=64, dtype=np.float64,
                   normalize=True, reg=True, opt=0):
    """x [B,S,T,2] ('tx_ofdm' feed), bits [B,D,nbits] -> (ce_mean, reg_loss, grads, aux).

    opt 0: equalizer_ofdm.  opt 7: equalizer_separateIQ (dev/py/model.py:1088-1218) -- the same wiring with
    layers_conv2d_vector in place of the complex convs and tanh on all three chain layers (:1140-1162).

    grads: TF variable name -> d total_loss / d var for every 'Equalizer/*' variable, in the
    reference layout.  aux holds intermediate tensors for layer-level checks.
    """
    assert opt in (0, 7)
    vec = opt == 7
    acts = (1, 1, 1) if vec else (0, 0, 1)
    _pack_1xk = _pack_vec_1xk if vec else globals()['_pack_1xk']
    _unpack_1xk = _unpack_vec_1xk if vec else globals()['_unpack_1xk']
    _pack_toeplitz = _pack_vec_toeplitz if vec else globals()['_pack_toeplitz']
    _unpack_toeplitz = _unpack_vec_toeplitz if vec else globals()['_unpack_toeplitz']
    A = lambda n: np.asarray(w[n], dtype=dtype)
    x = np.asarray(x, dtype=dtype)
    B, S, T, _ = x.shape
    K, F = nfft, nfilter
    z = orc.batch_moment_norm(x, dtype)[0] if normalize else x
    # ---- equalizer_ofdm forward (dev/py/model.py:349-462), GEMM form -------------------------
    a0 = orc.layer_norm(z, dtype)
    a0 = a0.reshape(B * S, T * 2) if use_cp else a0[:, :, cp_len:cp_len + K, :].reshape(B * S, K * 2)
    W1, b1 = A('Equalizer/dense/kernel'), A('Equalizer/dense/bias')
    Bp2, bp2 = _pack_1xk(A('Equalizer/conv3d/kernel'), A('Equalizer/conv3d/bias'))
    W3, b3 = A('Equalizer/dense_1/kernel'), A('Equalizer/dense_1/bias')
    W4, b4 = A('Equalizer/dense_2/kernel'), A('Equalizer/dense_2/bias')
    W5, b5 = A('Equalizer/dense_3/kernel'), A('Equalizer/dense_3/bias')
    W6, b6 = A('Equalizer/dense_4/kernel'), A('Equalizer/dense_4/bias')
    Bp7, bp7 = _pack_toeplitz(A('Equalizer/conv3d_1/kernel'), A('Equalizer/conv3d_1/bias'), S, K)
    Bp8, bp8 = _pack_1xk(A('Equalizer/conv3d_2/kernel'), A('Equalizer/conv3d_2/bias'))
    Bp9, bp9 = _pack_1xk(A('Equalizer/conv3d_3/kernel'), A('Equalizer/conv3d_3/bias'))
    W10, b10 = A('Equalizer/dense_5/kernel'), A('Equalizer/dense_5/bias')

    t1 = a0 @ W1 + b1                                   # [BS, 2K]              model.py:370
    f = t1 @ Bp2 + bp2                                  # [BS, 2K] learned DFT  model.py:377-379
    fl = f.reshape(B, S * K * 2)                        # model.py:391
    p = fl @ W3 + b3                                    # model.py:393
    c2 = p @ W4 + b4                                    # model.py:401
    c2 = np.tanh(c2) if acts[0] else c2
    c3 = c2 @ W5 + b5                                   # model.py:407
    c3 = np.tanh(c3) if acts[1] else c3
    c4 = np.tanh(c3 @ W6 + b6)                          # model.py:419
    ch = c4 @ Bp7 + bp7                                 # [B, 2SK] chest        model.py:426
    cr, ci = ch[:, 0::2], ch[:, 1::2]
    fr, fi = fl[:, 0::2], fl[:, 1::2]
    ab = np.sqrt(cr * cr + ci * ci)                     # model.py:431
    nr, ni = cr / ab, -ci / ab                          # model.py:432-433
    er, ei = fr * nr - fi * ni, fr * ni + fi * nr       # model.py:434
    corr = er * er + ei * ei                            # model.py:437 (imag identically 0)
    eqv = np.stack([er, ei], -1).reshape(B * S, 2 * K)
    corrv = np.stack([corr, np.zeros_like(corr)], -1).reshape(B * S, 2 * K)
    eqo = eqv @ Bp9 + bp9                               # model.py:442
    corro = corrv @ Bp8 + bp8                           # model.py:438
    cat = np.concatenate([eqo.reshape(B * S, K, 2), corro.reshape(B * S, K, 2)], -1).reshape(B * S, 4 * K)
    oeq = (cat @ W10 + b10).reshape(B, S, T, 2)         # model.py:457-462
    # ---- frozen ofdm_dense_rx forward (dev/py/model.py:1222-1292) -----------------------------
    Tin = T if use_cp else K
    rin = oeq.reshape(B * S, 2 * T) if use_cp else oeq[:, :, cp_len:, :].reshape(B * S, 2 * K)
    kf = A('fft_like/conv3d/kernel')[0, (Tin - 1) // 2, 0]            # live tap [Tin, 2F]
    bfl = A('fft_like/conv3d/bias')
    BpR, bpR = orc.pack_complex_kernel(kf[:, :F], kf[:, F:], bfl[:F], bfl[F:], dtype=dtype)
    r1 = (rin @ BpR + bpR).reshape(B, S * F * 2)
    Wd, bd = A('demodulation/dense/kernel'), A('demodulation/dense/bias')
    oiq = (r1 @ Wd + bd).reshape(B, -1, 2)                             # [B, D, 2]
    D = oiq.shape[1]
    Wc = A('demodulation/conv2d/kernel').reshape(2, -1)
    bc = A('demodulation/conv2d/bias')
    W1h, b1h = A('demodulation/dense_1/kernel'), A('demodulation/dense_1/bias')
    hpre = oiq @ Wc + bc
    hh = np.maximum(orc.LEAKY_ALPHA * hpre, hpre)
    hcat = np.concatenate([hh, oiq], -1)
    lpre = hcat @ W1h + b1h
    lg = np.maximum(orc.LEAKY_ALPHA * lpre, lpre).reshape(B, D, nbits, 2)
    m = lg.max(-1, keepdims=True)
    e = np.exp(lg - m)
    soft = e / e.sum(-1, keepdims=True)
    # ---- loss (dev/py/ofdmreceiver_np.py:154-162) -----------------------------------------------
    y = np.asarray(bits).astype(np.int64)
    oh = np.stack([1 - y, y], -1).astype(dtype)
    lse = np.log(np.exp(soft).sum(-1, keepdims=True))
    N = B * D * nbits
    ce_mean = float(np.sum(lse[..., 0] - np.sum(soft * oh, -1)) / N)
    # ---- backward ----------------------------------------------------------------------------
    dsoft = (np.exp(soft - lse) - oh) / N                             # softmax(p) - onehot
    dlg = soft * (dsoft - np.sum(dsoft * soft, -1, keepdims=True))    # through the model's softmax
    dlpre = dlg.reshape(B, D, 2 * nbits) * np.where(lpre > 0, 1.0, orc.LEAKY_ALPHA)
    dhcat = dlpre @ W1h.T
    MO = Wc.shape[1]
    dhpre = dhcat[..., :MO] * np.where(hpre > 0, 1.0, orc.LEAKY_ALPHA)
    doiq = dhcat[..., MO:] + dhpre @ Wc.T                              # [B, D, 2]
    dr1 = doiq.reshape(B, 2 * D) @ Wd.T                                # [B, S*F*2]
    drin = dr1.reshape(B * S, 2 * F) @ BpR.T                           # [BS, 2Tin]
    if use_cp:
        doeq = drin
    else:
        doeq = np.zeros((B * S, T, 2), dtype=dtype)
        doeq[:, cp_len:, :] = drin.reshape(B * S, K, 2)
        doeq = doeq.reshape(B * S, 2 * T)
    g = {}
    pre = 'Equalizer/'
    g[pre + 'dense_5/kernel'] = cat.T @ doeq
    g[pre + 'dense_5/bias'] = doeq.sum(0)
    dcat = (doeq @ W10.T).reshape(B * S, K, 4)
    deqo = np.ascontiguousarray(dcat[:, :, 0:2]).reshape(B * S, 2 * K)
    dcorro = np.ascontiguousarray(dcat[:, :, 2:4]).reshape(B * S, 2 * K)
    g[pre + 'conv3d_3/kernel'], g[pre + 'conv3d_3/bias'] = _unpack_1xk(
        eqv.T @ deqo, deqo.sum(0), w['Equalizer/conv3d_3/kernel'].shape)
    g[pre + 'conv3d_2/kernel'], g[pre + 'conv3d_2/bias'] = _unpack_1xk(
        corrv.T @ dcorro, dcorro.sum(0), w['Equalizer/conv3d_2/kernel'].shape)
    deqv = (deqo @ Bp9.T).reshape(B, S * K, 2)
    dcorr = (dcorro @ Bp8.T).reshape(B, S * K, 2)[:, :, 0]
    der = deqv[:, :, 0] + 2 * er * dcorr
    dei = deqv[:, :, 1] + 2 * ei * dcorr
    dfr = der * nr + dei * ni
    dfi = -der * ni + dei * nr
    dnr = der * fr + dei * fi
    dni = -der * fi + dei * fr
    com = (dnr * ci + dni * cr) / (ab ** 3)
    dcr, dci = ci * com, -cr * com
    dch = np.stack([dcr, dci], -1).reshape(B, 2 * S * K)
    g[pre + 'conv3d_1/kernel'], g[pre + 'conv3d_1/bias'] = _unpack_toeplitz(c4.T @ dch, dch.sum(0), S, K)
    dpre4 = (dch @ Bp7.T) * (1 - c4 * c4)
    g[pre + 'dense_4/kernel'] = c3.T @ dpre4
    g[pre + 'dense_4/bias'] = dpre4.sum(0)
    dc3 = dpre4 @ W6.T
    if acts[1]:
        dc3 = dc3 * (1 - c3 * c3)
    g[pre + 'dense_3/kernel'] = c2.T @ dc3
    g[pre + 'dense_3/bias'] = dc3.sum(0)
    dc2 = dc3 @ W5.T
    if acts[0]:
        dc2 = dc2 * (1 - c2 * c2)
    g[pre + 'dense_2/kernel'] = p.T @ dc2
    g[pre + 'dense_2/bias'] = dc2.sum(0)
    dp = dc2 @ W4.T
    g[pre + 'dense_1/kernel'] = fl.T @ dp
    g[pre + 'dense_1/bias'] = dp.sum(0)
    dfl = dp @ W3.T + np.stack([dfr, dfi], -1).reshape(B, 2 * S * K)
    df = dfl.reshape(B * S, 2 * K)
    g[pre + 'conv3d/kernel'], g[pre + 'conv3d/bias'] = _unpack_1xk(
        t1.T @ df, df.sum(0), w['Equalizer/conv3d/kernel'].shape)
    dt1 = df @ Bp2.T
    g[pre + 'dense/kernel'] = a0.T @ dt1
    g[pre + 'dense/bias'] = dt1.sum(0)
    # ---- regularisation: REG_COEFF * l * sum(w^2) over dense kernels + biases ---------------------
    reg_loss = 0.0
    for n in DENSE_NAMES:
        for s in ('/kernel', '/bias'):
            wv = A(pre + n + s)
            reg_loss += L2_L * float(np.sum(wv * wv))
            if reg:
                g[pre + n + s] = g[pre + n + s] + (2.0 * REG_COEFF * L2_L) * wv
    g = {k: np.asarray(v, dtype=dtype).reshape(np.shape(w[k])) for k, v in g.items()}
    aux = dict(soft=soft, oeq=oeq, doeq=doeq.reshape(B, S, T, 2), doiq=doiq, dch=dch, dfl=dfl, chest=ch, dt1=dt1)
    return ce_mean, reg_loss, g, aux


class Adam:
    """tf.train.AdamOptimizer (TF 1.15 formulation, epsilon outside the square root)."""

    def __init__(self, names, weights, dtype=np.float64):
        self.m = {n: np.zeros(np.shape(weights[n]), dtype=dtype) for n in names}
        self.v = {n: np.zeros(np.shape(weights[n]), dtype=dtype) for n in names}
        self.t = 0
        self.dtype = dtype

    def step(self, weights, grads, lr):
        """Updates ``weights`` (dict name -> array) in place for the names this optimiser tracks."""
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - ADAM_B2 ** self.t) / (1.0 - ADAM_B1 ** self.t)
        for n in self.m:
            gr = np.asarray(grads[n], dtype=self.dtype)
            self.m[n] = ADAM_B1 * self.m[n] + (1 - ADAM_B1) * gr
            self.v[n] = ADAM_B2 * self.v[n] + (1 - ADAM_B2) * gr * gr
            upd = lr_t * self.m[n] / (np.sqrt(self.v[n]) + ADAM_EPS)
            weights[n] = (np.asarray(weights[n], dtype=self.dtype) - upd).astype(np.asarray(weights[n]).dtype)


def train_steps(xs, bits, w, nbits, init_learning=1e-3, global_step0=0, **kw):
    """Run len(xs) reference training steps (one minibatch each); returns (weights, [ce_mean...])."""
    w = {k: np.array(v, copy=True) for k, v in w.items()}
    names = trainable_names()
    opt = Adam(names, w)
    losses = []
    for i, (x, b) in enumerate(zip(xs, bits)):
        ce, _, g, _ = loss_and_grads(x, b, w, nbits, **kw)
        opt.step(w, g, learning_rate(init_learning, global_step0 + i))
        losses.append(ce)
    return w, losses


# =============================================================================================
# Training of the basic receiver itself (dev/py/ofdmreceiver_np.py:154-198, loop :211-274)
# =============================================================================================
# graph: tx_ofdm -> batch-moment norm -> ofdm_dense_rx -> softmax;  every variable is trainable:
#   fft_like/conv3d/{kernel,bias}, demodulation/dense/{kernel,bias}, demodulation/conv2d/{kernel,bias},
#   demodulation/dense_1/{kernel,bias}                                   (dev/py/model.py:1246-1288)
# total_loss = ce_mean + berlin * REG_COEFF * sum(REGULARIZATION_LOSSES) + BER_COEFF * ber      (:173)
#   REG_COEFF = 0.0001 (:162); regularisers l2(0.01) sit on demodulation/dense and demodulation/dense_1
#   (kernel and bias; model.py:1269-1272, 1283-1286); conv3d / conv2d have none.
#   berlin / ber come from tf.confusion_matrix(argmax(...)) (:165-169): integer ops, no gradient path -- for the
#   gradient berlin is a per-step constant (the BER of THIS minibatch) and the BER_COEFF term contributes nothing.
# optimiser: AdamOptimizer(exponential_decay(0.001, global_step, 500, 0.98, staircase)).minimize(total_loss)  (:186-189)
# Dead taps of the (1,T) 'same' fft_like kernel only ever multiply zero padding: their gradient is exactly 0 and Adam
# (m = v = 0) leaves them untouched, so only the centre tap [0,(T-1)//2,0] moves.
REG_COEFF_RX = 0.0001
RX_DENSE = ['demodulation/dense', 'demodulation/dense_1']
RX_NAMES = ['fft_like/conv3d', 'demodulation/dense', 'demodulation/conv2d', 'demodulation/dense_1']


def rx_trainable_names():
    return [n + s for n in RX_NAMES for s in ('/kernel', '/bias')]


def rx_loss_and_grads(x, bits, w, nbits, nfft=64, cp_len=16, use_cp=True, nfilter=64, dtype=np.float64,
                      normalize=True, reg=True):
    """x [B,S,T,2], bits [B,D,nbits] -> (ce_mean, reg_loss, berlin, grads, aux) for the basic receiver.

    grads: d total_loss / d var for the eight receiver variables in the reference layouts.
    """
    A = lambda n: np.asarray(w[n], dtype=dtype)
    x = np.asarray(x, dtype=dtype)
    B, S, T, _ = x.shape
    K, F = nfft, nfilter
    z = orc.batch_moment_norm(x, dtype)[0] if normalize else x
    Tin = T if use_cp else K
    rin = z.reshape(B * S, 2 * T) if use_cp else z[:, :, cp_len:, :].reshape(B * S, 2 * K)
    kfull = A('fft_like/conv3d/kernel')
    tap = (Tin - 1) // 2
    kf = kfull[0, tap, 0]                                              # live tap [Tin, 2F]
    bfl = A('fft_like/conv3d/bias')
    BpR, bpR = orc.pack_complex_kernel(kf[:, :F], kf[:, F:], bfl[:F], bfl[F:], dtype=dtype)
    r1 = (rin @ BpR + bpR).reshape(B, S * F * 2)
    Wd, bd = A('demodulation/dense/kernel'), A('demodulation/dense/bias')
    oiq = (r1 @ Wd + bd).reshape(B, -1, 2)
    D = oiq.shape[1]
    Wc = A('demodulation/conv2d/kernel').reshape(2, -1)
    bc = A('demodulation/conv2d/bias')
    W1h, b1h = A('demodulation/dense_1/kernel'), A('demodulation/dense_1/bias')
    MO = Wc.shape[1]
    hpre = oiq @ Wc + bc
    hh = np.maximum(orc.LEAKY_ALPHA * hpre, hpre)
    hcat = np.concatenate([hh, oiq], -1)
    lpre = hcat @ W1h + b1h
    lg = np.maximum(orc.LEAKY_ALPHA * lpre, lpre).reshape(B, D, nbits, 2)
    m = lg.max(-1, keepdims=True)
    e = np.exp(lg - m)
    soft = e / e.sum(-1, keepdims=True)
    y = np.asarray(bits).astype(np.int64)
    oh = np.stack([1 - y, y], -1).astype(dtype)
    lse = np.log(np.exp(soft).sum(-1, keepdims=True))
    N = B * D * nbits
    ce_mean = float(np.sum(lse[..., 0] - np.sum(soft * oh, -1)) / N)
    hard = (soft[..., 1] > soft[..., 0]).astype(np.int64)             # tf.argmax: first index on ties
    berlin = float(np.sum(hard != y)) / N                             # util.ber_tensor: (cm01 + cm10) / sum
    # ---- backward ----------------------------------------------------------------------------
    dsoft = (np.exp(soft - lse) - oh) / N
    dlg = soft * (dsoft - np.sum(dsoft * soft, -1, keepdims=True))
    dlpre = dlg.reshape(B, D, 2 * nbits) * np.where(lpre > 0, 1.0, orc.LEAKY_ALPHA)
    dhcat = dlpre @ W1h.T
    dhpre = dhcat[..., :MO] * np.where(hpre > 0, 1.0, orc.LEAKY_ALPHA)
    doiq = dhcat[..., MO:] + dhpre @ Wc.T
    g = {}
    g['demodulation/dense_1/kernel'] = hcat.reshape(-1, MO + 2).T @ dlpre.reshape(-1, 2 * nbits)
    g['demodulation/dense_1/bias'] = dlpre.reshape(-1, 2 * nbits).sum(0)
    g['demodulation/conv2d/kernel'] = oiq.reshape(-1, 2).T @ dhpre.reshape(-1, MO)
    g['demodulation/conv2d/bias'] = dhpre.reshape(-1, MO).sum(0)
    doiq2 = doiq.reshape(B, 2 * D)
    g['demodulation/dense/kernel'] = r1.T @ doiq2
    g['demodulation/dense/bias'] = doiq2.sum(0)
    dr1 = (doiq2 @ Wd.T).reshape(B * S, 2 * F)
    dBp, dbp = rin.T @ dr1, dr1.sum(0)
    gWa = dBp[0::2, 0::2] - dBp[1::2, 1::2]
    gWb = dBp[0::2, 1::2] - dBp[1::2, 0::2]
    gk = np.zeros(kfull.shape, dtype=dtype)
    gk[0, tap, 0] = np.concatenate([gWa, gWb], axis=1)
    gba = dbp[0::2] - dbp[1::2]
    g['fft_like/conv3d/kernel'] = gk
    g['fft_like/conv3d/bias'] = np.concatenate([gba, -gba])
    reg_loss = 0.0
    for n in RX_DENSE:
        for s in ('/kernel', '/bias'):
            wv = A(n + s)
            reg_loss += L2_L * float(np.sum(wv * wv))
            if reg:
                g[n + s] = g[n + s] + (2.0 * berlin * REG_COEFF_RX * L2_L) * wv
    g = {k: np.asarray(v, dtype=dtype).reshape(np.shape(w[k])) for k, v in g.items()}
    aux = dict(soft=soft, doiq=doiq, dr1=dr1, hard=hard)
    return ce_mean, reg_loss, berlin, g, aux


def rx_train_steps(xs, bits, w, nbits, init_learning=1e-3, global_step0=0, **kw):
    """len(xs) training steps of the basic receiver; returns (weights, [ce_mean...], [berlin...])."""
    w = {k: np.array(v, copy=True) for k, v in w.items()}
    opt = Adam(rx_trainable_names(), w)
    losses, bers = [], []
    for i, (x, b) in enumerate(zip(xs, bits)):
        ce, _, berl, g, _ = rx_loss_and_grads(x, b, w, nbits, **kw)
        opt.step(w, g, learning_rate(init_learning, global_step0 + i))
        losses.append(ce)
        bers.append(berl)
    return w, losses, bers


# =============================================================================================
# Transfer learning of the ablation equalizers (--opt 1, 2, 3, 4, 5) in front of the frozen receiver
# =============================================================================================
# Same loss / optimiser as config 4 (dev/py/ofdmreceiver_np_mp.py:335-347); only the graph inside scope 'Equalizer'
# changes (dev/py/model.py:482-1084, wiring table in dccn_oracle.EQ_SPECS).  Every tf.layers.dense of those functions
# carries l2(0.01) on kernel and bias; the conv3d layers have no regulariser; tf.ifft has no variables.
def variant_trainable_names(opt, prefix='Equalizer/'):
    return [prefix + n + s for _, n in orc.eq_layer_names(opt) for s in ('/kernel', '/bias')]


def _ifft_matrix(K, dtype):
    """tf.ifft over K points as a real [2K, 2K] map on interleaved (re, im): standard complex product, 1/K."""
    k = np.arange(K)
    ang = 2.0 * np.pi * np.outer(k, k) / K
    c, s_ = np.cos(ang) / K, np.sin(ang) / K
    M = np.zeros((2 * K, 2 * K), dtype=dtype)
    M[0::2, 0::2] = c
    M[1::2, 0::2] = -s_
    M[0::2, 1::2] = s_
    M[1::2, 1::2] = c
    return M


def variant_loss_and_grads(x, bits, w, nbits, opt, nfft=64, cp_len=16, use_cp=True, nfilter=64, dtype=np.float64,
                           normalize=True, reg=True):
    """(ce_mean, reg_loss, grads, aux) for the --opt graph: d total_loss / d (every Equalizer/* variable)."""
    sp = orc.EQ_SPECS[opt]
    assert not sp.get('vector', False), 'the layers_conv2d_vector graph (--opt 7) is inference only'
    names = dict(orc.eq_layer_names(opt))
    pre = 'Equalizer/'
    A = lambda n: np.asarray(w[n], dtype=dtype)
    KV = lambda role: A(pre + names[role] + '/kernel')
    BV = lambda role: A(pre + names[role] + '/bias')
    x = np.asarray(x, dtype=dtype)
    B, S, T, _ = x.shape
    K, F = nfft, nfilter
    SK2 = S * K * 2
    z = orc.batch_moment_norm(x, dtype)[0] if normalize else x
    a0 = orc.layer_norm(z, dtype)
    a0 = a0.reshape(B * S, T * 2) if use_cp else a0[:, :, cp_len:cp_len + K, :].reshape(B * S, K * 2)
    # ---- forward, GEMM form ---------------------------------------------------------------------------------------
    W1, b1 = KV('front1'), BV('front1')
    t1 = a0 @ W1 + b1
    if sp['front2'] == 'cconv':
        Bp2, bp2 = _pack_1xk(KV('front2'), BV('front2'))
    else:
        Bp2, bp2 = KV('front2'), BV('front2')
    f = t1 @ Bp2 + bp2                                              # [BS, 2K] inputs_complex
    fl = f.reshape(B, SK2)
    Wp, bpil = KV('pilot'), BV('pilot')
    p = fl @ Wp + bpil
    cin, couts = [p], []
    for i, act in enumerate(sp['chain']):
        y = cin[-1] @ KV('chain%d' % i) + BV('chain%d' % i)
        y = np.tanh(y) if act else y
        couts.append(y)
        cin.append(y)
    if sp['toeplitz']:
        Bp7, bp7 = _pack_toeplitz(KV('toeplitz'), BV('toeplitz'), S, K)
        ch = couts[-1] @ Bp7 + bp7
    else:
        ch = couts[-1]
    cr, ci = ch[:, 0::2], ch[:, 1::2]
    fr, fi = fl[:, 0::2], fl[:, 1::2]
    ab = np.sqrt(cr * cr + ci * ci)
    nr, ni = cr / ab, -ci / ab
    er, ei = fr * nr - fi * ni, fr * ni + fi * nr
    eqv = np.stack([er, ei], -1).reshape(B * S, 2 * K)
    if sp['tail'] == 'dense2':
        Wt1, bt1 = KV('tail1'), BV('tail1')
        mid = eqv @ Wt1 + bt1
    else:
        Wt1 = _ifft_matrix(K, dtype)
        mid = eqv @ Wt1
    Wt2, bt2 = KV('tail2'), BV('tail2')
    oeq = (mid @ Wt2 + bt2).reshape(B, S, T, 2)
    # ---- frozen receiver + loss: reuse the receiver part of the config-4 oracle through its helper ---------------------
    ce_mean, doeq, soft = _rx_loss_and_input_grad(oeq, bits, w, nbits, nfft, cp_len, use_cp, nfilter, dtype)
    # ---- backward --------------------------------------------------------------------------------------------------
    g = {}

    def put(role, gk, gb):
        g[pre + names[role] + '/kernel'] = gk
        g[pre + names[role] + '/bias'] = gb

    put('tail2', mid.T @ doeq, doeq.sum(0))
    dmid = doeq @ Wt2.T
    if sp['tail'] == 'dense2':
        put('tail1', eqv.T @ dmid, dmid.sum(0))
    deqv = (dmid @ Wt1.T).reshape(B, S * K, 2)
    der, dei = deqv[:, :, 0], deqv[:, :, 1]
    dfr = der * nr + dei * ni
    dfi = -der * ni + dei * nr
    dnr = der * fr + dei * fi
    dni = -der * fi + dei * fr
    com = (dnr * ci + dni * cr) / (ab ** 3)
    dch = np.stack([ci * com, -cr * com], -1).reshape(B, SK2)
    if sp['toeplitz']:
        gk, gb = _unpack_toeplitz(couts[-1].T @ dch, dch.sum(0), S, K)
        put('toeplitz', gk, gb)
        dy = dch @ Bp7.T
    else:
        dy = dch
    for i in range(len(sp['chain']) - 1, -1, -1):
        if sp['chain'][i]:
            dy = dy * (1 - couts[i] * couts[i])
        put('chain%d' % i, cin[i].T @ dy, dy.sum(0))
        dy = dy @ KV('chain%d' % i).T
    put('pilot', fl.T @ dy, dy.sum(0))
    dfl = dy @ Wp.T + np.stack([dfr, dfi], -1).reshape(B, SK2)
    df = dfl.reshape(B * S, 2 * K)
    if sp['front2'] == 'cconv':
        gk, gb = _unpack_1xk(t1.T @ df, df.sum(0), w[pre + names['front2'] + '/kernel'].shape)
        put('front2', gk, gb)
    else:
        put('front2', t1.T @ df, df.sum(0))
    dt1 = df @ Bp2.T
    put('front1', a0.T @ dt1, dt1.sum(0))
    # ---- regulariser on every tf.layers.dense (kernel + bias) --------------------------------------------------------
    reg_loss = 0.0
    for role, n in orc.eq_layer_names(opt):
        if not n.startswith('dense'):
            continue
        for s in ('/kernel', '/bias'):
            wv = A(pre + n + s)
            reg_loss += L2_L * float(np.sum(wv * wv))
            if reg:
                g[pre + n + s] = g[pre + n + s] + (2.0 * REG_COEFF * L2_L) * wv
    g = {k: np.asarray(v, dtype=dtype).reshape(np.shape(w[k])) for k, v in g.items()}
    return ce_mean, reg_loss, g, dict(soft=soft, oeq=oeq, doeq=doeq.reshape(B, S, T, 2), dch=dch)


def _rx_loss_and_input_grad(oeq, bits, w, nbits, nfft, cp_len, use_cp, nfilter, dtype):
    """Frozen ofdm_dense_rx + ce_mean on equalised frames oeq [B,S,T,2]: (ce_mean, d ce_mean / d oeq [BS, 2T], soft)."""
    A = lambda n: np.asarray(w[n], dtype=dtype)
    B, S, T, _ = oeq.shape
    K, F = nfft, nfilter
    Tin = T if use_cp else K
    rin = oeq.reshape(B * S, 2 * T) if use_cp else oeq[:, :, cp_len:, :].reshape(B * S, 2 * K)
    kf = A('fft_like/conv3d/kernel')[0, (Tin - 1) // 2, 0]
    bfl = A('fft_like/conv3d/bias')
    BpR, bpR = orc.pack_complex_kernel(kf[:, :F], kf[:, F:], bfl[:F], bfl[F:], dtype=dtype)
    r1 = (rin @ BpR + bpR).reshape(B, S * F * 2)
    Wd, bd = A('demodulation/dense/kernel'), A('demodulation/dense/bias')
    oiq = (r1 @ Wd + bd).reshape(B, -1, 2)
    D = oiq.shape[1]
    Wc = A('demodulation/conv2d/kernel').reshape(2, -1)
    bc = A('demodulation/conv2d/bias')
    W1h, b1h = A('demodulation/dense_1/kernel'), A('demodulation/dense_1/bias')
    hpre = oiq @ Wc + bc
    hh = np.maximum(orc.LEAKY_ALPHA * hpre, hpre)
    lpre = np.concatenate([hh, oiq], -1) @ W1h + b1h
    lg = np.maximum(orc.LEAKY_ALPHA * lpre, lpre).reshape(B, D, nbits, 2)
    e = np.exp(lg - lg.max(-1, keepdims=True))
    soft = e / e.sum(-1, keepdims=True)
    y = np.asarray(bits).astype(np.int64)
    oh = np.stack([1 - y, y], -1).astype(dtype)
    lse = np.log(np.exp(soft).sum(-1, keepdims=True))
    N = B * D * nbits
    ce_mean = float(np.sum(lse[..., 0] - np.sum(soft * oh, -1)) / N)
    dsoft = (np.exp(soft - lse) - oh) / N
    dlg = soft * (dsoft - np.sum(dsoft * soft, -1, keepdims=True))
    dlpre = dlg.reshape(B, D, 2 * nbits) * np.where(lpre > 0, 1.0, orc.LEAKY_ALPHA)
    dhcat = dlpre @ W1h.T
    MO = Wc.shape[1]
    dhpre = dhcat[..., :MO] * np.where(hpre > 0, 1.0, orc.LEAKY_ALPHA)
    doiq = dhcat[..., MO:] + dhpre @ Wc.T
    drin = (doiq.reshape(B, 2 * D) @ Wd.T).reshape(B * S, 2 * F) @ BpR.T
    if use_cp:
        return ce_mean, drin, soft
    doeq = np.zeros((B * S, T, 2), dtype=dtype)
    doeq[:, cp_len:, :] = drin.reshape(B * S, K, 2)
    return ce_mean, doeq.reshape(B * S, 2 * T), soft


def variant_train_steps(xs, bits, w, nbits, opt, init_learning=1e-3, global_step0=0, **kw):
    w = {k: np.array(v, copy=True) for k, v in w.items()}
    optm = Adam(variant_trainable_names(opt), w)
    losses = []
    for i, (x, b) in enumerate(zip(xs, bits)):
        ce, _, g, _ = variant_loss_and_grads(x, b, w, nbits, opt, **kw)
        optm.step(w, g, learning_rate(init_learning, global_step0 + i))
        losses.append(ce)
    return w, losses
