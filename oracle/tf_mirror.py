"""Op-for-op torch-CPU mirror of the reference TF-1 graph  --  TEST / BASELINE ONLY.

Part of ``oracle/`` (see dccn_oracle.py header for who may import this).  It
reproduces the *cost structure* of the reference path, including the waste the
reference inherits from emulating complex layers with ``tf.layers.conv3d``:

  * ``layers_conv2d_complex`` (dev/py/complex.py:168-192) is run as a real,
    zero-padded ``conv3d`` with 2F filters over [B, L, W, IQ, C] followed by the
    [...,2,2F] -> [...,4,F] reshape and the c0-c3 / c1-c2 recombination;
  * the 'same' (1,K) learned-DFT layer therefore convolves an 80-wide kernel
    over a width-1 axis (79/80 of the MACs multiply padding zeros);
  * fp32 throughout, intra-op threading = torch.set_num_threads().

It is (a) an independent check of the NumPy restatement in dccn_oracle.py and
(b) the "restated reference" CPU arm of bench.py (``--impl reference`` and
``cpu_baseline``), because TensorFlow 1.x itself cannot run on this image.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as Fn

LEAKY_ALPHA = 0.2
BN_EPS = 1e-9
LN_EPS = 1e-12


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)


def conv2d_complex_tf(x, kernel, bias, padding):
    """complex.py:140-196 with torch ops.  x [B,L,W,C,2] -> [B,L',W',F,2]."""
    B, L, W, C, _ = x.shape
    kl, kw, _, _, F2 = kernel.shape
    F = F2 // 2
    conv = x.permute(0, 1, 2, 4, 3)                              # [B,L,W,IQ,C]   :168
    # torch conv3d wants [B, C, D, H, W]; TF kernel [kd,kh,kw,Cin,Cout]
    inp = conv.permute(0, 4, 1, 2, 3)                            # [B,C,L,W,IQ]
    wt = kernel.permute(4, 3, 0, 1, 2)                           # [2F,C,kl,kw,1]
    if padding == 'same':
        pl, pw = (kl - 1) // 2, (kw - 1) // 2
        inp = Fn.pad(inp, (0, 0, pw, kw - 1 - pw, pl, kl - 1 - pl))
    out = Fn.conv3d(inp, wt, bias)                               # [B,2F,L',W',IQ]  :183
    out = out.permute(0, 2, 3, 4, 1)                             # [B,L',W',IQ,2F]
    Lo, Wo = out.shape[1], out.shape[2]
    c4 = out.reshape(B, Lo, Wo, 4, F)                            # :185
    re = c4[:, :, :, 0] - c4[:, :, :, 3]                         # :187
    im = c4[:, :, :, 1] - c4[:, :, :, 2]                         # :188
    o = torch.stack([re, im], dim=3)                             # [B,L',W',2,F]   :191
    return o.permute(0, 1, 2, 4, 3)                              # [B,L',W',F,2]   :192


class TFMirror:
    """Holds fp32 torch copies of the weights and runs the reference dataflow."""

    def __init__(self, w, nbits, nfft=64, cp_len=16, use_cp=True, head='dev', nfilter=64,
                 equalizer=False, dtype=torch.float32):
        """dtype float32 = the reference's arithmetic (the CPU baseline); float64 makes the mirror an independent
        high-precision check of the NumPy restatement (tests/test_oracle_golden.py)."""
        self.dtype = dtype
        self.w = {k: _t(v, dtype) for k, v in w.items()}
        self.nbits, self.K, self.CP = nbits, nfft, cp_len
        self.use_cp, self.head, self.F, self.eq = use_cp, head, nfilter, equalizer

    # dev/py/ofdmreceiver_np.py:128-129
    def norm(self, x):
        mean = x.mean(dim=0)
        var = ((x - mean) ** 2).mean(dim=0)
        inv = torch.rsqrt(var + BN_EPS)
        return (x * inv + (-mean * inv)) / float(np.sqrt(2.0))

    # dev/py/model.py:1222-1292
    def dense_rx(self, z):
        w = self.w
        B, S, T, _ = z.shape
        out = z if self.use_cp else z[:, :, self.CP:, :]
        K = out.shape[2]
        conv = out.reshape(B, S, 1, K, 2)
        fft = conv2d_complex_tf(conv, w['fft_like/conv3d/kernel'], w['fft_like/conv3d/bias'], 'same')
        flat = fft.reshape(B, S * self.F * 2)
        out_iq = (flat @ w['demodulation/dense/kernel'] + w['demodulation/dense/bias']).reshape(B, -1, 2)
        h = out_iq @ w['demodulation/conv2d/kernel'].reshape(2, -1) + w['demodulation/conv2d/bias']
        if self.head == 'v1':
            k1 = w['demodulation/conv2d_1/kernel']
            h = h @ k1.reshape(k1.shape[2], k1.shape[3]) + w['demodulation/conv2d_1/bias']
        h = torch.maximum(LEAKY_ALPHA * h, h)
        cat = torch.cat([h, out_iq], dim=-1)
        lg = cat @ w['demodulation/dense_1/kernel'] + w['demodulation/dense_1/bias']
        lg = torch.maximum(LEAKY_ALPHA * lg, lg).reshape(B, -1, self.nbits, 2)
        return torch.softmax(lg, dim=-1)

    # dev/py/model.py:349-478
    def equalizer(self, z):
        w = {k[len('Equalizer/'):]: v for k, v in self.w.items() if k.startswith('Equalizer/')}
        B, S, T, _ = z.shape
        K = self.K
        flat = z.reshape(B, -1)
        mean = flat.mean(dim=1, keepdim=True)
        var = ((flat - mean) ** 2).mean(dim=1, keepdim=True)
        inv = torch.rsqrt(var + LN_EPS)
        chest = (flat * inv + (-mean * inv)).reshape(B, S, T, 2)
        if not self.use_cp:
            chest = chest[:, :, self.CP:self.CP + K, :].reshape(B, S, K * 2)
        else:
            chest = chest.reshape(B, S, T * 2)
        t1 = (chest @ w['dense/kernel'] + w['dense/bias']).reshape(B, S, K, 1, 2)
        f = conv2d_complex_tf(t1, w['conv3d/kernel'], w['conv3d/bias'], 'valid').permute(0, 1, 3, 2, 4)
        inr, ini = f[..., 0], f[..., 1]                          # [B,S,K,1]
        c = f.reshape(B, S * K * 2)
        c = c @ w['dense_1/kernel'] + w['dense_1/bias']
        c = c @ w['dense_2/kernel'] + w['dense_2/bias']
        c = c @ w['dense_3/kernel'] + w['dense_3/bias']
        c = torch.tanh(c @ w['dense_4/kernel'] + w['dense_4/bias'])
        c5 = conv2d_complex_tf(c.reshape(B, S, K, 1, 2), w['conv3d_1/kernel'], w['conv3d_1/bias'], 'same')
        cr, ci = c5[..., 0], c5[..., 1]
        ab = torch.sqrt(cr * cr + ci * ci)
        nr, ni = cr / ab, (-ci) / ab
        er = inr * nr - ini * ni
        ei = inr * ni + ini * nr
        corr = torch.stack([er * er + ei * ei, torch.zeros_like(er)], dim=-1)
        corr_o = conv2d_complex_tf(corr, w['conv3d_2/kernel'], w['conv3d_2/bias'], 'valid')
        corr_o = corr_o.permute(0, 1, 3, 2, 4)[:, :, :, 0, :]
        eq5 = torch.stack([er, ei], dim=-1)
        eq_o = conv2d_complex_tf(eq5, w['conv3d_3/kernel'], w['conv3d_3/bias'], 'valid')
        eq_o = eq_o.permute(0, 1, 3, 2, 4)[:, :, :, 0, :]
        cat = torch.cat([eq_o, corr_o], dim=-1).reshape(B, S, K * 4)
        out = cat @ w['dense_5/kernel'] + w['dense_5/bias']
        return out.reshape(B, S, T, 2)

    @torch.no_grad()
    def forward(self, x):
        z = self.norm(_t(x, self.dtype))
        if self.eq:
            z = self.equalizer(z)
        return self.dense_rx(z)
