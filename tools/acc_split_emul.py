"""CPU emulation (no GPU): operand-split schemes for an fp32-class tensor-core GEMM, measured on the shipped v1
16-QAM checkpoint (same frames / metric as tools/acc_probe.py: |soft - soft_fp64|, hard-bit flips).

Each scheme represents an fp32 operand v as hi + lo and evaluates  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (lo*lo dropped);
the products of two <= 11-bit significands are exact in the tensor core's fp32 accumulator, so the scheme's own error
is the REPRESENTATION error of hi + lo -- emulated here with the three matmuls in float64.  What the emulation does
not contain is the accumulator's truncating adds (measured on the GPU: profiles/accuracy_r1.txt, `kc`).

  tf32x3        : hi = rna_tf32(v), lo = rna_tf32(v - hi)            (the library's DCCN_PREC_PARITY today)
  fp16x3        : hi = fp16(v), lo = fp16(v - hi), no scaling        (kind::f16 runs at twice the kind::tf32 rate)
  fp16x3 scaled : weights pre-multiplied by a power of two so that max|W| sits in [2^13, 2^14) (exact; undone in the
                  epilogue), activations as they are
Usage: python tools/acc_split_emul.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import v1_weights, GOLDEN          # noqa: E402
from oracle import dccn_oracle as orc            # noqa: E402
from oracle.v1_recipe import v1_frames           # noqa: E402


def rna_tf32(v):
    b = np.asarray(v, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)


def split_tf32(v):
    v = np.asarray(v, np.float32)
    hi = rna_tf32(v)
    return hi.astype(np.float64), rna_tf32(v - hi).astype(np.float64)


def split_fp16(v, scale=1.0):
    v = np.asarray(v, np.float32) * np.float32(scale)
    hi = v.astype(np.float16)
    assert np.isfinite(hi).all(), 'fp16 overflow'
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64) / scale, lo.astype(np.float64) / scale


def split_none(v):
    v = np.asarray(v, np.float32).astype(np.float64)
    return v, np.zeros_like(v)


def gemm3(a, b, split_a, split_b):
    ah, al = split_a(a)
    bh, bl = split_b(b)
    return ah @ bh + ah @ bl + al @ bh


def receiver(x, w, nb, split_a, split_b_of):
    """basic receiver with the two big contractions through gemm3, everything else in float32 like the GPU path."""
    z = orc.batch_moment_norm(x, dtype=np.float64)[0].astype(np.float32)
    B, S, T, _ = z.shape
    # fft_like: centre tap, packed [[a,b],[-b,-a]] (oracle/dccn_oracle.py pack_complex_kernel)
    k = np.asarray(w['fft_like/conv3d/kernel'], np.float64)[0, (T - 1) // 2, 0]
    F = k.shape[1] // 2
    bias = np.asarray(w['fft_like/conv3d/bias'], np.float64)
    Bp, bp = orc.pack_complex_kernel(k[:, :F], k[:, F:], bias[:F], bias[F:])
    y = gemm3(z.reshape(B * S, T * 2), Bp.astype(np.float32), split_a, split_b_of(Bp)) + bp
    flat = y.astype(np.float32).reshape(B, S * F * 2)
    Wd = np.asarray(w['demodulation/dense/kernel'], np.float32)
    out_iq = gemm3(flat, Wd, split_a, split_b_of(Wd)) + np.asarray(w['demodulation/dense/bias'], np.float64)
    return out_iq.astype(np.float32)


def head_from_out_iq(out_iq, w, nb):
    """the rest of the v1 head in float64 on the given out_iq (isolates the GEMM error)."""
    B = out_iq.shape[0]
    o = out_iq.astype(np.float64).reshape(B, -1, 2)
    Wc = np.asarray(w['demodulation/conv2d/kernel'], np.float64).reshape(2, -1)
    h = o @ Wc + np.asarray(w['demodulation/conv2d/bias'], np.float64)
    Wc1 = np.asarray(w['demodulation/conv2d_1/kernel'], np.float64)
    h = h @ Wc1.reshape(Wc1.shape[2], Wc1.shape[3]) + np.asarray(w['demodulation/conv2d_1/bias'], np.float64)
    h = np.where(h > 0, h, 0.2 * h)
    cat = np.concatenate([h, o], -1)
    lg = cat @ np.asarray(w['demodulation/dense_1/kernel'], np.float64) + np.asarray(w['demodulation/dense_1/bias'], np.float64)
    lg = np.where(lg > 0, lg, 0.2 * lg).reshape(B, -1, nb, 2)
    e = np.exp(lg - lg.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def run(n_frames=700, verbose=True):
    nb = 4
    w = v1_weights(np.load(os.path.join(GOLDEN, 'v1_4mod_cpTrue.npz')))
    x, _ = v1_frames(nb, 10, n_frames)
    ref = orc.basic_receiver(x, w, nb, 16, head='v1', dtype=np.float64)
    hard_ref = ref[..., 1] > ref[..., 0]

    def wscale(W):
        m = float(np.abs(W).max())
        return 2.0 ** (13 - int(np.floor(np.log2(m))))

    schemes = [
        ('fp32 operands (no split)', split_none, lambda W: split_none),
        ('tf32x3', split_tf32, lambda W: split_tf32),
        ('fp16x3 unscaled', split_fp16, lambda W: split_fp16),
        ('fp16x3, weights x 2^k', split_fp16, lambda W: (lambda v, s=wscale(W): split_fp16(v, s))),
    ]
    out = {}
    if verbose:
        print('# tools/acc_split_emul.py (CPU emulation): shipped v1 checkpoint OFDM_Dense3_4mod_snr12_cpTrue, %d frames @ 10 dB' % n_frames)
        print('# (%d bit decisions), |soft - soft_fp64oracle|; exact products, float64 accumulation (operand error only)' % hard_ref.size)
    for name, sa, sb in schemes:
        soft = head_from_out_iq(receiver(x, w, nb, sa, sb), w, nb)
        e = np.abs(soft - ref)
        flips = int(((soft[..., 1] > soft[..., 0]) != hard_ref).sum())
        out[name] = (float(np.quantile(e, .999)), float(e.max()), flips)
        if verbose:
            print('%-26s: p99.9 %.3g max %.3g flips %d' % ((name,) + out[name]))
    if verbose:
        a = orc.batch_moment_norm(x, dtype=np.float32)[0]
        print('# operand ranges: max|z| = %.3g, max|W fft_like| = %.3g, max|W dense| = %.3g (fp16 max 65504, min normal 6.1e-5)'
              % (np.abs(a).max(), np.abs(w['fft_like/conv3d/kernel']).max(), np.abs(w['demodulation/dense/kernel']).max()))
    return out


if __name__ == '__main__':
    run()
