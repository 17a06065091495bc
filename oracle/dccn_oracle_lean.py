"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a fast form of the oracle for full-size parity runs.

`dccn_oracle.py` restates the reference op by op -- every `layers_conv2d_complex` as a literal tap loop -- and evaluates
about 100 frames/s, which is fine for the seeded small cases but not for BASELINE config 3 (65 536 frames per batch).
This module evaluates the SAME functions as dense matrix products (the packing of SURVEY.md Appendix D; the (7,64) 'same'
conv as its Toeplitz matrix, built from the literal definition) in float64 BLAS, a few thousand frames/s per core.
It is not an independent restatement: tests/test_oracle_golden.py::test_lean_oracle_equals_literal pins it to
`dccn_oracle.equalized_receiver` / `basic_receiver` (fp64, <= 1e-11) on trained and on seeded weights, and the full-size
GPU parity test then uses it on every frame of the batch.

Reference lines: dev/py/model.py:349-478 (equalizer_ofdm), :1222-1292 (ofdm_dense_rx), dev/py/complex.py:140-196,
dev/py/ofdmreceiver_np.py:128-129 (batch-moment norm).
"""
import numpy as np

from . import dccn_oracle as orc


def _packed_1xK(kernel, bias, dtype):
    """(1,K) 'valid' complex conv with F filters over a width-K axis == dense [2K -> 2F]."""
    k = np.asarray(kernel, dtype=np.float64)
    K, F = k.shape[1], k.shape[4] // 2
    Wa, Wb = k[0, :, 0, 0, :F], k[0, :, 0, 0, F:]
    b = np.asarray(bias, dtype=np.float64)
    Bp, bp = orc.pack_complex_kernel(Wa, Wb, b[:F], b[F:], np.float64)
    return Bp.astype(dtype), bp.astype(dtype)


def _toeplitz(kernel, bias, S, K, dtype):
    """(S,K) 'same' complex conv, one filter (model.py:426): re[d,h] = sum xr[d+i-pl, h+j-pw] Wa[i,j] - xi[.] Wb[i,j] + (b0-b1),
    im[d,h] = sum xr[.] Wb[i,j] - xi[.] Wa[i,j] + (b1-b0)  as a [2SK, 2SK] matrix on (s, k, iq)-flattened rows."""
    k = np.asarray(kernel, dtype=np.float64)
    assert k.shape == (S, K, 1, 1, 2), k.shape
    Wa, Wb = k[:, :, 0, 0, 0], k[:, :, 0, 0, 1]
    pl, pw = (S - 1) // 2, (K - 1) // 2
    n = 2 * S * K
    M = np.zeros((n, n), dtype=np.float64)
    for d in range(S):
        for i in range(S):
            di = d + i - pl
            if di < 0 or di >= S:
                continue
            for h in range(K):
                j0, j1 = max(0, pw - h), min(K, K + pw - h)          # 0 <= h + j - pw < K
                j = np.arange(j0, j1)
                rows = (di * K + (h + j - pw)) * 2
                col = (d * K + h) * 2
                M[rows, col] = Wa[i, j]
                M[rows, col + 1] = Wb[i, j]
                M[rows + 1, col] = -Wb[i, j]
                M[rows + 1, col + 1] = -Wa[i, j]
    b = np.asarray(bias, dtype=np.float64)
    bp = np.tile(np.array([b[0] - b[1], b[1] - b[0]]), S * K)
    return M.astype(dtype), bp.astype(dtype)


class LeanModel:
    """[equalizer_ofdm ->] ofdm_dense_rx on NORMALISED input z [B,S,T,2] as dense products in `dtype`."""

    def __init__(self, w, nbits, nfft=64, cp_len=16, nsymbol=7, nfilter=64, use_cp=True, equalizer=True, head='dev',
                 dtype=np.float64):
        self.nb, self.K, self.CP, self.S, self.F, self.use_cp, self.eq, self.head = \
            nbits, nfft, cp_len, nsymbol, nfilter, use_cp, equalizer, head
        self.dtype = dtype
        g = lambda n: np.asarray(w[n], dtype=dtype)                                       # noqa: E731
        T = nfft + cp_len
        Tin = T if use_cp else nfft
        k = np.asarray(w['fft_like/conv3d/kernel'], dtype=np.float64)                     # [1,Tin,1,Tin,2F]
        tap = (Tin - 1) // 2                                                              # 'same' over a width-1 axis
        b = np.asarray(w['fft_like/conv3d/bias'], dtype=np.float64)
        Bp, bp = orc.pack_complex_kernel(k[0, tap, 0, :, :nfilter], k[0, tap, 0, :, nfilter:], b[:nfilter], b[nfilter:])
        self.r1 = (Bp.astype(dtype), bp.astype(dtype))
        self.r2 = (g('demodulation/dense/kernel'), g('demodulation/dense/bias'))
        self.Wc, self.bc = g('demodulation/conv2d/kernel').reshape(2, -1), g('demodulation/conv2d/bias')
        if head == 'v1':
            k1 = g('demodulation/conv2d_1/kernel')
            self.Wc1, self.bc1 = k1.reshape(k1.shape[2], k1.shape[3]), g('demodulation/conv2d_1/bias')
        self.W1, self.b1 = g('demodulation/dense_1/kernel'), g('demodulation/dense_1/bias')
        if equalizer:
            e = 'Equalizer/'
            self.g1 = (g(e + 'dense/kernel'), g(e + 'dense/bias'))
            self.g2 = _packed_1xK(w[e + 'conv3d/kernel'], w[e + 'conv3d/bias'], dtype)
            self.g3 = (g(e + 'dense_1/kernel'), g(e + 'dense_1/bias'))
            self.g4 = (g(e + 'dense_2/kernel'), g(e + 'dense_2/bias'))
            self.g5 = (g(e + 'dense_3/kernel'), g(e + 'dense_3/bias'))
            self.g6 = (g(e + 'dense_4/kernel'), g(e + 'dense_4/bias'))
            self.g7 = _toeplitz(w[e + 'conv3d_1/kernel'], w[e + 'conv3d_1/bias'], nsymbol, nfft, dtype)
            self.g8 = _packed_1xK(w[e + 'conv3d_2/kernel'], w[e + 'conv3d_2/bias'], dtype)
            self.g9 = _packed_1xK(w[e + 'conv3d_3/kernel'], w[e + 'conv3d_3/bias'], dtype)
            self.g10 = (g(e + 'dense_5/kernel'), g(e + 'dense_5/bias'))

    def equalizer(self, z):
        """model.py:349-478 -> (equalized [B,S,T,2], chest complex [B,S,K])."""
        dt = self.dtype
        B, S, T, _ = z.shape
        K = self.K
        x = orc.layer_norm(z, dt)                                                          # :363
        x = x.reshape(B, S, T * 2) if self.use_cp else x[:, :, self.CP:self.CP + K, :].reshape(B, S, K * 2)
        t1 = x @ self.g1[0] + self.g1[1]                                                   # :370
        f = t1 @ self.g2[0] + self.g2[1]                                                   # :377-379  [B,S,2K] (k, iq)
        flat = f.reshape(B, S * K * 2)                                                     # :391
        c = flat @ self.g3[0] + self.g3[1]                                                 # :393
        c = c @ self.g4[0] + self.g4[1]                                                    # :401
        c = c @ self.g5[0] + self.g5[1]                                                    # :407
        c = np.tanh(c @ self.g6[0] + self.g6[1])                                           # :419
        ch = (c @ self.g7[0] + self.g7[1]).reshape(B, S, K, 2)                             # :426
        fc = f.reshape(B, S, K, 2)
        inputs_c = fc[..., 0] + 1j * fc[..., 1]
        chest = ch[..., 0] + 1j * ch[..., 1]                                               # :428
        ab = np.abs(chest)                                                                 # :431
        conj_n = np.real(chest) / ab - 1j * (np.imag(chest) / ab)                          # :432-433 (no eps)
        eq = inputs_c * conj_n                                                             # :434
        corr = eq * np.conj(eq)                                                            # :437
        eq2 = np.stack([eq.real, eq.imag], -1).reshape(B, S, 2 * K).astype(dt)
        corr2 = np.stack([corr.real, corr.imag], -1).reshape(B, S, 2 * K).astype(dt)
        corr_o = (corr2 @ self.g8[0] + self.g8[1]).reshape(B, S, K, 2)                     # :438-440
        eq_o = (eq2 @ self.g9[0] + self.g9[1]).reshape(B, S, K, 2)                         # :442-448
        cat = np.concatenate([eq_o, corr_o], -1).reshape(B, S, 4 * K)                      # :455-456
        out = cat @ self.g10[0] + self.g10[1]                                              # :457
        return out.reshape(B, S, T, 2), chest

    def receiver(self, z):
        """model.py:1222-1292 -> softmax [B,D,nb,2]."""
        B, S, T, _ = z.shape
        x = z if self.use_cp else z[:, :, self.CP:, :]
        Tin = x.shape[2]
        fft = x.reshape(B, S, Tin * 2) @ self.r1[0] + self.r1[1]                           # :1246-1264
        out_iq = (fft.reshape(B, S * self.F * 2) @ self.r2[0] + self.r2[1]).reshape(B, -1, 2)   # :1268-1275
        h = out_iq @ self.Wc + self.bc                                                     # :1278
        if self.head == 'v1':
            h = h @ self.Wc1 + self.bc1
        h = orc._leaky(h)
        cat = np.concatenate([h, out_iq], -1)                                              # :1282
        logits = orc._leaky(cat @ self.W1 + self.b1).reshape(B, -1, self.nb, 2)            # :1283-1290
        return orc._softmax2(logits)                                                       # :1291

    def forward(self, z):
        z = np.asarray(z, dtype=self.dtype)
        if not self.eq:
            return self.receiver(z), None, None
        eq, chest = self.equalizer(z)
        return self.receiver(eq.astype(self.dtype)), eq, chest


def batch_norm_with(x, mean, inv, dtype=np.float64):
    """a2 with GIVEN batch statistics (ofdmreceiver_np.py:128-129): chunks of a big batch share the whole batch's
    moments.  TF evaluates x*inv + (-mean*inv), then / sqrt(2)."""
    x = np.asarray(x, dtype=dtype)
    return (x * inv.astype(dtype) + (-mean.astype(dtype) * inv.astype(dtype))) / dtype(orc.SQRT2)
