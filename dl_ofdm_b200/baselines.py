"""Expert OFDM receivers as BER comparators for the learned receiver (second half).

The reference compares DCCN with classical receivers computed in MATLAB (`dev/m/OFDM_Benchmark_dev.m`, driven by
`dev/m/script_rayleigh.m:53-69`): remove the cyclic prefix, FFT, estimate the channel on the LTE pilot cells, one-tap
equalise, hard-decide.  MATLAB / Octave are not available here, so the estimators the paper plots are restated in NumPy:

  perfect     the true frequency response (OFDM_Benchmark_dev.m:339-341, eq_idx 1)
  ls_spline   LS at the pilots + MATLAB `griddata(..., 'v4')` = biharmonic-spline interpolation over (subcarrier, symbol)
              (:346-348, eq_idx 2)
  ls_linear   LS at the pilots + `scatteredInterpolant` (linear, Delaunay) over the same grid (:349-352, eq_idx 3)
  lmmse       the "ideal LMMSE" of :353-363 (eq_idx 4): W = Rhh (Rhh + beta/snr I)^-1 with Rhh = H H^H of the TRUE per-symbol
              response, applied to the spline LS estimate; beta = 1, 1, 17/9, 17/9 for BPSK..16-QAM (:222)

Both interpolators are linear in the pilot values for a fixed pilot geometry, so each is ONE [S*K, n_pilot] matrix built at
construction; a batch of frames is then two small matrix products.  This is host NumPy on purpose: the comparators are
not part of the DCCN hot path (the reference runs them in MATLAB on the CPU) -- they only put expert BER curves next to the
GPU sweep's.  Frames are the same ones the receiver sees (made by the library's feeder with injected path gains, so the
true channel is known).  Restatement caveats: MATLAB's RNG, its `awgn(...,'measured')` per-call power and
scatteredInterpolant's extrapolation rule outside the pilots' convex hull (here: the nearest triangle's plane) are not
reproduced bit for bit -- the curves are comparators, not parity targets.
"""
from __future__ import annotations

import numpy as np

from .ofdm import const_map

BETA = {1: 1.0, 2: 1.0, 3: 17.0 / 9.0, 4: 17.0 / 9.0}          # OFDM_Benchmark_dev.m:222


def _green(r):
    """Biharmonic Green's function of griddata 'v4' (Sandwell 1987): r^2 (ln r - 1), 0 at r = 0."""
    out = np.zeros_like(r)
    nz = r > 0
    out[nz] = r[nz] ** 2 * (np.log(r[nz]) - 1.0)
    return out


def spline_matrix(px, py, qx, qy):
    """A with  v(q) = A @ v(p)  for MATLAB griddata(px, py, v, qx, qy, 'v4')."""
    p = np.stack([px, py], 1).astype(np.float64)
    q = np.stack([qx, qy], 1).astype(np.float64)
    gpp = _green(np.linalg.norm(p[:, None] - p[None], axis=-1))
    gqp = _green(np.linalg.norm(q[:, None] - p[None], axis=-1))
    return gqp @ np.linalg.inv(gpp)


def linear_matrix(px, py, qx, qy):
    """A with  v(q) = A @ v(p)  for a linear Delaunay interpolant (scatteredInterpolant, 'linear'); queries outside the
    convex hull use the plane of the triangle with the least-negative barycentric coordinate."""
    from scipy.spatial import Delaunay
    p = np.stack([px, py], 1).astype(np.float64)
    q = np.stack([qx, qy], 1).astype(np.float64)
    tri = Delaunay(p)
    T = tri.transform                                            # [nsimplex, 3, 2]
    A = np.zeros((len(q), len(p)))
    simp = tri.find_simplex(q)
    for i, (pt, sidx) in enumerate(zip(q, simp)):
        if sidx < 0:
            b2 = np.einsum('sij,sj->si', T[:, :2], pt[None] - T[:, 2])
            bary = np.concatenate([b2, 1.0 - b2.sum(1, keepdims=True)], 1)
            sidx = int(np.argmax(bary.min(axis=1)))
            b = bary[sidx]
        else:
            b2 = T[sidx, :2] @ (pt - T[sidx, 2])
            b = np.append(b2, 1.0 - b2.sum())
        A[i, tri.simplices[sidx]] = b
    return A


class ClassicReceiver:
    """CP removal + FFT + pilot-based channel estimate + one-tap equaliser + hard QAM decision for the frames of `ofdmobj`."""

    def __init__(self, ofdmobj, nbits):
        self.o, self.nb = ofdmobj, nbits
        S, K = ofdmobj.nSymbol, ofdmobj.K
        psc = np.asarray(ofdmobj.pilotSc)
        # MATLAB grids: gt = symbol 1..S, gf = subcarrier 1..N (OFDM_Benchmark_dev.m:130-172)
        pf, pt = (psc % K) + 1.0, (psc // K) + 1.0
        gt, gf = np.meshgrid(np.arange(1, S + 1.0), np.arange(1, K + 1.0))       # [K, S]
        qf, qt = gf.T.reshape(-1), gt.T.reshape(-1)                             # symbol-major like the frame (s*K + k)
        self.A = {'ls_spline': spline_matrix(pf, pt, qf, qt), 'ls_linear': linear_matrix(pf, pt, qf, qt)}
        self.const = const_map(nbits).astype(np.complex128)
        self.pilot = complex(ofdmobj.pilotValue)

    def true_response(self, g):
        """Frequency response the FFT window sees for the centred 'same' FIR g [B, M] (radio.py:436): tap j sits at delay
        j - off, off = (M-1) - M//2."""
        K = self.o.K
        M = g.shape[1]
        off = (M - 1) - (M >> 1)
        d = np.arange(M) - off
        E = np.exp(-2j * np.pi * np.outer(d, np.arange(K)) / K)                   # [M, K]
        return g @ E                                                              # [B, K], the same in every symbol (static channel)

    def decide(self, x, estimator, snr_db=None, g=None):
        """x float [B,S,T,2] received frames -> hard bits uint8 [B, D, nbits]."""
        o, K, S = self.o, self.o.K, self.o.nSymbol
        B = x.shape[0]
        y = x[..., 0].astype(np.float64) + 1j * x[..., 1].astype(np.float64)
        Y = np.fft.fft(y[:, :, o.CP:], axis=-1).reshape(B, S * K)                 # remove CP, to the frequency domain
        if estimator == 'perfect':
            H = np.tile(self.true_response(g), (1, S))
            # the received frames were scaled by 1/sqrt(mean power) of the batch (AWGN_channel_np): a real gain LS would absorb;
            # estimate it from the pilots so that 'perfect' knows the channel, not the AGC
            ph = Y[:, o.pilotSc] / self.pilot
            gain = (np.abs(ph) ** 2).sum() / np.real(ph * np.conj(H[:, o.pilotSc])).sum()
            H = H * gain
        else:
            hls = Y[:, o.pilotSc] / self.pilot                                    # LS at the pilots (:342)
            if estimator in ('ls_spline', 'ls_linear'):
                H = hls @ self.A[estimator].T
            elif estimator == 'lmmse':
                hs = (hls @ self.A['ls_spline'].T).reshape(B, S, K)
                Ht = self.true_response(g)                                        # [B, K]
                ph = hls
                gain = (np.abs(ph) ** 2).sum() / np.real(ph * np.conj(np.tile(Ht, (1, S))[:, o.pilotSc])).sum()
                Ht = Ht * gain
                lam = BETA[self.nb] / (10.0 ** (float(snr_db) / 10.0))
                # Rhh = h h^H (rank one): W = h h^H / (|h|^2 + lam)  (Sherman-Morrison of :358-359)
                n2 = (np.abs(Ht) ** 2).sum(1, keepdims=True)
                proj = np.einsum('bk,bsk->bs', np.conj(Ht), hs)                   # h^H hs per symbol
                H = (Ht[:, None, :] * (proj / (n2 + lam))[:, :, None]).reshape(B, S * K)
            else:
                raise ValueError(estimator)
        eq = Y[:, o.dataSc] / H[:, o.dataSc]
        idx = np.abs(eq[..., None] - self.const[None, None, :]).argmin(-1)        # nearest constellation point
        shifts = np.arange(self.nb - 1, -1, -1)
        return ((idx[..., None] >> shifts) & 1).astype(np.uint8)

    def ber(self, x, bits, estimator, snr_db=None, g=None):
        hard = self.decide(x, estimator, snr_db, g)
        return float((hard != np.asarray(bits)).mean())
