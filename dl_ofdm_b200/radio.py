"""Channel emulators on the GPU with the reference's ``radio.py`` surface.

``rayleigh_chan_lte(FLAGS, sample_rate, mobile, mix).run(iq)`` and
``AWGN_channel_np(x, SNR)`` keep the reference signatures (dev/py/radio.py:277-526)
but run as CUDA kernels (``dccn_chan_fir_awgn``): per-frame CN(0,1) path gains,
``g = (z*ch_coeff) @ alpha``, centred 'same' FIR with zero history, then batch-power
normalisation + AWGN.  Random numbers come from a Philox counter stream (seeded, so
sweeps are reproducible -- the reference seeds NumPy from the wall clock), or can be
injected for parity tests.  Only the static (non-Doppler) branch is implemented;
``mobile=True`` raises.
"""
from __future__ import annotations

import os

import numpy as np
import torch

_TAPS = {   # LTE tapped-delay-line profiles (delay ns, relative power dB), dev/py/radio.py:340-366
    'etu': ([0, 50, 120, 200, 230, 500, 1600, 2300, 5000], [-1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -3.0, -5.0, -7.0]),
    'epa': ([0, 30, 70, 90, 110, 190, 410], [0.0, -1.0, -2.0, -3.0, -8.0, -17.2, -20.8]),
    'eva': ([0, 30, 150, 310, 370, 710, 1090, 1730, 2510], [0.0, -1.5, -1.4, -3.6, -0.6, -9.1, -7.0, -12.0, -16.9]),
    'custom': ([0, 70, 200, 230, 500, 1600, 2700, 3000], [0.0, -1.4, -1.4, -1.0, -3.0, -9.1, -15.0, -19.0]),
    'flat': ([0], [0.0]),
}
_ALPHA = None


def _alpha(chan):
    global _ALPHA
    if chan == 'flat':
        return np.ones((1, 1), dtype=np.float64)
    if _ALPHA is None:
        _ALPHA = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'lte_alpha.npz'))
    return _ALPHA[chan]


def channel_profile(chan):
    """-> (ch_coeff [n_taps], alpha [n_taps, n_fir]).  ch_coeff = linear power / sqrt(sum power)
    used as the path AMPLITUDE, a quirk of the reference that is kept (radio.py:367-371)."""
    chan = chan.lower()
    if chan not in _TAPS:
        chan = 'flat'                      # radio.py:361: every other name is the single-tap channel
    pw = 10.0 ** (np.asarray(_TAPS[chan][1], dtype=np.float64) / 10.0)
    return pw / np.sqrt(pw.sum()), _alpha(chan)


class rayleigh_chan_lte:
    def __init__(self, FLAGS, sample_rate=0.96e6, mobile=False, mix=False, engine=None, seed=0):
        if mobile:
            raise NotImplementedError('Doppler (mobile=True) fading is not implemented on the GPU path yet')
        self.nSymbol = FLAGS.nsymbol
        self.chan = FLAGS.channel.lower()
        self.sample_rate = sample_rate
        self.nfft = FLAGS.nfft
        self.engine = engine
        self.seed = seed
        self._calls = 0
        if self.chan in ('mixrayleigh', 'mixall'):
            raise NotImplementedError("channel '%s' cycles profiles per frame; use one profile per call" % self.chan)
        self.ch_coeff, self.alpha_matrix = (None, None) if self.chan == 'awgn' else channel_profile(self.chan)
        self.n_taps = 0 if self.ch_coeff is None else len(self.ch_coeff)

    def device_profile(self, device):
        if self.ch_coeff is None:
            return None, None
        return (torch.as_tensor(self.alpha_matrix, dtype=torch.float64, device=device).contiguous(),
                torch.as_tensor(self.ch_coeff, dtype=torch.float64, device=device).contiguous())

    def run(self, inputs, snr_db, z=None, normals=None):
        """inputs: float32 CUDA tensor [B,S,T,2] (transmitted IQ); snr_db: float32 CUDA tensor [B].
        Returns the received float32 tensor [B,S,T,2] (fading + AWGN fused: radio.py:228-229)."""
        alpha, coeff = self.device_profile(inputs.device)
        self._calls += 1
        return self.engine.channel(inputs, snr_db, alpha=alpha, coeff=coeff, z=z, normals=normals,
                                   seed=(self.seed << 20) + self._calls)

    __call__ = run
