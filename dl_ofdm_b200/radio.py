"""Channel emulators on the GPU with the reference's ``radio.py`` surface.

``rayleigh_chan_lte(FLAGS, sample_rate, mobile, mix).run(iq)`` and
``AWGN_channel_np(x, SNR)`` keep the reference signatures (dev/py/radio.py:277-526)
but run as CUDA kernels (``dccn_chan_fir_awgn``): per-frame CN(0,1) path gains,
``g = (z*ch_coeff) @ alpha``, centred 'same' FIR with zero history, then batch-power
normalisation + AWGN.  Random numbers come from a Philox counter stream (seeded, so
sweeps are reproducible -- the reference seeds NumPy from the wall clock), or can be
injected for parity tests.  Both the static branch and the mobile (Doppler, sum-of-sinusoids)
branch are implemented, as is the per-frame profile cycling of 'mixRayleigh' / 'mixAll'.
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


_FD = {'etu': 300.0, 'epa': 5.0, 'eva': 70.0, 'custom': 80.0}      # Doppler spread when mobile (radio.py:343-360)


def doppler_hz(chan, mobile):
    """Maximum Doppler shift of a profile: 0 when static; the single-tap channel uses 5 Hz (radio.py:363-366).
    'awgn' has no path to fade: the reference passes those frames through untouched whatever Fd is
    (radio.py:443-446, 470-474), so its Doppler is 0."""
    if not mobile or chan.lower() == 'awgn':
        return 0.0
    return _FD.get(chan.lower(), 5.0)


class rayleigh_chan_lte:
    """GPU version of dev/py/radio.py:277-510.  ``channel`` may be a profile (EPA/EVA/ETU/Custom/Flat),
    'AWGN' (no fading), or 'mixRayleigh' / 'mixAll' which deal the frames round-robin to
    flat/etu/eva/epa (and awgn); with ``mobile`` the Doppler (sum-of-sinusoids) branch is used, with
    ``mix`` only on every third (fourth for mixAll) frame like the reference."""

    def __init__(self, FLAGS, sample_rate=0.96e6, mobile=False, mix=False, engine=None, seed=0):
        self.nSymbol = FLAGS.nsymbol
        self.chan = FLAGS.channel.lower()
        self.sample_rate = sample_rate
        self.nfft = FLAGS.nfft
        self.mobile, self.mix = bool(mobile), bool(mix)
        self.engine = engine
        self.seed = seed
        self._calls = 0
        self._dev = {}
        if self.chan == 'mixrayleigh':
            self.profiles = ['flat', 'etu', 'eva', 'epa']
        elif self.chan == 'mixall':
            self.profiles = ['awgn', 'flat', 'etu', 'eva', 'epa']
        else:
            self.profiles = [self.chan]
        self.Fd = doppler_hz(self.profiles[-1], self.mobile) if len(self.profiles) == 1 else None
        if len(self.profiles) == 1 and self.chan != 'awgn':
            self.ch_coeff, self.alpha_matrix = channel_profile(self.chan)
            self.n_taps = len(self.ch_coeff)
        else:
            self.ch_coeff, self.alpha_matrix, self.n_taps = None, None, 0

    def _profile(self, name, device):
        key = (name, str(device))
        if key not in self._dev:
            if name == 'awgn':
                self._dev[key] = (None, None)
            else:
                c, a = channel_profile(name)
                self._dev[key] = (torch.as_tensor(a, dtype=torch.float64, device=device).contiguous(),
                                  torch.as_tensor(c, dtype=torch.float64, device=device).contiguous())
        return self._dev[key]

    def fade(self, inputs, draws=None):
        """Fading only: float32 CUDA [B,S,T,2] -> faded [B,S,T,2] (complex64 like radio.py:492)."""
        eng = self.engine
        faded = torch.empty_like(inputs)
        self._calls += 1
        n = len(self.profiles)
        dmod = 3 if self.chan == 'mixrayleigh' else 4
        first = True
        for k, name in enumerate(self.profiles):
            alpha, coeff = self._profile(name, inputs.device)
            fd = doppler_hz(name, self.mobile)
            seed = (self.seed << 24) + (self._calls << 4) + k
            if n == 1:
                eng.fading(inputs, faded, alpha, coeff, fd, self.sample_rate, draws, seed, 0, 1, first)
                first = False
                continue
            if n > 1 and fd > 0.1 and self.mix:
                # frames with i % n == k and i % dmod == 0 are Doppler, the others static; lcm stride
                lcm = n * dmod // np.gcd(n, dmod)
                for f0 in range(k, lcm, n):
                    dop = fd if (f0 % dmod == 0) else 0.0
                    eng.fading(inputs, faded, alpha, coeff, dop, self.sample_rate, None, seed + f0 * 131, f0, lcm, first)
                    first = False
            else:
                eng.fading(inputs, faded, alpha, coeff, 0.0, self.sample_rate, None, seed, k, n, first)
                first = False
        return faded

    def run(self, inputs, snr_db, z=None, normals=None):
        """inputs: float32 CUDA tensor [B,S,T,2] (transmitted IQ); snr_db: float32 CUDA tensor [B].
        Returns the received float32 tensor [B,S,T,2] (fading + AWGN: radio.py:228-229)."""
        faded = self.fade(inputs, z)
        return self.engine.awgn(faded, snr_db, normals, seed=(self.seed << 20) + self._calls)

    __call__ = run

    def run_bits(self, bits, ofdmobj, constellation, snr_db, z=None, normals=None):
        """bits uint8 CUDA [B,D,nbits] -> received float32 [B,S,T,2]: the reference's chain
        ``ofdm_tx_frame_np -> fading.run -> AWGN_channel_np`` (dev/py/ofdmreceiver_np.py:227-229).  For a single static
        profile (or 'AWGN') on the nfft = 64 geometry the transmitter and the FIR run as ONE kernel (`dccn_tx_fade`: the
        transmitted frames never reach HBM) with the same bits as the separate calls; otherwise it is
        ``run(engine.transmit(bits))``."""
        eng = self.engine
        static = len(self.profiles) == 1 and doppler_hz(self.profiles[0], self.mobile) == 0.0
        if not (static and eng.can_tx_fade()):
            return self.run(eng.transmit(bits, ofdmobj, constellation), snr_db, z, normals)
        self._calls += 1
        alpha, coeff = self._profile(self.profiles[0], bits.device)
        seed = (self.seed << 24) + (self._calls << 4)           # == fade()'s seed for profile index 0
        faded, _ = eng.transmit_fade(bits, ofdmobj, constellation, alpha, coeff, z, seed)
        return eng.awgn(faded, snr_db, normals, seed=(self.seed << 20) + self._calls)
