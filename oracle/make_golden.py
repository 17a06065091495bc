#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE  (build container only).

Imports the reference's NumPy modules from /root/reference/dev/py (``ofdm.py``,
``radio.py``) behind a stub ``tensorflow`` module (they import TF at the top but
the functions used here are pure NumPy), and reads the shipped v1 checkpoints
under /root/reference/test_v1/model.  /root/reference does not exist on the GPU
box, hence the committed fixtures.  Re-run:  python oracle/make_golden.py

Fixtures written
  ofdm_tx_<pilot>_<nb>b.npz   geometry index sets + bits -> IQ of ofdm_tx_frame_np
  const_map.npz               the four constellation tables
  rayleigh_<chan>.npz         tx, the N(0,1/2) path draws, rx of rayleigh_chan_lte.run
  awgn.npz                    x, snr, normals, out of AWGN_channel_np
  v1_<tag>.npz                live tensors of two shipped v1 checkpoints
  v1_known_answers.npz        BER known answers of all 8 checkpoints (oracle fp64 run)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference'
OUT = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REPO)


def _import_reference():
    tf = types.ModuleType('tensorflow')
    tf.disable_eager_execution = lambda: None
    tf.__version__ = '1.15.0'
    sys.modules['tensorflow'] = tf
    if not hasattr(np, 'complex'):
        np.complex = complex            # removed alias used at ofdm.py:163
    sys.path.insert(0, os.path.join(REF, 'dev', 'py'))
    os.chdir(os.path.join(REF, 'dev', 'py'))           # radio.py loads ./3gpp/*.csv
    import ofdm as ref_ofdm
    import radio as ref_radio
    return ref_ofdm, ref_radio


class Flags:
    def __init__(self, **kw):
        self.nbits, self.nfft, self.nsymbol = 1, 64, 7
        self.npilot, self.nguard, self.nfilter = 8, 8, 64
        self.pilot, self.channel = 'lte', 'EPA'
        self.cp, self.longcp = True, True
        self.__dict__.update(kw)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_ofdm, ref_radio = _import_reference()

    # ---- constellation tables ------------------------------------------------
    np.savez(os.path.join(OUT, 'const_map.npz'),
             **{'ord%d' % o: ref_ofdm.const_map(o) for o in (1, 2, 3, 4)})

    # ---- transmitter + geometry ----------------------------------------------
    cases = [('lte', 7, True, nb) for nb in (1, 2, 3, 4)] + [('scattered', 8, True, 4),
                                                          ('lte', 7, False, 2)]
    for pilot, nsym, longcp, nb in cases:
        fl = Flags(nbits=nb, pilot=pilot, nsymbol=nsym, longcp=longcp)
        tx = ref_ofdm.ofdm_tx(fl)
        rng = np.random.default_rng(100 + nb)
        bits = rng.integers(0, 2, (6, tx.frame_size, nb))
        cpx, real, pilots = tx.ofdm_tx_frame_np(bits.copy())
        tag = '%s_%db%s' % (pilot, nb, '' if longcp else '_shortcp')
        np.savez(os.path.join(OUT, 'ofdm_tx_%s.npz' % tag), bits=bits.astype(np.uint8),
                 cpx=cpx, real=real.astype(np.float64), dataSc=tx.dataSc, pilotSc=tx.pilotSc,
                 guardSc=tx.guardSc, effecCarriers=tx.effecCarriers,
                 pilotCarriers=tx.pilotCarriers, dataCarriers=tx.dataCarriers,
                 meta=np.array([tx.K, tx.CP, tx.P, tx.G, tx.DC, tx.frame_size, tx.pilot_size, tx.nRB]),
                 Fs=tx.Fs)

    # ---- Rayleigh static FIR ---------------------------------------------------
    fl = Flags(nbits=2)
    txo = ref_ofdm.ofdm_tx(fl)
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2, (8, txo.frame_size, 2))
    cpx, _, _ = txo.ofdm_tx_frame_np(bits)
    for chan in ('EPA', 'EVA', 'ETU', 'Flat', 'Custom'):
        fading = ref_radio.rayleigh_chan_lte(Flags(channel=chan), txo.Fs)
        np.random.seed(1234)
        y, H = fading.run(cpx)
        # replay the generator to expose the per-frame draws radio.py:432 made
        np.random.seed(1234)
        zs = []
        for _ in range(cpx.shape[0]):
            zr = np.random.normal(loc=0.0, scale=1.0 / np.sqrt(2), size=[fading.n_taps, 2])
            zs.append(zr[:, 0] + 1j * zr[:, 1])
        np.savez(os.path.join(OUT, 'rayleigh_%s.npz' % chan.lower()), tx=cpx, z=np.array(zs),
                 rx=y, H=H, ch_coeff=fading.ch_coeff, alpha=fading.alpha_matrix)

    # ---- AWGN ------------------------------------------------------------------
    rng = np.random.default_rng(11)
    x = rng.standard_normal((16, 7, 80, 2)) * 0.13
    snr = rng.choice(np.linspace(0, 27, 10), size=(16, 1))
    np.random.seed(4321)
    out, npw = ref_radio.AWGN_channel_np(x, snr)
    np.random.seed(4321)
    normals = np.random.randn(*x.shape)
    np.savez(os.path.join(OUT, 'awgn.npz'), x=x, snr=snr, normals=normals, out=out, noise_power=npw)

    # ---- v1 checkpoints (live tensors only: centre tap of fft_like) -------------
    from dl_ofdm_b200 import tfbundle
    for nb, snr_tag, cp in ((4, 12, True), (1, 3, False)):
        name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, snr_tag, cp)
        w = tfbundle.read_checkpoint(os.path.join(REF, 'test_v1', 'model', name))
        k = w['fft_like/conv3d/kernel']
        T = k.shape[1]
        live = {n: v for n, v in w.items() if n != 'fft_like/conv3d/kernel' and n != 'global_step'}
        live['fft_like/conv3d/kernel_center'] = k[0, (T - 1) // 2, 0]       # [T,128]
        live['fft_like/conv3d/kernel_shape'] = np.array(k.shape)
        np.savez_compressed(os.path.join(OUT, 'v1_%dmod_cp%s.npz' % (nb, cp)),
                            **{n.replace('/', '.'): v for n, v in live.items()})

    # ---- v1 known answers (all 8 checkpoints, oracle fp64, BASELINE.md recipe) ---
    from oracle import dccn_oracle as orc
    from oracle.v1_recipe import v1_frames
    rows = {}
    for nb in (1, 2, 3, 4):
        for cp in (True, False):
            name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, 3 * nb, cp)
            w = tfbundle.read_checkpoint(os.path.join(REF, 'test_v1', 'model', name))
            bers = []
            for snr_db in (0, 5, 10, 15, 22):
                x, bits = v1_frames(nb, snr_db, 2000)
                soft = orc.basic_receiver(x, w, nb, 16, use_cp=cp, head='v1', dtype=np.float64)
                _, conf, ber, _ = orc.ber_head(soft, bits)
                bers.append(ber)
                print(name, snr_db, '%.4e' % ber, flush=True)
            rows['%dmod_cp%s' % (nb, cp)] = np.array(bers)
    np.savez(os.path.join(OUT, 'v1_known_answers.npz'), snr=np.array([0, 5, 10, 15, 22]), **rows)



def make_doppler():
    """rayleigh_<chan>_mobile.npz: Doppler (sum-of-sinusoids) branch of rayleigh_chan_lte.run
    (dev/py/radio.py:387-422, 491-506) with the per-frame uniform phase draws exposed."""
    ref_ofdm, ref_radio = _import_reference()
    fl = Flags(nbits=2)
    txo = ref_ofdm.ofdm_tx(fl)
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, (5, txo.frame_size, 2))
    cpx, _, _ = txo.ofdm_tx_frame_np(bits)
    for chan in ('ETU', 'EVA', 'Flat'):
        fading = ref_radio.rayleigh_chan_lte(Flags(channel=chan), txo.Fs, mobile=True)
        np.random.seed(777)
        y, H = fading.run(cpx)
        np.random.seed(777)
        th = []
        for _ in range(cpx.shape[0]):
            tre = np.random.uniform(0, 2 * np.pi, size=(fading.ss, fading.n_taps))
            tim = np.random.uniform(0, 2 * np.pi, size=(fading.ss, fading.n_taps))
            th.append(np.stack([tre, tim]))
        np.savez(os.path.join(OUT, 'rayleigh_%s_mobile.npz' % chan.lower()), tx=cpx, theta=np.array(th), rx=y,
                 ch_coeff=fading.ch_coeff, alpha=fading.alpha_matrix, Fd=fading.Fd, Fs=txo.Fs)


def make_alpha_npz():
    """dl_ofdm_b200/data/lte_alpha.npz: the reference's fractional-delay interpolation matrices
    (dev/py/3gpp/AM_*.csv, exported from MATLAB's channelFilter.alphaMatrix, README.md:60) packed
    as one npz -- numeric channel-model data the product reads at run time."""
    d = {}
    for ch, fn in (('epa', 'AM_EPA.csv'), ('eva', 'AM_EVA.csv'), ('etu', 'AM_ETU.csv'), ('custom', 'AM_Custom.csv')):
        d[ch] = np.genfromtxt(os.path.join(REF, 'dev', 'py', '3gpp', fn), delimiter=',')
    np.savez(os.path.join(REPO, 'dl_ofdm_b200', 'data', 'lte_alpha.npz'), **d)


if __name__ == '__main__':
    if 'doppler' in sys.argv:
        make_doppler()
    else:
        make_alpha_npz()
        make_doppler()
        main()
