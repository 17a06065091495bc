"""Variable creation with the reference's initialisers (no TensorFlow).

``tf.layers.dense/conv2d/conv3d`` default to glorot-uniform kernels and zero
biases; names, shapes and creation order follow the variable inventory of
``ofdm_dense_rx`` (dev/py/model.py:1246-1288) and ``equalizer_ofdm``
(dev/py/model.py:370-461), i.e. exactly what ``dccn_set_weight`` expects.
"""
from __future__ import annotations

import numpy as np


def _glorot(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def receiver_variables(rng, nbits, nfft=64, cp_len=16, nsymbol=7, nfilter=64, n_data=320,
                       use_cp=True, head='dev'):
    T = nfft + cp_len if use_cp else nfft
    M = 1 << nbits
    w = {
        'fft_like/conv3d/kernel': _glorot(rng, (1, T, 1, T, 2 * nfilter), T * T, T * 2 * nfilter),
        'fft_like/conv3d/bias': np.zeros(2 * nfilter, np.float32),
        'demodulation/dense/kernel': _glorot(rng, (nsymbol * nfilter * 2, n_data * 2),
                                             nsymbol * nfilter * 2, n_data * 2),
        'demodulation/dense/bias': np.zeros(n_data * 2, np.float32),
        'demodulation/conv2d/kernel': _glorot(rng, (1, 1, 2, M), 2, M),
        'demodulation/conv2d/bias': np.zeros(M, np.float32),
    }
    if head == 'v1':
        w['demodulation/conv2d_1/kernel'] = _glorot(rng, (1, 1, M, M), M, M)
        w['demodulation/conv2d_1/bias'] = np.zeros(M, np.float32)
    w['demodulation/dense_1/kernel'] = _glorot(rng, (M + 2, 2 * nbits), M + 2, 2 * nbits)
    w['demodulation/dense_1/bias'] = np.zeros(2 * nbits, np.float32)
    return w


def equalizer_variables(rng, nfft=64, cp_len=16, nsymbol=7, pilot_size=16, use_cp=True,
                        chest_bias=(1.0, 0.0)):
    """Variables of scope 'Equalizer'.  ``chest_bias`` seeds conv3d_1's bias so that an UNTRAINED
    channel estimate starts near 1+0j instead of 0 (the phase-only equaliser divides by |chest|
    without an epsilon, dev/py/model.py:430-433); pass (0, 0) for TF's literal zero init."""
    K, S = nfft, nsymbol
    Tin = K + cp_len if use_cp else K
    e = 'Equalizer/'
    SK2 = S * K * 2

    def dense(name, i, o):
        return {e + name + '/kernel': _glorot(rng, (i, o), i, o), e + name + '/bias': np.zeros(o, np.float32)}

    def conv3d(name, kl, kw, cout):
        rf = kl * kw
        return {e + name + '/kernel': _glorot(rng, (kl, kw, 1, 1, cout), rf, rf * cout),
                e + name + '/bias': np.zeros(cout, np.float32)}

    w = {}
    w.update(dense('dense', Tin * 2, K * 2))
    w.update(conv3d('conv3d', 1, K, 2 * K))
    w.update(dense('dense_1', SK2, pilot_size * 2))
    w.update(dense('dense_2', pilot_size * 2, SK2))
    w.update(dense('dense_3', SK2, SK2))
    w.update(dense('dense_4', SK2, SK2))
    w.update(conv3d('conv3d_1', S, K, 2))
    w.update(conv3d('conv3d_2', 1, K, 2 * K))
    w.update(conv3d('conv3d_3', 1, K, 2 * K))
    w.update(dense('dense_5', K * 4, (K + cp_len) * 2))
    w[e + 'conv3d_1/bias'] = np.asarray(chest_bias, np.float32)
    return w
