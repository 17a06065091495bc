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


# Wiring of the equalizer graphs selected by --opt (dev/py/ofdmreceiver_np_mp.py:292-311):
#   opt 0 equalizer_ofdm (model.py:349), 1 equalizer_nocconv (:482), 2 equalizer_noresdl (:612), 3 equalizer_dnnE (:953),
#   4 equalizer_noresdl2 (:718), 5 equalizer_noresdl4 (:829).
# front2: second per-symbol layer ('cconv' = (1,K) valid complex conv, 'dense'); chain: frame-level dense layers after
# the pilot bottleneck (0 linear, 1 tanh); toeplitz: (S,K) 'same' complex conv; tail: 'corr' | 'dense2' | 'ifft'.
#   opt 7 equalizer_separateIQ (:1088): equalizer_ofdm's wiring with layers_conv2d_vector (complex.py:199-255) and tanh chain.
#   (opt 6 names equalizer_doppler, which dev/py/model.py does not define; opt 9 / 10 build equalizer_ofdm.)
EQ_SPECS = {
    7: dict(front2='cconv', chain=(1, 1, 1), toeplitz=True, tail='corr', vector=True),
    0: dict(front2='cconv', chain=(0, 0, 1), toeplitz=True, tail='corr'),
    1: dict(front2='dense', chain=(0, 0, 1), toeplitz=True, tail='dense2'),
    2: dict(front2='cconv', chain=(0,), toeplitz=False, tail='ifft'),
    4: dict(front2='cconv', chain=(0, 1), toeplitz=False, tail='ifft'),
    5: dict(front2='cconv', chain=(0, 1, 1, 1), toeplitz=False, tail='ifft'),
    3: dict(front2='dense', chain=(1, 1, 1, 1), toeplitz=False, tail='dense2'),
}


def eq_layer_roles(opt):
    """[(role, tf layer name)] in creation order; TF-1 numbers the layers of a scope dense, dense_1, ..., conv3d, ..."""
    sp = EQ_SPECS[opt]
    cnt = {'dense': 0, 'conv3d': 0}
    out = []

    def add(role, kind):
        out.append((role, kind if cnt[kind] == 0 else '%s_%d' % (kind, cnt[kind])))
        cnt[kind] += 1

    add('front1', 'dense')
    add('front2', 'conv3d' if sp['front2'] == 'cconv' else 'dense')
    add('pilot', 'dense')
    for i in range(len(sp['chain'])):
        add('chain%d' % i, 'dense')
    if sp['toeplitz']:
        add('toeplitz', 'conv3d')
    if sp['tail'] == 'corr':
        add('tail_corr', 'conv3d')
        add('tail_eq', 'conv3d')
    elif sp['tail'] == 'dense2':
        add('tail1', 'dense')
    add('tail2', 'dense')
    return out


def detect_eq_opt(weights):
    """Which --opt graph a variable dict belongs to (the layer counts of the six graphs are all different)."""
    names = {k.split('/')[1] for k in weights if k.startswith('Equalizer/') and k.endswith('/kernel')}
    for opt in (0, 1, 2, 3, 4, 5):
        if names == {n for _, n in eq_layer_roles(opt)}:
            if opt == 0 and np.shape(weights['Equalizer/conv3d/kernel'])[2] == 2:
                return 7          # same layer list, conv3d kernels of depth 2 across IQ (layers_conv2d_vector)
            return opt
    raise ValueError('Equalizer/* variables match none of the implemented graphs (--opt 0,1,2,3,4,5): %s' % sorted(names))


def equalizer_variables(rng, nfft=64, cp_len=16, nsymbol=7, pilot_size=16, use_cp=True,
                        chest_bias=(0.0, 0.0), opt=0):
    """Variables of scope 'Equalizer' with TF's initialisers (glorot-uniform kernels, zero biases).
    ``chest_bias`` is an explicit opt-in that is NOT in the reference: (1, 0) seeds conv3d_1's bias (the last chain
    bias in the graphs without that layer) so that an UNTRAINED channel estimate starts near 1+0j instead of 0 -- the
    phase-only equaliser divides by |chest| without an epsilon (dev/py/model.py:430-433).  The default (0, 0) is the
    reference's trajectory."""
    K, S = nfft, nsymbol
    Tin = K + cp_len if use_cp else K
    e = 'Equalizer/'
    SK2 = S * K * 2
    if opt in (9, 10):
        opt = 0
    if opt != 0:
        return _variant_variables(rng, opt, K, S, Tin, K + cp_len, pilot_size, chest_bias)

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


def _variant_variables(rng, opt, K, S, Tin, T, pilot_size, chest_bias):
    sp = EQ_SPECS[opt]
    e = 'Equalizer/'
    SK2 = S * K * 2
    shapes = {'front1': (Tin * 2, K * 2), 'front2': (1, K, 2 * K) if sp['front2'] == 'cconv' else (K * 2, K * 2),
              'pilot': (SK2, pilot_size * 2), 'toeplitz': (S, K, 2), 'tail_corr': (1, K, 2 * K), 'tail_eq': (1, K, 2 * K),
              'tail1': (K * 2, K * 2), 'tail2': (K * 4 if sp['tail'] == 'corr' else K * 2, T * 2)}
    for i in range(len(sp['chain'])):
        shapes['chain%d' % i] = (pilot_size * 2 if i == 0 else SK2, SK2)
    w = {}
    roles = eq_layer_roles(opt)
    for role, name in roles:
        sh = shapes[role]
        if len(sh) == 2:
            w[e + name + '/kernel'] = _glorot(rng, sh, sh[0], sh[1])
            w[e + name + '/bias'] = np.zeros(sh[1], np.float32)
        elif sp.get('vector'):
            rf = sh[0] * sh[1] * 2
            w[e + name + '/kernel'] = _glorot(rng, (sh[0], sh[1], 2, 1, sh[2]), rf, rf * sh[2])
            w[e + name + '/bias'] = np.zeros(sh[2], np.float32)
        else:
            rf = sh[0] * sh[1]
            w[e + name + '/kernel'] = _glorot(rng, (sh[0], sh[1], 1, 1, sh[2]), rf, rf * sh[2])
            w[e + name + '/bias'] = np.zeros(sh[2], np.float32)
    names = dict(roles)
    if sp['toeplitz']:
        w[e + names['toeplitz'] + '/bias'] = np.asarray(chest_bias, np.float32)
    else:   # the chest is the last chain layer's output: start it near chest_bias instead of 0 (no epsilon in |chest|)
        b = np.empty(SK2, np.float32)
        b[0::2], b[1::2] = chest_bias[0], chest_bias[1]
        w[e + names['chain%d' % (len(sp['chain']) - 1)] + '/bias'] = b
    return w
