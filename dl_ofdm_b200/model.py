"""Model surface of the reference (dev/py/model.py) on top of libdccn.

``ofdm_dense_rx`` and ``equalizer_ofdm`` keep the reference call signatures; instead of adding
ops to a TF graph they run the corresponding sub-graph entry point of the CUDA library on a
CUDA tensor and return CUDA tensors.  ``load_model_np`` restores a TF-bundle checkpoint written by
the reference (or by ``save_model``) into a ``Session`` whose ``run`` serves the reference's named
fetches (conf_matrix, linear_ber, log_ber, ce_mean, output).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, tfbundle
from .engine import DCCN
from .init import detect_eq_opt

_CACHE = []          # [(weights object, config key, engine)], newest last
_CACHE_MAX = 8


def _engine(FLAGS, ofdmobj, weights, equalizer, head, precision, eq_opt=0):
    """One engine per (weights dict, configuration).  The entry HOLDS the dict (a bare id() could be reused by a later
    object after garbage collection and serve stale weights) and is matched by identity; the oldest engines are closed
    when more than _CACHE_MAX are alive."""
    if weights is None:
        raise ValueError('weights: a dict of TF variable name -> array is required')
    key = (equalizer, head, precision, FLAGS.nbits, FLAGS.cp, FLAGS.nfilter, ofdmobj.K, ofdmobj.CP, ofdmobj.nSymbol,
           ofdmobj.frame_size, torch.cuda.current_device(), eq_opt)
    for w, k, m in _CACHE:
        if w is weights and k == key:
            return m
    m = DCCN.from_ofdm(FLAGS, ofdmobj, equalizer=equalizer, precision=precision, head=head, eq_opt=eq_opt)
    m.load_weights(weights)
    _CACHE.append((weights, key, m))
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.pop(0)[2].close()
    return m


def ofdm_dense_rx(inputs, FLAGS, ofdmobj, outshape=None, weights=None, head='dev', precision='parity'):
    """Normalised IQ [B,S,T,2] -> softmax [B, frame_size, nbits, 2] (dev/py/model.py:1222-1292)."""
    m = _engine(FLAGS, ofdmobj, weights, False, head, precision)
    out = m.forward(inputs.contiguous(), want_hard=False, flags=_lib.FWD_NO_NORM | _lib.FWD_SKIP_EQ)['soft']
    if outshape is not None:
        assert tuple(out.shape[1:]) == tuple(int(v) for v in outshape[1:])
    return out


def equalizer_ofdm(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """Normalised IQ [B,S,T,2] -> (equalized [B,S,T,2], snr_db [B,1], chest complex [B,S,K])
    (dev/py/model.py:349-478; snr_db is the monitor of :464-475)."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 0)


def _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, opt):
    m = _engine(FLAGS, ofdmobj, weights, True, 'dev', precision, opt)
    o = m.forward(inputs.contiguous(), want_soft=False, want_hard=False, want_eq=True, want_chest=True,
                  flags=_lib.FWD_NO_NORM | _lib.FWD_EQ_ONLY,
                  snr_pilot_carriers=ofdmobj.pilotCarriers if opt == 0 else None)    # the ablation graphs: monitor not served
    return o['eq'], o['snr_db'], torch.view_as_complex(o['chest'])


def equalizer_nocconv(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=1 (dev/py/model.py:482-609): dense instead of the learned-DFT conv, dense -> dense tail."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 1)


def equalizer_noresdl(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=2 (dev/py/model.py:612-714): one dense after the pilots, tf.ifft tail."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 2)


def equalizer_dnnE(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=3 (dev/py/model.py:953-1084): all-dense equalizer."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 3)


def equalizer_noresdl2(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=4 (dev/py/model.py:718-826)."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 4)


def equalizer_separateIQ(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=7 (dev/py/model.py:1088-1218): equalizer_ofdm's wiring with layers_conv2d_vector and a tanh chain."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 7)


def equalizer_noresdl4(inputs, FLAGS, ofdmobj, weights=None, precision='parity'):
    """--opt=5 (dev/py/model.py:829-950)."""
    return _run_equalizer(inputs, FLAGS, ofdmobj, weights, precision, 5)


class Session:
    """Stands in for tf.Session + imported graph: holds the engine, serves the named fetches."""

    # the named tensors of dev/py/ofdmreceiver_np.py:172-183 (+ 'hard', the argmax the confusion matrix is built from)
    FETCHES = ('conf_matrix', 'linear_ber', 'log_ber', 'ce_mean', 'output', 'hard', 'cost', 'tx_power', 'noise_power',
               'input', 'iq_tx', 'iq_rx')
    MONITORS = ('tx_power', 'noise_power', 'input', 'iq_tx', 'iq_rx')

    def __init__(self, FLAGS, ofdmobj, weights, precision='parity', head=None, chunk_frames=0):
        self.FLAGS, self.ofdm = FLAGS, ofdmobj
        self.weights = weights
        has_eq = any(k.startswith('Equalizer/') for k in weights)
        if head is None:
            head = 'v1' if 'demodulation/conv2d_1/kernel' in weights else 'dev'
        # the Equalizer/* variable names identify the graph (--opt) a checkpoint was trained with
        self.engine = DCCN.from_ofdm(FLAGS, ofdmobj, equalizer=has_eq, precision=precision, head=head,
                                     chunk_frames=chunk_frames, eq_opt=detect_eq_opt(weights) if has_eq else 0)
        self.engine.load_weights(weights)

    def run(self, fetches, feed):
        """feed: {'tx_ofdm': float32 [B,S,T,2], 'bits_in': uint8/int [B,D,nbits]} (CUDA or numpy)."""
        dev = self.engine.device
        x = torch.as_tensor(feed['tx_ofdm'], dtype=torch.float32, device=dev).contiguous()
        y = feed.get('bits_in')
        if y is not None:
            y = torch.as_tensor(y, device=dev).to(torch.uint8).contiguous()
        single = isinstance(fetches, str)
        names = [fetches] if single else list(fetches)
        for n in names:
            if n not in self.FETCHES:
                raise KeyError('unknown fetch %r (have %s)' % (n, ', '.join(self.FETCHES)))
        o = self.engine.forward(x, y, want_soft='output' in names, want_hard='hard' in names)
        mon = None
        if any(n in self.MONITORS for n in names):
            snr = feed.get('SNR')
            mon = self.engine.monitors(x, snr, seed=int(feed.get('seed', 0)), want_input='input' in names,
                                       want_iq=('iq_tx' in names or 'iq_rx' in names))
        res = []
        conf = o['conf'].cpu().numpy() if o['conf'] is not None else None
        for n in names:
            if n in self.MONITORS:
                res.append(mon[n])
                continue
            if n == 'cost':
                # total_loss = ce_mean + berlin * 1e-4 * sum(reg) + log(BER)  (ofdmreceiver_np.py:162-171); reg = the keras
                # l2(0.01) terms of the two dense layers of the receiver (model.py:1270-1286)
                ber = (conf[0, 1] + conf[1, 0]) / conf.sum()
                reg = sum(0.01 * float(np.sum(np.square(np.asarray(self.weights[k], dtype=np.float64))))
                          for k in ('demodulation/dense/kernel', 'demodulation/dense/bias', 'demodulation/dense_1/kernel',
                                    'demodulation/dense_1/bias'))
                ce = float(o['ce_sum'].cpu()[0]) / o['n_bits']
                res.append(np.float32(ce + ber * 1e-4 * reg + (np.log(ber) if ber > 0 else -np.inf)))
                continue
            if n == 'conf_matrix':
                res.append(conf)
            elif n == 'linear_ber':
                res.append(np.float32((conf[0, 1] + conf[1, 0]) / conf.sum()))
            elif n == 'log_ber':
                res.append(np.log((conf[0, 1] + conf[1, 0]) / conf.sum()) if conf[0, 1] + conf[1, 0] else -np.inf)
            elif n == 'ce_mean':
                res.append(np.float32(float(o['ce_sum'].cpu()[0]) / o['n_bits']))
            elif n == 'output':
                res.append(o['soft'])
            elif n == 'hard':
                res.append(o['hard'])
        return res[0] if single else res

    def close(self):
        self.engine.close()


def load_model_np(path, session=None, FLAGS=None, ofdmobj=None, precision='parity'):
    """Restore ``path``.index/.data (TF bundle) -> Session (dev/py/model.py:51-72)."""
    weights = tfbundle.read_checkpoint(path)
    for n in ('global_step', 'optimizer/global_step'):
        weights.pop(n, None)
    return Session(FLAGS, ofdmobj, weights, precision=precision)


def save_model(path, weights, global_step=0, step_name='global_step', FLAGS=None, ofdmobj=None):
    """Write a TF-bundle-v2 checkpoint with the reference's variable names.  The basic receiver's step counter is
    ``global_step`` (dev/py/ofdmreceiver_np.py:185), the equalizer driver's lives in its 'optimizer' variable scope
    (``optimizer/global_step``, dev/py/ofdmreceiver_np_mp.py:335-343).
    With ``FLAGS`` / ``ofdmobj`` a ``.meta`` graph is written next to a BASIC-RECEIVER bundle (tfmeta.write_meta: the
    graph of ofdmreceiver_np.py with the named tensors the reference's ``load_model_np`` fetches after
    ``import_meta_graph``, dev/py/model.py:51-72).  Equalizer bundles get no ``.meta``: the spliced graph of
    ofdmreceiver_np_mp.py:264-320 is not emitted."""
    w = dict(weights)
    w[step_name] = np.asarray(global_step, dtype=np.float32).reshape(())
    tfbundle.write_checkpoint(path, w)
    if FLAGS is not None and ofdmobj is not None and not any(k.startswith('Equalizer/') for k in w):
        from . import tfmeta
        head = 'v1' if 'demodulation/conv2d_1/kernel' in w else 'dev'
        tfmeta.write_meta(path, FLAGS.nbits, ofdmobj, nfilter=FLAGS.nfilter, cp=FLAGS.cp, head=head)
