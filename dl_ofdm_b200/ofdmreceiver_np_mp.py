"""Driver surface of dev/py/ofdmreceiver_np_mp.py (equalizer): transfer learning + the cross-channel BER test.

``train_equalizer`` follows the epoch loop of dev/py/ofdmreceiver_np_mp.py:394-466 on the GPU: per epoch
``frame_cnt = msg_length // nsymbol`` frames are generated (Philox bits -> OFDM TX -> Rayleigh -> AWGN with the
SNR mix of :386,405), cut into minibatches of ``batch_size // nsymbol`` frames, and every minibatch is one
``dccn_train_step`` (forward + backward of ce_mean + 0.001 * L2 w.r.t. the Equalizer variables + Adam with the
staircase learning-rate decay); then a 1024-frame test, best-train-loss checkpoint (TF bundle, the reference's
file names) and the early-stop rule.

``test_model_cross`` follows dev/py/ofdmreceiver_np_mp.py:62-104: for every test channel in
ETU/EVA/EPA/Flat/Custom and SNR -10..30 step 5, 30000 frames, one CSV per channel named
``Test_DCCN_<token>_Equalizer<opt>_<train channel>_test_chan_<chan>.csv``.  The whole
(channel x SNR) grid is one sharded sweep (config 5 of BASELINE.json).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from . import sweep, tfbundle
from .flags import parse_flags
from .init import eq_layer_roles, equalizer_variables
from .model import Session, load_model_np, save_model
from .ofdm import const_map, ofdm_tx
from .radio import rayleigh_chan_lte

# SNR mix of the training set (dev/py/ofdmreceiver_np_mp.py:386,405)
TRAIN_SNRS = np.linspace(0, 27, 10, dtype=np.float32)
TRAIN_SNR_P = [0.01, 0.01, 0.02, 0.02, 0.02, 0.02, 0.1, 0.5, 0.2, 0.1]
TRAINABLE = ['Equalizer/' + n + s for n in ('dense', 'conv3d', 'dense_1', 'dense_2', 'dense_3', 'dense_4', 'conv3d_1',
                                            'conv3d_2', 'conv3d_3', 'dense_5') for s in ('/kernel', '/bias')]


def learning_rate(init_learning, global_step):
    """tf.train.exponential_decay(init_learning, global_step, 500, 0.98, staircase=True) (_mp.py:343-344)."""
    return init_learning * 0.98 ** (int(global_step) // 500)


def save_model_name(FLAGS):
    """dev/py/ofdmreceiver_np_mp.py:349-352."""
    return FLAGS.token + ('_Equalizer_' if FLAGS.opt == 0 else '_Equalizer%d_' % FLAGS.opt) + FLAGS.channel


class _DataGen:
    """bits -> OFDM TX -> fading -> AWGN on the GPU (the host NumPy chain of _mp.py:403-413)."""

    def __init__(self, FLAGS, ofdmobj, engine, seed):
        self.fl, self.ofdm, self.eng = FLAGS, ofdmobj, engine
        self.const = const_map(FLAGS.nbits)
        self.rng = np.random.default_rng(seed)
        self.seed = int(seed)
        self.calls = 0
        self.fading0 = rayleigh_chan_lte(FLAGS, ofdmobj.Fs, mobile=False, engine=engine, seed=self.seed * 2 + 1)
        self.fading1 = rayleigh_chan_lte(FLAGS, ofdmobj.Fs, mobile=True, mix=True, engine=engine,
                                         seed=self.seed * 2 + 2) if FLAGS.mobile else None

    def make(self, frames, phase2=True):
        from .engine import bit_source_gpu
        self.calls += 1
        dev = self.eng.device
        D, nb = self.ofdm.frame_size, self.fl.nbits
        bits = bit_source_gpu(frames * D * nb, seed=(self.seed << 20) + self.calls, device=dev).view(frames, D, nb)
        snr = torch.as_tensor(self.rng.choice(TRAIN_SNRS, frames, p=TRAIN_SNR_P), dtype=torch.float32, device=dev)
        fading = self.fading1 if (phase2 and self.fading1 is not None) else self.fading0
        return fading.run_bits(bits, self.ofdm, self.const, snr), bits


def train_equalizer(FLAGS, ofdmobj, rx_weights, eq_weights=None, max_epoch_num=None, frame_cnt=None, test_frames=1024,
                    save=True, seed=None, log=print, chest_bias=(0.0, 0.0)):
    """Transfer learning of equalizer_ofdm in front of the frozen receiver ``rx_weights`` (TF names -> arrays).
    Returns (session, history) where history is a list of per-epoch dicts (train_loss, test_loss, test_ber)."""
    opt = 0 if FLAGS.opt in (9, 10) else FLAGS.opt                   # 9 / 10 build equalizer_ofdm (_mp.py:309-312)
    if opt not in (0, 1, 2, 3, 4, 5, 7):
        raise NotImplementedError('--opt=%d: dev/py/model.py defines the graphs --opt 0..5, 7 (9, 10 = 0)' % FLAGS.opt)
    seed = FLAGS.seed if seed is None else seed
    rng = np.random.default_rng(seed)
    weights = dict(rx_weights)
    weights.update(eq_weights if eq_weights is not None else
                   equalizer_variables(rng, ofdmobj.K, ofdmobj.CP, ofdmobj.nSymbol, ofdmobj.pilot_size, FLAGS.cp, opt=opt,
                                       chest_bias=chest_bias))
    trainable = ['Equalizer/' + n + sfx for _, n in eq_layer_roles(opt) for sfx in ('/kernel', '/bias')]
    batch = FLAGS.batch_size // ofdmobj.nSymbol                       # _mp.py:358
    frame_cnt = FLAGS.msg_length // FLAGS.nsymbol if frame_cnt is None else frame_cnt
    session = Session(FLAGS, ofdmobj, weights, precision=FLAGS.precision,
                      chunk_frames=max(batch, test_frames, 1024))
    eng = session.engine
    eng.train_init(batch)
    gen = _DataGen(FLAGS, ofdmobj, eng, seed)
    name = os.path.join(FLAGS.save_dir, save_model_name(FLAGS))
    test_loss_min, epoch_min_loss, history = 100.0, 0, []
    max_epoch_num = FLAGS.max_epoch_num if max_epoch_num is None else max_epoch_num
    for epoch in range(max_epoch_num):
        xs, ys = gen.make(frame_cnt)
        ce = torch.zeros(1, dtype=torch.float64, device=eng.device)
        nbits_seen = 0
        for i in range(frame_cnt // batch):
            out = eng.train_step(xs[i * batch:(i + 1) * batch], ys[i * batch:(i + 1) * batch],
                                 learning_rate(FLAGS.init_learning, eng.global_step))
            ce += out['ce_sum']
            nbits_seen += out['n_bits']
        train_loss = float(ce.cpu()[0]) / max(nbits_seen, 1)
        xt, yt = gen.make(test_frames)
        conf, test_loss = session.run(['conf_matrix', 'ce_mean'], {'tx_ofdm': xt, 'bits_in': yt})
        ber = float(conf[0, 1] + conf[1, 0]) / float(conf.sum())
        history.append(dict(epoch=epoch, train_loss=train_loss, test_loss=float(test_loss), test_ber=ber,
                            global_step=eng.global_step))
        log('Epoch: %d  Train Loss: %f  Test Loss: %f  Test BER: %.8f' % (epoch, train_loss, test_loss, ber))
        if train_loss < test_loss_min:                                # _mp.py:456-459
            epoch_min_loss, test_loss_min = epoch, train_loss
            if save:
                w = dict(weights)
                for n in trainable:
                    w[n] = eng.get_weight(n).reshape(np.shape(weights[n]))
                save_model(name, w, global_step=eng.global_step, step_name='optimizer/global_step')
        if epoch - FLAGS.early_stop > epoch_min_loss:                 # _mp.py:460-461
            break
    return session, history

TEST_CHANNELS = ['ETU', 'EVA', 'EPA', 'Flat', 'Custom']


def test_model_cross(FLAGS, path_prefix_min, ofdmobj, session=None, frame_cnt=30000, snrs=range(-10, 31, 5),
                     out_dir='.', seed=1, channels=TEST_CHANNELS):
    own = session is None
    if own:
        session = load_model_np(path_prefix_min, FLAGS=FLAGS, ofdmobj=ofdmobj, precision=FLAGS.precision)
    cells = sweep.make_cells(channels, snrs, (FLAGS.nbits,))
    conf, ce = sweep.run_sweep(cells, sweep.CellRunner(session, frame_cnt, seed), device=session.engine.device)
    rows = sweep.ber_table(cells, conf, ce)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0:
        for ch in channels:
            sub = [r for r in rows if r['channel'] == ch]
            name = 'Test_DCCN_%s_test_chan_%s%s.csv' % (FLAGS.token + '_Equalizer%d_' % FLAGS.opt + FLAGS.channel, ch,
                                                        '_mobile' if FLAGS.mobile else '')        # _mp.py:98-101
            sweep.write_csv(os.path.join(out_dir, name), sub)
    if own:
        session.close()
    return rows


def main(argv=None):
    FLAGS = parse_flags(argv, driver='mp')
    ofdmobj = ofdm_tx(FLAGS)
    rank, _ = sweep.init_distributed()
    name = FLAGS.token + ('_Equalizer_' if FLAGS.opt == 0 else '_Equalizer%d_' % FLAGS.opt) + FLAGS.channel
    path = os.path.join(FLAGS.save_dir, name)
    if not FLAGS.test and rank == 0:     # rank 0 trains and saves; every rank then takes its share of the test grid
        # transfer learning in front of the basic receiver saved by ofdmreceiver_np.py as save_dir/token (_mp.py:265-266)
        base = os.path.join(FLAGS.save_dir, FLAGS.token)
        if not os.path.exists(base + '.index'):
            raise FileNotFoundError('%s.index: no basic-receiver checkpoint to put the equalizer in front of' % base)
        rx = {k: v for k, v in tfbundle.read_checkpoint(base).items()
              if k.startswith(('fft_like/', 'demodulation/')) and '/Adam' not in k}
        session, _ = train_equalizer(FLAGS, ofdmobj, rx)
        session.close()
    sweep.barrier()
    if not os.path.exists(path + '.index'):
        raise FileNotFoundError('%s.index: no equalizer checkpoint to evaluate' % path)
    return test_model_cross(FLAGS, path, ofdmobj, frame_cnt=FLAGS.frames or 30000)    # _mp.py:73


if __name__ == '__main__':
    main(sys.argv[1:])
