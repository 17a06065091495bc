"""Driver surface of dev/py/ofdmreceiver_np.py (basic receiver): training and the final BER sweep.

``train_receiver`` follows the epoch loop of dev/py/ofdmreceiver_np.py:193-274 on the GPU: per epoch
``frame_cnt = msg_length // nsymbol`` frames (bits -> OFDM TX -> fading -> AWGN at ``FLAGS.SNR``, :211-229), cut into
minibatches of ``batch_size // nsymbol`` frames, every minibatch one ``dccn_train_step`` in DCCN_TRAIN_RX mode (forward,
backward of ``ce_mean + berlin * 1e-4 * L2`` w.r.t. all eight receiver variables, Adam with the staircase decay of
:186-189); the minibatch grows with the reference's ``idealbatchsize`` rule (:243-244); then a 1024-frame test, the
best-train-loss checkpoint ``save_dir/token`` in TF-bundle format (:266-270) and the early-stop rule (:271-272).

``test_model`` reproduces the reference's -10..30 dB sweep (dev/py/ofdmreceiver_np.py:59-91) on the
GPU and writes ``Test_DCCN_<token>_<channel>.csv``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from . import sweep
from .flags import parse_flags
from .init import receiver_variables
from .model import Session, load_model_np, save_model
from .ofdm import const_map, ofdm_tx
from .radio import rayleigh_chan_lte

RX_TRAINABLE = [n + s for n in ('fft_like/conv3d', 'demodulation/dense', 'demodulation/conv2d', 'demodulation/dense_1')
                for s in ('/kernel', '/bias')]


def learning_rate(global_step, init_learning=0.001):
    """tf.train.exponential_decay(0.001, global_step, 500, 0.98, staircase=True) (ofdmreceiver_np.py:186-187)."""
    return init_learning * 0.98 ** (int(global_step) // 500)


def train_receiver(FLAGS, ofdmobj, weights=None, max_epoch_num=None, frame_cnt=None, test_frames=1024, save=True,
                   seed=None, log=print):
    """Training of the basic receiver from glorot-uniform variables (or ``weights``).
    Returns (session, history); history holds per-epoch train_loss / test_loss / test_ber / batch."""
    from .engine import bit_source_gpu
    seed = FLAGS.seed if seed is None else seed
    rng = np.random.default_rng(seed)
    nb, D, S = FLAGS.nbits, ofdmobj.frame_size, ofdmobj.nSymbol
    if weights is None:
        weights = receiver_variables(rng, nb, ofdmobj.K, ofdmobj.CP, S, FLAGS.nfilter, D, FLAGS.cp)
    frame_cnt = FLAGS.msg_length // FLAGS.nsymbol if frame_cnt is None else frame_cnt
    batch = max(1, min(FLAGS.batch_size // FLAGS.nsymbol, frame_cnt))           # :196
    session = Session(FLAGS, ofdmobj, weights, precision=FLAGS.precision, chunk_frames=max(frame_cnt, test_frames, 1024))
    eng = session.engine
    eng.train_init(frame_cnt, mode='rx')                                        # the minibatch may grow up to the epoch
    const = const_map(nb)
    fading = rayleigh_chan_lte(FLAGS, ofdmobj.Fs, engine=eng, seed=seed * 2 + 1)
    dev = eng.device
    calls = 0

    def make(frames):
        nonlocal calls
        calls += 1
        bits = bit_source_gpu(frames * D * nb, seed=(int(seed) << 20) + calls, device=dev).view(frames, D, nb)
        snr = torch.full((frames,), float(FLAGS.SNR), dtype=torch.float32, device=dev)   # snr_seq is all zeros (:207)
        return fading.run_bits(bits, ofdmobj, const, snr), bits

    name = os.path.join(FLAGS.save_dir, FLAGS.token)
    test_loss_min, epoch_min_loss, history = 100.0, 0, []
    max_epoch_num = FLAGS.max_epoch_num if max_epoch_num is None else max_epoch_num
    for epoch in range(max_epoch_num):
        xs, ys = make(frame_cnt)
        ce = torch.zeros(1, dtype=torch.float64, device=dev)
        nseen, conf_last = 0, None
        for i in range(frame_cnt // batch):
            out = eng.train_step(xs[i * batch:(i + 1) * batch], ys[i * batch:(i + 1) * batch], learning_rate(eng.global_step))
            ce += out['ce_sum']
            nseen += out['n_bits']
            conf_last = out['conf']
        train_loss = float(ce.cpu()[0]) / max(nseen, 1)
        c = conf_last.cpu().numpy()
        berl = float(c[0, 1] + c[1, 0]) / float(c.sum())                       # BER of the last minibatch (:242)
        idealbatchsize = int(min(200.0 / max(berl, 1.e-6), 900000.) / (55 * nb)) // 8      # :243
        batch = min(max(batch, idealbatchsize), frame_cnt)                                  # :244
        xt, yt = make(test_frames)
        conf, test_loss = session.run(['conf_matrix', 'ce_mean'], {'tx_ofdm': xt, 'bits_in': yt})
        ber = float(conf[0, 1] + conf[1, 0]) / float(conf.sum())
        history.append(dict(epoch=epoch, train_loss=train_loss, test_loss=float(test_loss), test_ber=ber, batch=batch,
                            global_step=eng.global_step))
        log('Epoch: %d  Train Loss: %f  Test Loss: %f  Test BER: %.8f' % (epoch, train_loss, test_loss, ber))
        if train_loss < test_loss_min:                                          # :266-270
            epoch_min_loss, test_loss_min = epoch, train_loss
            if save:
                w = dict(weights)
                for n in RX_TRAINABLE:
                    w[n] = eng.get_weight(n).reshape(np.shape(weights[n]))
                save_model(name, w, global_step=eng.global_step, FLAGS=FLAGS, ofdmobj=ofdmobj)
        if epoch - FLAGS.early_stop > epoch_min_loss:                           # :271-272
            break
    return session, history


def test_model(FLAGS, path_prefix_min, ofdmobj, session=None, frame_cnt=20000, snrs=range(-10, 31),
               out_dir='.', seed=1):
    own = session is None
    if own:
        session = load_model_np(path_prefix_min, FLAGS=FLAGS, ofdmobj=ofdmobj, precision=FLAGS.precision)
    cells = sweep.make_cells([FLAGS.channel], snrs, (FLAGS.nbits,))
    conf, ce = sweep.run_sweep(cells, sweep.CellRunner(session, frame_cnt, seed), device=session.engine.device)
    rows = sweep.ber_table(cells, conf, ce)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0:
        for r in rows:
            print('SNR: %.2f, BER: %.8f, Loss: %f' % (r['SNR'], r['BER'], r['Loss']))
        sweep.write_csv(os.path.join(out_dir, 'Test_DCCN_%s.csv' % (FLAGS.token + '_' + FLAGS.channel)), rows)
    if own:
        session.close()
    return rows


def main(argv=None):
    FLAGS = parse_flags(argv, driver='np')
    ofdmobj = ofdm_tx(FLAGS)
    rank, _ = sweep.init_distributed()
    path = os.path.join(FLAGS.save_dir, FLAGS.token)
    if not FLAGS.test and rank == 0:
        # training is one sequential trajectory (the reference has no multi-GPU training): rank 0 trains and writes the
        # checkpoint (atomically, tfbundle.write_checkpoint), the other ranks wait and then share the BER sweep
        os.makedirs(FLAGS.save_dir, exist_ok=True)
        session, _ = train_receiver(FLAGS, ofdmobj)
        session.close()
    sweep.barrier()
    if not os.path.exists(path + '.index'):
        raise FileNotFoundError('%s.index: no checkpoint to evaluate' % path)
    return test_model(FLAGS, path, ofdmobj, frame_cnt=FLAGS.frames or 20000)      # :69


if __name__ == '__main__':
    main(sys.argv[1:])
