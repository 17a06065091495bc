"""Driver surface of dev/py/ofdmreceiver_np.py (basic receiver): the final BER sweep.

``test_model`` reproduces the reference's -10..30 dB sweep (dev/py/ofdmreceiver_np.py:59-91) on the
GPU and writes ``Test_DCCN_<token>_<channel>.csv``.  Training of the basic receiver is not part of
this round (out of scope / next); ``main`` therefore requires ``--test=True`` or a
checkpoint to evaluate.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

from . import sweep
from .flags import parse_flags
from .model import load_model_np
from .ofdm import ofdm_tx


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
    FLAGS = parse_flags(argv)
    ofdmobj = ofdm_tx(FLAGS)
    if 'LOCAL_RANK' in os.environ and int(os.environ.get('WORLD_SIZE', 1)) > 1:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group('nccl')
    path = os.path.join(FLAGS.save_dir, FLAGS.token)
    if not os.path.exists(path + '.index'):
        raise FileNotFoundError('%s.index: no checkpoint to evaluate (training the basic receiver on the GPU '
                                'is not implemented in this round)' % path)
    return test_model(FLAGS, path, ofdmobj, frame_cnt=FLAGS.frames)


if __name__ == '__main__':
    main(sys.argv[1:])
