"""Driver surface of dev/py/ofdmreceiver_np_mp.py (equalizer): the cross-channel BER test.

``test_model_cross`` follows dev/py/ofdmreceiver_np_mp.py:62-104: for every test channel in
ETU/EVA/EPA/Flat/Custom and SNR -10..30 step 5, 30000 frames, one CSV per channel named
``Test_DCCN_<token>_Equalizer<opt>_<train channel>_test_chan_<chan>.csv``.  The whole
(channel x SNR) grid is one sharded sweep (config 5 of BASELINE.json).
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
            name = 'Test_DCCN_%s_test_chan_%s.csv' % (FLAGS.token + '_Equalizer%d_' % FLAGS.opt + FLAGS.channel, ch)
            sweep.write_csv(os.path.join(out_dir, name), sub)
    if own:
        session.close()
    return rows


def main(argv=None):
    FLAGS = parse_flags(argv)
    ofdmobj = ofdm_tx(FLAGS)
    if 'LOCAL_RANK' in os.environ and int(os.environ.get('WORLD_SIZE', 1)) > 1:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group('nccl')
    name = FLAGS.token + ('_Equalizer_' if FLAGS.opt == 0 else '_Equalizer%d_' % FLAGS.opt) + FLAGS.channel
    path = os.path.join(FLAGS.save_dir, name)
    if not os.path.exists(path + '.index'):
        raise FileNotFoundError('%s.index: no equalizer checkpoint to evaluate' % path)
    return test_model_cross(FLAGS, path, ofdmobj, frame_cnt=FLAGS.frames)


if __name__ == '__main__':
    main(sys.argv[1:])
