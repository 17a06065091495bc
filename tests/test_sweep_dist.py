"""CPU tests of the multi-rank sweep logic: gloo, world_size 2 (the N>1 path of bench / drivers)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist
import torch.multiprocessing as mp

from dl_ofdm_b200 import sweep


def _fake_cell(i, cell):
    nb, ch, snr = cell
    rng = np.random.default_rng(1000 + i)
    c = rng.integers(0, 1000, (2, 2))
    return c, float(c.sum()) * 0.25


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cells = sweep.make_cells(['ETU', 'EPA', 'Flat'], range(-10, 31, 5), (4,))
    owned = sweep.shard(cells, rank, world)
    conf, ce = sweep.run_sweep(cells, _fake_cell)
    q.put((rank, owned, conf, ce))
    dist.destroy_process_group()


def test_sweep_allreduce_world2():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    cells = sweep.make_cells(['ETU', 'EPA', 'Flat'], range(-10, 31, 5), (4,))
    ref_conf = np.stack([_fake_cell(i, c)[0] for i, c in enumerate(cells)])
    ref_ce = np.array([_fake_cell(i, c)[1] for i, c in enumerate(cells)])
    owned = sorted(sum((r[1] for r in res), []))
    assert owned == list(range(len(cells)))                      # every cell owned exactly once
    for rank, own, conf, ce in res:
        assert own == [i for i in range(len(cells)) if i % 2 == rank]
        assert np.array_equal(conf, ref_conf)                    # == single-process sum on every rank
        assert np.allclose(ce, ref_ce)


def test_single_process_sweep_and_csv(tmp_path):
    cells = sweep.make_cells(['EPA'], [0, 5], (2,))
    conf, ce = sweep.run_sweep(cells, _fake_cell)
    rows = sweep.ber_table(cells, conf, ce)
    assert len(rows) == 2 and rows[0]['SNR'] == 0.0
    c0 = _fake_cell(0, cells[0])[0]
    assert abs(rows[0]['BER'] - (c0[0, 1] + c0[1, 0]) / c0.sum()) < 1e-12
    p = tmp_path / 'out' / 'Test_DCCN_x_EPA.csv'
    sweep.write_csv(str(p), rows)
    lines = p.read_text().strip().splitlines()
    assert lines[0] == 'SNR,BER,Loss' and len(lines) == 3


def test_launcher_job_list_matches_reference_order():
    from dl_ofdm_b200.run_local_ofdm import job_list
    from dl_ofdm_b200.flags import parse_flags
    save_dir, result_dir, jobs = job_list(True)
    assert save_dir == './ofdm_lte_ext_64_longcp_mobile/' and result_dir == './test_ext_64_long_cross_mobile'
    assert len(jobs) == 2 * (4 * 2 + 2)
    script, flags, csv = jobs[0]                      # reversed([4,3,2,1]) starts with BPSK (run_local_ofdm.py:66)
    f = parse_flags(flags.split())
    assert script == 'ofdmreceiver_np' and csv == 'Test_DCCN_OFDM_Dense3_1mod_snr5_cpFalse_AWGN.csv'
    assert (f.nbits, f.cp, f.longcp, f.channel, f.nfilter, f.SNR, f.max_epoch_num, f.early_stop, f.test) == \
        (1, False, False, 'AWGN', 64, 5.0, 1200, 200, False)
    script, flags, csv = jobs[8]
    f = parse_flags(flags.split())
    assert script == 'ofdmreceiver_np_mp' and csv == 'Test_DCCN_OFDM_Dense3_1mod_snr5_cpTrue_Equalizer0_mixRayleigh_test_chan_Custom.csv'
    assert (f.channel, f.opt, f.nbits, f.cp, f.mobile, f.max_epoch_num, f.token) == \
        ('mixRayleigh', 0, 1, True, True, 4000, 'OFDM_Dense3_1mod_snr5_cpTrue')
    f = parse_flags(jobs[-1][1].split())
    assert (f.longcp, f.cp) == (True, False)
