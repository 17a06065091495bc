"""BASELINE config 5 on two GPUs: the grid sharded over two NCCL ranks == the single-GPU sweep, cell by cell.
Needs two CUDA devices (skipped on the 1-GPU box the driver uses for `-m gpu`; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _grid(rank, world):
    import types
    from conftest import dev_weights, GOLDEN
    from dl_ofdm_b200 import sweep
    from dl_ofdm_b200.engine import DCCN
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.ofdm import ofdm_tx
    fl = Flags(nbits=4, channel='EPA')
    o = ofdm_tx(fl)
    m = DCCN.from_ofdm(fl, o, equalizer=True, precision='parity')
    m.load_weights(dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz'))))
    cells = sweep.make_cells(['Flat', 'EPA', 'ETU'], [0, 10, 20], (4,))
    runner = sweep.CellRunner(types.SimpleNamespace(engine=m, FLAGS=fl, ofdm=o), 3000, seed=9)
    conf, ce = sweep.run_sweep(cells, runner, device=m.device)
    m.close()
    return cells, conf, ce


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    _, conf, ce = _grid(rank, world)
    q.put((rank, conf, ce))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_grid_equals_single_gpu(libdccn):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    cells, conf1, ce1 = _grid(0, 1)                                  # single process, no process group
    assert conf1.sum() == len(cells) * 3000 * 320 * 4
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=300) for _ in ps]
    [p.join(timeout=60) for p in ps]
    for rank, conf, ce in res:
        assert np.array_equal(conf, conf1), rank                    # every rank holds the whole reduced grid
        assert np.allclose(ce, ce1, rtol=1e-12), rank
    ber = (conf1[:, 0, 1] + conf1[:, 1, 0]) / conf1.sum(axis=(1, 2))
    assert ber[2] < ber[1] < ber[0]                                  # Flat: BER falls with SNR
