"""BER sweep over the (modulation, channel, SNR) grid, sharded across the GPUs of one node.

The reference loops the grid serially in one process (dev/py/ofdmreceiver_np.py:72,
dev/py/ofdmreceiver_np_mp.py:74,81, dev/py/run_local_ofdm.py:61-72).  Cells are independent, and the
batch-moment norm (Q3) / batch-power AWGN (Q9) make one cell an indivisible statistical unit, so
the grid is dealt round-robin to the ranks with a full weight replica each; the ONLY collective is
one all-reduce(sum) of the int64 confusion matrices (+ the loss sums) at the end -- NCCL over NVLink
on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed():
    """Join the torchrun job once (idempotent: the launcher calls the drivers' main() once per job in one process).
    -> (rank, world)."""
    if int(os.environ.get('WORLD_SIZE', 1)) > 1 and 'LOCAL_RANK' in os.environ:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        if not dist.is_initialized():
            dist.init_process_group('nccl')
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def make_cells(channels, snrs, nbits_list=(None,)):
    """Deterministic cell order: modulation-major, then channel, then SNR (the reference's loop nest)."""
    return [(nb, ch, float(s)) for nb in nbits_list for ch in channels for s in snrs]


def shard(cells, rank, world):
    """Round-robin ownership: cell i belongs to rank i % world."""
    return [i for i in range(len(cells)) if i % world == rank]


def reduce_results(conf, ce, device=None):
    """Sum [n_cells,2,2] int64 confusion matrices and [n_cells] float64 loss sums over all ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            conf, ce = conf.to(device), ce.to(device)
        dist.all_reduce(conf, op=dist.ReduceOp.SUM)
        dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    return conf, ce


def run_sweep(cells, run_cell, device=None):
    """Run the cells this rank owns with ``run_cell(index, cell) -> (conf int64[2,2], ce_sum)`` and return the reduced
    (conf [n,2,2] numpy, ce_sum [n] numpy) on every rank.  A runner may return DEVICE tensors (CellRunner does): they are
    gathered on the device without a host synchronisation per cell, so the cells of a rank run back to back, and the only
    collective -- one all-reduce of the stacked matrices and loss sums -- runs on them directly."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    n = len(cells)
    conf = torch.zeros((n, 2, 2), dtype=torch.int64, device=device)
    ce = torch.zeros((n,), dtype=torch.float64, device=device)
    for i in shard(cells, rank, world):
        c, l = run_cell(i, cells[i])
        if torch.is_tensor(c):
            conf[i].copy_(c.reshape(2, 2), non_blocking=True)
            ce[i:i + 1].copy_(l.reshape(1), non_blocking=True)
        else:
            conf[i] = torch.as_tensor(np.asarray(c), dtype=torch.int64)
            ce[i] = float(l)
    conf, ce = reduce_results(conf, ce, device)
    return conf.cpu().numpy(), ce.cpu().numpy()


def ber_table(cells, conf, ce):
    rows = []
    for (nb, ch, snr), c, l in zip(cells, conf, ce):
        tot = float(c.sum())
        rows.append({'nbits': nb, 'channel': ch, 'SNR': snr, 'BER': float(c[0, 1] + c[1, 0]) / tot if tot else float('nan'),
                     'Loss': float(l) / tot if tot else float('nan'), 'bits': int(tot)})
    return rows


def write_csv(path, rows):
    """Reference result format: columns SNR,BER,Loss indexed by SNR (ofdmreceiver_np.py:70,85-89)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as f:
        f.write('SNR,BER,Loss\n')
        for r in rows:
            f.write('%g,%.10g,%.10g\n' % (r['SNR'], r['BER'], r['Loss']))


class CellRunner:
    """GPU cell: Philox bits -> OFDM TX -> [Rayleigh FIR] + AWGN -> norm -> [equalizer] -> receiver -> BER."""

    def __init__(self, session, frames, seed=0):
        from .ofdm import const_map
        self.s, self.frames, self.seed = session, int(frames), int(seed)
        self.const = const_map(session.FLAGS.nbits)

    def __call__(self, index, cell):
        from .engine import bit_source_gpu
        from .radio import rayleigh_chan_lte
        _, chan_name, snr = cell
        eng, fl, ofdm = self.s.engine, self.s.FLAGS, self.s.ofdm
        B, D, nb = self.frames, ofdm.frame_size, fl.nbits
        dev = eng.device
        bits = bit_source_gpu(B * D * nb, seed=(self.seed << 24) + index, device=dev).view(B, D, nb)
        chan = rayleigh_chan_lte(fl.copy(channel=chan_name), ofdm.Fs, mobile=getattr(fl, 'mobile', False),
                                 engine=eng, seed=(self.seed << 12) + index)
        x = chan.run_bits(bits, ofdm, self.const, torch.full((B,), snr, dtype=torch.float32, device=dev))
        o = eng.forward(x, bits, want_soft=False, want_hard=False)
        return o['conf'], o['ce_sum']                      # device tensors: no host synchronisation per cell
