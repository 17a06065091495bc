"""Measure tcgen05.mma (kind::tf32, M=128) execution rate of an SM: clk per MMA for SS / TS forms,
dependent vs alternating accumulators, and with a tcgen05.commit every n MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from dl_ofdm_b200 import _lib
lib = _lib.load()
n = 4800
for bn in (128, 256):
    for mode in ((0, 1) if bn == 128 else (0,)):
        for dep in (1, 0):
            for pc in (0, 48, 12, 4):
                clks = torch.zeros(148, dtype=torch.int64, device='cuda')
                _lib.check(lib.dccn_debug_mma_rate(bn, n, pc, mode, dep, 148, C.c_void_p(clks.data_ptr())))
                c = clks.float().mean().item()
                print('N=%3d %s %s commit every %2d : %.1f clk / MMA  (ideal %d)' %
                      (bn, 'TS' if mode else 'SS', 'same-acc' if dep else 'alt-acc ', pc, c / n, bn // 2))
