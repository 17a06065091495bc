"""Measure the raw TMA -> shared-memory delivery rate of an SM (B/clk) for [128 x 32 fp32] boxes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from dl_ofdm_b200 import _lib
lib = _lib.load()
for rows, cols, name in ((21504, 896, 'act 77MB'), (896, 896, 'weight 3.2MB'), (2048, 896, 'act 7MB')):
    mat = torch.randn((rows, cols), device='cuda')
    for grid in (148, 37):
        for stages, boxes in ((6, 2), (4, 3), (12, 1), (3, 4), (2, 6)):
            clks = torch.zeros(grid, dtype=torch.int64, device='cuda')
            iters = 2000
            _lib.check(lib.dccn_debug_tma_rate(C.c_void_p(mat.data_ptr()), rows, cols, cols, stages, boxes, iters, grid,
                                               C.c_void_p(clks.data_ptr())))
            c = clks.float().mean().item()
            print('%-12s grid %3d stages %2d boxes %d : %.0f clk / stage, %.1f B/clk/SM' %
                  (name, grid, stages, boxes, c / iters, iters * boxes * 16384 / c))
