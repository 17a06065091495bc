"""tcgen05.mma rate with the parity GEMM's real operand addressing (4 smem stages, hi/lo planes, k-step offsets,
rotating TMEM staging columns, alternating accumulators, one commit per 12 MMAs) vs the same-operand loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from dl_ofdm_b200 import _lib
lib = _lib.load()
n = 4800
for mode, name in ((1, 'TS same operands       '), (2, 'TS real pattern (shipped order)'), (3, 'TS real pattern (operand-sharing order)')):
    for dep in (1, 0):
        for pc in (12, 24):
            clks = torch.zeros(148, dtype=torch.int64, device='cuda')
            _lib.check(lib.dccn_debug_mma_rate(128, n, pc, mode, dep, 148, C.c_void_p(clks.data_ptr())))
            c = clks.float().mean().item()
            print('%s %s commit every %2d : %.1f clk / MMA' % (name, 'same-acc' if dep else 'alt-acc ', pc, c / n), flush=True)
