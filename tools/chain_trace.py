"""In-kernel timeline of the chained per-symbol kernels (csrc/chain.cu): CTA 0 records clock64() at its hand-off points.
Prints, for the first tiles of the front and the tail chain, when each k-block's MMAs were issued, when its drain started /
ended and when each stage was finalised (cycles relative to the first MMA issue).   Usage: python tools/chain_trace.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import dev_weights, GOLDEN         # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402
from dl_ofdm_b200 import _lib                    # noqa: E402
import test_gpu_parity as tp                     # noqa: E402

lib = _lib.load()      # run with DCCN_LIB=dl_ofdm_b200/libdccn_chtrace.so (tools/build_chain_trace.sh)
lib.dccn_debug_chain_trace.restype = C.c_int
lib.dccn_debug_chain_trace.argtypes = [C.c_int, C.c_void_p]
wt = dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz')))
m = DCCN(nbits=4, equalizer=True, precision='parity')
m.load_weights(wt)
x, bits = tp._config3_frames(m, 65536, 15.0, seed=4)
for _ in range(2):
    m.forward(x, bits)
torch.cuda.synchronize()
bufs = [torch.zeros(11 * 1024, dtype=torch.int64, device='cuda') for _ in range(2)]
for i in range(2):
    lib.dccn_debug_chain_trace(i, C.c_void_p(bufs[i].data_ptr()))
m.forward(x, bits)
torch.cuda.synchronize()
for i in range(2):
    lib.dccn_debug_chain_trace(i, None)
for name, buf, steps in (('front', bufs[0], 5), ('tail', bufs[1], 11)):
    t = buf.cpu().numpy().reshape(11, 1024)
    t0 = t[0, 0]
    print('== %s chain, CTA 0: cycles since the first MMA issue; %d k-block steps per tile' % (name, steps))
    print('%5s %9s %9s %9s %9s %9s %9s %9s   %s' % ('step', 'mma', 'at_wait', 'drain0', 'drain1', 'bias', 'pack', 'final', 'd(mma)') + '   last stage: patch free / written / fenced')
    for k in range(3 * steps + 2):
        if t[0, k] == 0:
            break
        fin = t[3, k] - t0 if t[3, k] else -1
        print('%5d %9d %9d %9d %9d %9d %9d %9d   %6d%s' % (k, t[0, k] - t0, t[7, k] - t0, t[1, k] - t0, t[2, k] - t0,
                                                     t[5, k] - t0 if t[5, k] else -1, t[6, k] - t0 if t[6, k] else -1, fin,
                                               t[0, k] - t[0, k - 1] if k else 0, '   <- tile' if k % steps == 0 else '') +
              ('   %d / %d / %d' % (t[8, k] - t0, t[9, k] - t0, t[10, k] - t0) if t[8, k] else ''))
    sp = t[4][t[4] > 0][:12] - t0
    print('splitter k-blocks staged at', sp.tolist())
m.close()
