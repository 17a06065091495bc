"""Timing ablations of the parity GEMM pipeline (debug build with -DDCCN_TRACE, tools/build_trace.sh).
Each mask switches pieces of the pipeline off (results are garbage, timing is what is measured):
1 splitter skips lds+split, 2 splitter skips tcgen05.st, 4 epilogue skips tcgen05.ld, 8 no B TMA, 16 no A TMA,
32 no MMA issue, 64 cvt.rna split instead of the integer split."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('DCCN_LIB', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dl_ofdm_b200', 'libdccn_trace.so'))
import numpy as np, torch
from oracle import dccn_oracle as orc
from dl_ofdm_b200 import _lib
from dl_ofdm_b200.engine import DCCN
B = int(os.environ.get('B', 21504))
rng = np.random.default_rng(0)
wd = orc.glorot_weights(rng, 4, equalizer=True, bias_scale=0.02, chest_bias=(0.6, -0.4))
xg = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
m = DCCN(nbits=4, equalizer=True, precision='parity', chunk_frames=B)
m.load_weights(wd)
lib = _lib.load()
lib.dccn_debug_abl.restype = C.c_int
lib.dccn_debug_abl.argtypes = [C.c_int]
masks = [int(x) for x in os.environ.get('MASKS', '0,64,1,2,3,4,8,16,24,32,7,15,31').split(',')]
show = os.environ.get('SHOW', 'eq_dft,eq_dense4_tanh,eq_dense5,rx_demod_gemm').split(',')
for mask in masks:
    lib.dccn_debug_abl(mask)
    for _ in range(2):
        m.forward(xg, want_soft=False)
    torch.cuda.synchronize()
    m.profile(True)
    for _ in range(3):
        m.forward(xg, want_soft=False)
    prof = m.profile_collect()
    m.profile(False)
    print('abl %3d ' % mask + '  '.join('%s %.1f us' % (k, 1e3 * prof[k][0] / max(1, prof[k][1])) for k in show), flush=True)
