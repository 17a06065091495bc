"""In-kernel timeline of one GEMM layer (debug build with -DDCCN_TRACE, see tools/build_trace.sh).
Prints, for CTA 0, the clock of every hand-off: roles 0 B-producer issue, 1 A-producer issue, 2 splitter got A,
3 splitter got TMEM slot, 4 splitter signalled ready, 5 MMA warp has operands, 6 epilogue got accumulator,
7 epilogue released accumulator."""
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
lib.dccn_debug_trace.restype = C.c_int
lib.dccn_debug_trace.argtypes = [C.c_void_p]
m.forward(xg, want_soft=False)
torch.cuda.synchronize()
buf = torch.zeros(10 * 4096, dtype=torch.int64, device='cuda')
# trace ONLY the layer selected by the profile slot order: run EQ_ONLY pass and keep the last writer.  Simplest: trace
# the whole pass; every GEMM overwrites the buffer, so the LAST GEMM of the pass is what remains -> choose flags.
# DCCN_TRACE_SLOT (library profile slot index: 2 eq_dense, 3 eq_dft, 6 eq_dense3, 7 eq_dense4_tanh, 8 conv7x64,
# 9 corr_idft, 10 idft, 11 dense5, 12 rx_fft_like, 16 rx_demod_gemm) selects the GEMM whose CTA 0 writes the buffer
lib.dccn_debug_trace(C.c_void_p(buf.data_ptr()))
if os.environ.get('ABL'):
    lib.dccn_debug_abl.argtypes = [C.c_int]
    lib.dccn_debug_abl(int(os.environ['ABL']))
m.forward(xg, want_soft=False)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(10, 4096)
t0 = t[t > 0].min()
names = ['epi tile0', 'epi tile12', 'mma top', 'mma tempty', 'spl ready', 'mma ops', 'epi got', 'epi rel', 'mma commit', 'mma issued']
n = int(os.environ.get('N', 40))
for r in range(10):
    v = t[r][t[r] > 0] - t0
    print('%-10s n=%4d' % (names[r], len(v)), ' '.join('%6d' % x for x in v[:n]))
    if len(v) > 8:
        d = np.diff(v)
        print('           diff median %d  p90 %d  last-first %d' % (np.median(d), np.quantile(d, 0.9), v[-1] - v[0]))
