"""One config-3 cell for profiling under ncu (not a benchmark): Philox bits -> fused transmitter + EPA FIR -> AWGN ->
norm -> equalizer_ofdm -> ofdm_dense_rx -> BER on the trained weights, B frames, ITERS times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                   # noqa: E402
import torch                                         # noqa: E402
from conftest import dev_weights, GOLDEN             # noqa: E402
from dl_ofdm_b200.engine import DCCN, bit_source_gpu  # noqa: E402
from dl_ofdm_b200.flags import Flags                 # noqa: E402
from dl_ofdm_b200.ofdm import const_map, ofdm_tx     # noqa: E402
from dl_ofdm_b200.radio import rayleigh_chan_lte     # noqa: E402

B = int(os.environ.get('B', 65536))
iters = int(os.environ.get('ITERS', 2))
fl = Flags(nbits=4, channel='EPA')
o = ofdm_tx(fl)
m = DCCN.from_ofdm(fl, o, equalizer=True, precision='parity')
m.load_weights(dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz'))))
snr = torch.full((B,), 15.0, device=m.device)
for it in range(iters):
    bits = bit_source_gpu(B * o.frame_size * 4, seed=it, device=m.device).view(B, o.frame_size, 4)
    x = rayleigh_chan_lte(fl, o.Fs, engine=m, seed=it).run_bits(bits, o, const_map(4), snr)
    out = m.forward(x, bits)
torch.cuda.synchronize()
c = out['conf'].cpu().numpy()
print('BER %.5f' % ((c[0, 1] + c[1, 0]) / c.sum()))
