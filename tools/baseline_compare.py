"""BER of the learned receiver next to the expert receivers the reference compares it with (dev/m/OFDM_Benchmark_dev.m via
dev/m/script_rayleigh.m), on the SAME frames: 16-QAM, 20 000 frames per point (the MATLAB script's Nframes), SNR -10..30
step 5 (its SNRs), channels Flat / EPA / ETU.  Frames come from the library's GPU feeder with injected path gains (so the
true channel is known to the 'perfect' and 'ideal LMMSE' comparators); DCCN = the committed GPU-trained 16-QAM receiver +
equalizer_ofdm (tests/golden/dev_4mod_eq_trained.npz: 400 + 300 epochs, EPA only -- far short of the reference's
1200 x nbits / 4000 epoch schedules); comparators = dl_ofdm_b200/baselines.py (NumPy, host).
Usage: python tools/baseline_compare.py > profiles/baseline_compare_r2.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                                   # noqa: E402
import torch                                         # noqa: E402
from conftest import dev_weights, GOLDEN             # noqa: E402
from dl_ofdm_b200.baselines import ClassicReceiver   # noqa: E402
from dl_ofdm_b200.engine import DCCN, bit_source_gpu  # noqa: E402
from dl_ofdm_b200.flags import Flags                 # noqa: E402
from dl_ofdm_b200.ofdm import const_map, ofdm_tx     # noqa: E402
from dl_ofdm_b200.radio import rayleigh_chan_lte, channel_profile   # noqa: E402

nb, B = 4, 20000
fl = Flags(nbits=nb)
o = ofdm_tx(fl)
m = DCCN.from_ofdm(fl, o, equalizer=True, precision='parity')
m.load_weights(dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz'))))
cr = ClassicReceiver(o, nb)
rng = np.random.default_rng(7)
print('# %s' % __doc__.strip().splitlines()[0])
print('# channel SNR_dB   DCCN(trained here)  perfect-CSI  ideal-LMMSE  LS-spline  LS-linear      (BER, %d frames x 1280 bits per point)' % B)
for chan in ('Flat', 'EPA', 'ETU'):
    coeff, alpha = channel_profile(chan)
    for snr in range(-10, 31, 5):
        bits = bit_source_gpu(B * o.frame_size * nb, seed=1000 + snr, device=m.device).view(B, o.frame_size, nb)
        z = (rng.standard_normal((B, len(coeff))) + 1j * rng.standard_normal((B, len(coeff)))) * np.sqrt(.5)
        zt = torch.as_tensor(np.stack([z.real, z.imag], -1)).cuda().contiguous()
        ch = rayleigh_chan_lte(fl.copy(channel=chan), o.Fs, engine=m, seed=snr + 50)
        x = ch.run_bits(bits, o, const_map(nb), torch.full((B,), float(snr), device=m.device), z=zt)
        out = m.forward(x, bits, want_soft=False, want_hard=False)
        c = out['conf'].cpu().numpy()
        g = (z * coeff) @ alpha
        xs, bs = x.cpu().numpy(), bits.cpu().numpy()
        r = [cr.ber(xs, bs, e, float(snr), g) for e in ('perfect', 'lmmse', 'ls_spline', 'ls_linear')]
        print('%-6s %4d        %.4e         %.4e   %.4e   %.4e  %.4e' % (chan, snr, (c[0, 1] + c[1, 0]) / c.sum(), *r), flush=True)
