"""GPU probe of the chained per-symbol kernels (csrc/chain.cu): accuracy against the fp64 oracle with DCCN_CHAIN=0 / 1 on the
trained eq + rx model, chained against layer-by-layer outputs directly, then burst timing + per-kernel times of a
65 536-frame pass.   Usage: timeout 300 python tools/chain_probe.py [small]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import dev_weights, GOLDEN         # noqa: E402
from oracle import dccn_oracle as orc            # noqa: E402
from oracle.dccn_oracle_lean import LeanModel    # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402
import test_gpu_parity as tp                     # noqa: E402

wt = dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz')))
small = len(sys.argv) > 1 and sys.argv[1] == 'small'
time_only = len(sys.argv) > 1 and sys.argv[1] == 'time'      # only the chained 65 536-frame timing


def report(tag, soft, hard, ref):
    e = np.abs(soft - ref)
    flips = int((hard.astype(bool) != (ref[..., 1] > ref[..., 0])).sum())
    print('%-40s: p99.9 %.3g max %.3g flips %d' % (tag, np.quantile(e, .999), np.nanmax(e), flips), flush=True)


outs = {}
for chain in (() if time_only else ('0', '1')):
    os.environ['DCCN_CHAIN'] = chain
    m = DCCN(nbits=4, equalizer=True, precision='parity', chunk_frames=512)
    m.load_weights(wt)
    for B, snr in ((900, 15.0), (2000, 30.0)):
        x, bits = tp._config3_frames(m, B, snr, seed=3)
        o = m.forward(x, bits, want_eq=True)
        torch.cuda.synchronize()
        z, _, _ = orc.batch_moment_norm(x.cpu().numpy(), np.float64)
        ref, eq_ref, _ = LeanModel(wt, 4).forward(z)
        soft = o['soft'].cpu().numpy()
        report('trained eq+rx B=%d %g dB chain=%s' % (B, snr, chain), soft, o['hard'].cpu().numpy(), ref)
        eq = o['eq'].cpu().numpy()
        print('    eq (dense_5 output) max err %.3g of max %.3g' % (np.abs(eq - eq_ref).max(), np.abs(eq_ref).max()), flush=True)
        outs[(chain, B)] = (soft, eq)
    m.close()
for B in (() if time_only else (900, 2000)):
    d = np.abs(outs[('0', B)][0] - outs[('1', B)][0])
    de = np.abs(outs[('0', B)][1] - outs[('1', B)][1])
    print('chain vs layer-by-layer B=%d: soft max %.3g p99.9 %.3g; eq max %.3g' % (B, d.max(), np.quantile(d, .999), de.max()), flush=True)

if not small:
    for chain in (('1',) if time_only else ('0', '1')):
        os.environ['DCCN_CHAIN'] = chain
        m = DCCN(nbits=4, equalizer=True, precision='parity')
        m.load_weights(wt)
        xg, bg = tp._config3_frames(m, 65536, 15.0, seed=4)
        for _ in range(3):
            o = m.forward(xg, bg)
        torch.cuda.synchronize()
        conf = o['conf'].cpu().numpy()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            m.forward(xg, bg)
        t1.record()
        torch.cuda.synchronize()
        m.profile(True)
        for _ in range(5):
            m.forward(xg, bg)
        prof = m.profile_collect()
        m.profile(False)
        print('chain=%s: burst %.3f ms / pass  BER %.5f  %s' % (chain, t0.elapsed_time(t1) / 10,
              (conf[0, 1] + conf[1, 0]) / conf.sum(), {k: round(v[0] / 5, 3) for k, v in sorted(prof.items())}), flush=True)
        m.close()
