"""Round-2 GPU probe 2: per-kernel times of a config-3 pass with / without the phase-equaliser epilogue's L2 prefetch.
Usage: timeout 300 python tools/r2_probe2.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import dev_weights, GOLDEN         # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402
import test_gpu_parity as tp                     # noqa: E402

wt = dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz')))


def run_variant(tag, env):
    for k, v in env.items():
        os.environ[k] = v
    m = DCCN(nbits=4, equalizer=True, precision='parity')
    m.load_weights(wt)
    xg, bg = tp._config3_frames(m, 65536, 15.0, seed=4)
    for _ in range(3):
        m.forward(xg, bg)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        o = m.forward(xg, bg)
    t1.record()
    torch.cuda.synchronize()
    m.profile(True)
    for _ in range(5):
        m.forward(xg, bg)
    prof = m.profile_collect()
    m.profile(False)
    print('%s: burst %.3f ms / pass  ber %.5f  %s' % (tag, t0.elapsed_time(t1) / 20, float((lambda c: (c[0, 1] + c[1, 0]) / c.sum())(o['conf'].cpu().numpy().astype(float))),
                                                     {k: round(v[0] / 5, 3) for k, v in sorted(prof.items())}), flush=True)
    m.close()
    for k in env:
        del os.environ[k]


for rep in range(2):
    run_variant('prefetch on ', {})
    run_variant('prefetch off', {'DCCN_EPI_PREFETCH': '0'})
