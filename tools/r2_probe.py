"""Round-2 GPU probe: (1) accuracy + per-kernel time of DCCN_KC_WHOLE_K (short-K layers keep their whole K in one TMEM
accumulator) on the shipped v1 checkpoint and on the trained eq + rx model; (2) ms per pass against elapsed time for a
3 s back-to-back run (is the sustained figure power-limited or launch-limited?).
Usage: timeout 300 python tools/r2_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import v1_weights, dev_weights, GOLDEN          # noqa: E402
from oracle import dccn_oracle as orc            # noqa: E402
from oracle.dccn_oracle_lean import LeanModel    # noqa: E402
from oracle.v1_recipe import v1_frames           # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402


def report(tag, soft, hard, ref):
    e = np.abs(soft - ref)
    flips = int((hard.astype(bool) != (ref[..., 1] > ref[..., 0])).sum())
    print('%-34s: p99.9 %.3g max %.3g flips %d' % (tag, np.quantile(e, .999), np.nanmax(e), flips), flush=True)


w1 = v1_weights(np.load(os.path.join(GOLDEN, 'v1_4mod_cpTrue.npz')))
x1, b1 = v1_frames(4, 10, 700)
ref1 = orc.basic_receiver(x1, w1, 4, 16, head='v1', dtype=np.float64)
wt = dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz')))
rng = np.random.default_rng(0)
def run_variant(tag, env, timing=True):
    for k, v in env.items():
        os.environ[k] = v
    import test_gpu_parity as tp
    m = DCCN(nbits=4, nsymbol=8, n_data=368, head='v1', precision='parity')
    m.load_weights(w1)
    o = m.forward(torch.as_tensor(x1).cuda(), torch.as_tensor(b1).cuda())
    report('v1 receiver   %s' % tag, o['soft'].cpu().numpy(), o['hard'].cpu().numpy(), ref1)
    m.close()
    m = DCCN(nbits=4, equalizer=True, precision='parity')
    m.load_weights(wt)
    for snr in (15.0, 30.0):
        x, bits = tp._config3_frames(m, 2000, snr, seed=3)
        o = m.forward(x, bits)
        z, _, _ = orc.batch_moment_norm(x.cpu().numpy(), np.float64)
        ref, _, _ = LeanModel(wt, 4).forward(z)
        report('trained eq+rx %g dB %s' % (snr, tag), o['soft'].cpu().numpy(), o['hard'].cpu().numpy(), ref)
        if tag == 'base':
            z32, _, _ = orc.batch_moment_norm(x.cpu().numpy(), np.float32)
            r32, _, _ = LeanModel(wt, 4, dtype=np.float32).forward(z32)
            report('   (numpy fp32 oracle itself, %g dB)' % snr, r32, (r32[..., 1] > r32[..., 0]), ref)
    if timing:
        xg, bg = tp._config3_frames(m, 65536, 15.0, seed=4)
        for _ in range(3):
            m.forward(xg, bg)
        torch.cuda.synchronize()
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
        print('%s: burst %.3f ms / pass   %s' % (tag, t0.elapsed_time(t1) / 10,
                                                  {k: round(v[0] / 5, 3) for k, v in sorted(prof.items())}), flush=True)
    m.close()
    for k in env:
        del os.environ[k]


run_variant('base', {})
run_variant('small_first', {'DCCN_SMALL_FIRST': '1'})
run_variant('small_first kc=2', {'DCCN_SMALL_FIRST': '1', 'DCCN_KC': '2'})
run_variant('kc=2', {'DCCN_KC': '2'}, timing=False)
run_variant('exact', {}, timing=False) if False else None
run_variant('head subs=1', {'DCCN_HEAD_SUBS': '1'})
run_variant('head subs=1 blocks=8', {'DCCN_HEAD_SUBS': '1', 'DCCN_HEAD_BLOCKS': '8'})
run_variant('head subs=2 blocks=5', {'DCCN_HEAD_SUBS': '2', 'DCCN_HEAD_BLOCKS': '5'})
