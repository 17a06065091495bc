"""GPU probe for the STAGED fp16 hi/lo GEMM form (DCCN_F16X3=1; gemm_tc.cuh `F16`): accuracy against the fp64 oracle
next to the default 3xTF32 parity mode, then timing of the 16-QAM eq + rx pass.  First thing to run in a round with GPU
time -- the form was written without a GPU and has never executed.
Usage: timeout 300 python tools/f16x3_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import v1_weights, GOLDEN          # noqa: E402
from oracle import dccn_oracle as orc            # noqa: E402
from oracle.v1_recipe import v1_frames           # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402


def report(tag, soft, hard, ref):
    e = np.abs(soft - ref)
    flips = int((hard.astype(bool) != (ref[..., 1] > ref[..., 0])).sum())
    print('%-28s: p99.9 %.3g max %.3g flips %d nan %d' % (tag, np.quantile(e, .999), np.nanmax(e), flips,
                                                           int(np.isnan(soft).sum())), flush=True)


# 1. shipped v1 checkpoint (two GEMMs: K = 160, N = 128 and K = 1024, N = 736)
w = v1_weights(np.load(os.path.join(GOLDEN, 'v1_4mod_cpTrue.npz')))
x, bits = v1_frames(4, 10, 700)
ref = orc.basic_receiver(x, w, 4, 16, head='v1', dtype=np.float64)
xc, bc = torch.as_tensor(x).cuda(), torch.as_tensor(bits).cuda()
for f16 in ('0', '1'):
    os.environ['DCCN_F16X3'] = f16
    m = DCCN(nbits=4, nsymbol=8, n_data=368, head='v1', precision='parity')
    m.load_weights(w)
    o = m.forward(xc, bc)
    torch.cuda.synchronize()
    report('v1 receiver   F16X3=%s' % f16, o['soft'].cpu().numpy(), o['hard'].cpu().numpy(), ref)
    m.close()

# 2. equalizer + receiver, seeded weights (all twelve GEMMs incl. the Toeplitz band skip, N = 32 and K = 32 layers)
rng = np.random.default_rng(0)
nb, B = 4, 1000
wd = orc.glorot_weights(rng, nb, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
xe = (rng.standard_normal((B, 7, 80, 2)) * 0.1).astype(np.float32)
be = rng.integers(0, 2, (B, 320, nb)).astype(np.uint8)
ref_e, _, _ = orc.equalized_receiver(xe, wd, nb, 64, 16, dtype=np.float64)
for f16 in ('0', '1'):
    os.environ['DCCN_F16X3'] = f16
    m = DCCN(nbits=nb, equalizer=True, precision='parity')
    m.load_weights(wd)
    o = m.forward(torch.as_tensor(xe).cuda(), torch.as_tensor(be).cuda())
    torch.cuda.synchronize()
    report('eq + rx       F16X3=%s' % f16, o['soft'].cpu().numpy(), o['hard'].cpu().numpy(), ref_e)
    m.close()

# 3. timing, 65 536 frames
Bt = 65536
xg = torch.randn((Bt, 7, 80, 2), device='cuda') * 0.2
bg = torch.randint(0, 2, (Bt, 320, nb), device='cuda', dtype=torch.uint8)
for f16, kc in (('0', '1'), ('1', '1'), ('1', '2')):
    os.environ['DCCN_F16X3'] = f16
    os.environ['DCCN_KC'] = kc
    m = DCCN(nbits=nb, equalizer=True, precision='parity')
    m.load_weights(wd)
    for _ in range(3):
        m.forward(xg, bg)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        m.forward(xg, bg)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    m.profile(True)
    for _ in range(3):
        m.forward(xg, bg)
    prof = m.profile_collect()
    m.profile(False)
    print('F16X3=%s kc=%s: %.3f ms / %d frames = %.3g frames/s   %s' % (
        f16, kc, ms, Bt, Bt / ms * 1e3, {k: round(v[0] / 3, 3) for k, v in sorted(prof.items())}), flush=True)
    m.close()
