"""GPU probe: accuracy of the 3xTF32 path vs the fp64 oracle for several TMEM chain lengths
(DCCN_KC) and a first timing of the eq+rx pass.  Usage: python tools/acc_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import v1_weights, GOLDEN
from oracle import dccn_oracle as orc
from oracle.v1_recipe import v1_frames
from dl_ofdm_b200.engine import DCCN

w = v1_weights(np.load(os.path.join(GOLDEN, 'v1_4mod_cpTrue.npz')))
x, bits = v1_frames(4, 10, 700)
ref = orc.basic_receiver(x, w, 4, 16, head='v1', dtype=np.float64)
ref32 = orc.basic_receiver(x, w, 4, 16, head='v1', dtype=np.float32)
e32 = np.abs(ref32 - ref)
print('fp32 oracle   : p99.9 %.3g max %.3g' % (np.quantile(e32, .999), e32.max()))
xc, bc = torch.as_tensor(x).cuda(), torch.as_tensor(bits).cuda()
for prec, kc in (('exact', 0), ('parity', 1), ('parity', 2), ('parity', 4), ('parity', 8), ('parity', 0), ('fast', 0)):
    os.environ['DCCN_KC'] = str(kc)
    m = DCCN(nbits=4, nsymbol=8, n_data=368, head='v1', precision=prec)
    m.load_weights(w)
    o = m.forward(xc, bc)
    e = np.abs(o['soft'].cpu().numpy() - ref)
    hard_ref = (ref[..., 1] > ref[..., 0])
    flips = int((o['hard'].cpu().numpy().astype(bool) != hard_ref).sum())
    print('%-6s kc=%d : p99.9 %.3g max %.3g flips %d' % (prec, kc, np.quantile(e, .999), e.max(), flips))
    m.close()

# timing: 16-QAM eq + rx, B = 16384 (dev geometry, seeded weights)
rng = np.random.default_rng(0)
wd = orc.glorot_weights(rng, 4, equalizer=True, bias_scale=0.02, chest_bias=(0.6, -0.4))
B = 16384
xg = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
bg = torch.randint(0, 2, (B, 320, 4), device='cuda', dtype=torch.uint8)
for prec, kc, wide in (('exact', 0, 0), ('parity', 1, 0), ('parity', 2, 0), ('parity', 4, 0), ('parity', 2, 1), ('fast', 0, 0), ('fast', 0, 1)):
    os.environ['DCCN_KC'] = str(kc)
    os.environ['DCCN_BN_WIDE'] = str(wide)
    for eq in (True, False):
        m = DCCN(nbits=4, equalizer=eq, precision=prec, chunk_frames=4096)
        m.load_weights({k: v for k, v in wd.items() if eq or not k.startswith('Equalizer')})
        for _ in range(2):
            m.forward(xg, bg, want_soft=False)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            m.forward(xg, bg, want_soft=False)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print('%-6s kc=%d wide=%d eq=%d: %.3f ms / %d frames = %.3g frames/s' % (prec, kc, wide, eq, ms, B, B / ms * 1e3))
        m.close()
