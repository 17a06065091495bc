"""Two config-4 training steps (QPSK, B frames) for profiling (ncu) -- not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dccn_oracle as orc
from dl_ofdm_b200.engine import DCCN
B = int(os.environ.get('B', 4096)); prec = os.environ.get('PREC', 'parity'); iters = int(os.environ.get('ITERS', 2))
rng = np.random.default_rng(0)
wd = orc.glorot_weights(rng, 2, equalizer=True, bias_scale=0.02, chest_bias=(0.6, -0.4))
xg = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
bg = torch.randint(0, 2, (B, 320, 2), device='cuda', dtype=torch.uint8)
m = DCCN(nbits=2, equalizer=True, precision=prec, chunk_frames=B)
m.load_weights(wd)
m.train_init(B)
for _ in range(iters):
    o = m.train_step(xg, bg, 1e-3)
torch.cuda.synchronize()
print(float(o['ce_sum'][0]) / o['n_bits'], m.global_step)
