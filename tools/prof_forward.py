"""One eq+rx pass at B frames for profiling (ncu) -- not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dccn_oracle as orc
from dl_ofdm_b200.engine import DCCN
B = int(os.environ.get('B', 16384)); prec = os.environ.get('PREC', 'parity'); eq = int(os.environ.get('EQ', 1))
chunk = int(os.environ.get('CHUNK', 4096)); iters = int(os.environ.get('ITERS', 2))
rng = np.random.default_rng(0)
wd = orc.glorot_weights(rng, 4, equalizer=bool(eq), bias_scale=0.02, chest_bias=(0.6, -0.4))
xg = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
bg = torch.randint(0, 2, (B, 320, 4), device='cuda', dtype=torch.uint8)
m = DCCN(nbits=4, equalizer=bool(eq), precision=prec, chunk_frames=chunk)
m.load_weights(wd)
for _ in range(iters):
    o = m.forward(xg, bg, want_soft=False)
torch.cuda.synchronize()
print(o['conf'].cpu().numpy().tolist())
