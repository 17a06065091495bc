"""Debug aid: does the result depend on the internal chunking (multi-tile persistent path vs single tile)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dccn_oracle as orc
from dl_ofdm_b200.engine import DCCN
B = int(os.environ.get('B', 8192))
rng = np.random.default_rng(int(os.environ.get('SEED', 42)))
w = orc.glorot_weights(rng, 4, equalizer=True, bias_scale=0.02, chest_bias=(0.6, -0.4))
g = torch.Generator(device='cuda').manual_seed(int(os.environ.get('SEED', 42)) + 1)
x = torch.randn((B, 7, 80, 2), generator=g, device='cuda') * 0.2
res = []
for chunk in [int(c) for c in os.environ.get('CHUNKS', '4096,1024').split(',')]:
    m = DCCN(nbits=4, equalizer=True, precision='parity', chunk_frames=chunk)
    m.load_weights(w)
    o = m.forward(x, None, want_soft=True, want_eq=True, want_chest=True)
    torch.cuda.synchronize()
    res.append({k: o[k].clone() for k in ('soft', 'hard', 'eq', 'chest')})
    m.close()
for k in ('chest', 'eq', 'soft', 'hard'):
    a, b = res[0][k], res[1][k]
    d = (a != b)
    print(os.environ.get('DCCN_LIB', 'default'), k, 'differing', int(d.sum()), 'of', d.numel(),
          'max abs', float((a.float() - b.float()).abs().max()), 'first frames', torch.nonzero(d.flatten())[:3].flatten().tolist())
