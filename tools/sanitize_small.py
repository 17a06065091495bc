"""Small eq + rx pass for compute-sanitizer (memcheck / racecheck): chained and layer-by-layer schedules, ragged batch.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from conftest import dev_weights, GOLDEN         # noqa: E402
from dl_ofdm_b200.engine import DCCN             # noqa: E402

wt = dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz')))
rng = np.random.default_rng(0)
B = int(os.environ.get('B', 700))
x = torch.as_tensor((rng.standard_normal((B, 7, 80, 2)) * 0.7).astype(np.float32)).cuda()
bits = torch.as_tensor(rng.integers(0, 2, (B, 320, 4)).astype(np.uint8)).cuda()
for chain in ('1', '0'):
    os.environ['DCCN_CHAIN'] = chain
    m = DCCN(nbits=4, equalizer=True, precision='parity', chunk_frames=512)
    m.load_weights(wt)
    o = m.forward(x, bits, want_eq=True, want_chest=True)
    torch.cuda.synchronize()
    c = o['conf'].cpu().numpy()
    print('chain=%s conf sum %d' % (chain, c.sum()), flush=True)
    m.close()
print('done')
