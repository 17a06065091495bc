"""Is the eq+rx pass bound by the host's launch rate?  Prints, per chunk size, the host time to enqueue one pass
(no sync), the device time of the pass and the sum of the per-kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dccn_oracle as orc
from dl_ofdm_b200.engine import DCCN
B = int(os.environ.get('B', 65536))
rng = np.random.default_rng(0)
wd = orc.glorot_weights(rng, 4, equalizer=True, bias_scale=0.02, chest_bias=(0.6, -0.4))
xg = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
bits = torch.randint(0, 2, (B, 320, 4), device='cuda', dtype=torch.uint8)
for chunk in [int(c) for c in os.environ.get('CHUNKS', '0,8192,10752,16384,21504,32768,65536').split(',')]:
    m = DCCN(nbits=4, equalizer=True, precision='parity', chunk_frames=chunk)
    m.load_weights(wd)
    for _ in range(3):
        m.forward(xg, bits, want_soft=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        m.forward(xg, bits, want_soft=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    dev = e0.elapsed_time(e1) / n
    m.profile(True)
    for _ in range(n):
        m.forward(xg, bits, want_soft=True)
    prof = m.profile_collect()
    m.profile(False)
    ksum = sum(v[0] for v in prof.values()) / n
    nl = sum(v[1] for v in prof.values()) / n
    print('chunk %6d  host enqueue %.2f ms  device %.2f ms  sum of kernels %.2f ms  (%d profiled launches)  -> %.3g frames/s'
          % (chunk, 1e3 * (t1 - t0) / n, dev, ksum, nl, B / dev * 1e3), flush=True)
    m.close()
