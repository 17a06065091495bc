"""Golden data for the BER-vs-SNR known-answer tests (north_star "+-0.1 dB over -10..29 dB").

Reads the reference's shipped v1 checkpoints (test_v1/model/*, read-only) and writes
  tests/golden/v1_2mod_cpTrue.npz, v1_3mod_cpTrue.npz   live tensors (centre tap of fft_like) of the QPSK / 8-QAM receivers
  tests/golden/v1_curves.npz                            bit errors per SNR point -10..29 dB of ALL 8 checkpoints on the seeded
                                                        recipe frames (oracle/v1_recipe.py, 2000 frames per point), fp64 oracle
Run here (CPU, needs /root/reference):  python oracle/make_golden_curves.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get('DL_OFDM_REF', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')


def main():
    from dl_ofdm_b200 import tfbundle
    from oracle import dccn_oracle as orc
    from oracle.v1_recipe import v1_frames
    for nb, cp in ((2, True), (3, True)):
        name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, 3 * nb, cp)
        w = tfbundle.read_checkpoint(os.path.join(REF, 'test_v1', 'model', name))
        k = w['fft_like/conv3d/kernel']
        T = k.shape[1]
        live = {n: v for n, v in w.items() if n not in ('fft_like/conv3d/kernel', 'global_step')}
        live['fft_like/conv3d/kernel_center'] = k[0, (T - 1) // 2, 0]
        live['fft_like/conv3d/kernel_shape'] = np.array(k.shape)
        np.savez_compressed(os.path.join(OUT, 'v1_%dmod_cp%s.npz' % (nb, cp)),
                            **{n.replace('/', '.'): v for n, v in live.items()})
    snrs = np.arange(-10, 30)
    rows = {'snr': snrs}
    for nb in (1, 2, 3, 4):
        frames = {s: v1_frames(nb, int(s), 2000) for s in snrs}
        for cp in (True, False):
            name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, 3 * nb, cp)
            w = tfbundle.read_checkpoint(os.path.join(REF, 'test_v1', 'model', name))
            errs = []
            for s in snrs:
                x, bits = frames[s]
                soft = orc.basic_receiver(x, w, nb, 16, use_cp=cp, head='v1', dtype=np.float64)
                _, conf, _, _ = orc.ber_head(soft, bits)
                errs.append(int(conf[0, 1] + conf[1, 0]))
            rows['%dmod_cp%s_errors' % (nb, cp)] = np.array(errs, dtype=np.int64)
            rows['%dmod_cp%s_bits' % (nb, cp)] = np.int64(2000 * 368 * nb)
            print(name, ' '.join('%.3e' % (e / (2000 * 368 * nb)) for e in errs[::5]), flush=True)
    np.savez(os.path.join(OUT, 'v1_curves.npz'), **rows)


if __name__ == '__main__':
    main()
