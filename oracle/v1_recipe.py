"""Seeded v1 test-frame recipe of BASELINE.md section 2  --  TEST INFRASTRUCTURE ONLY.

Generates the AWGN frames on which the reference's 8 shipped v1 checkpoints give
the known-answer BER table.  Geometry follows test_v1/OFDM_Benchmark.m:33-54
(N=64, 4+4 guards, DC {31,32}, 8 scattered pilots = 3+3j rotating +3 per symbol,
46 data carriers per symbol, 8 symbols, CP 16) which is what the reference's
``ofdm_tx`` builds for pilot='scattered', nsymbol=8, npilot=8, nguard=8
(dev/py/ofdm.py:198-238).  Constellations restate dev/py/ofdm.py:24-78.
"""
import numpy as np


def constellation(nbits):
    """Literal restatement of the reference dictionaries, MSB-first index."""
    if nbits == 1:
        tab = {(0,): -4.24264 + 0j, (1,): 4.24264 + 0j}
    elif nbits == 2:
        tab = {(0, 0): -3 + 3j, (1, 0): -3 - 3j, (0, 1): 3 + 3j, (1, 1): 3 - 3j}
    elif nbits == 3:
        s = 4.2426 / 3.1623
        tab = {(0, 0, 0): (-3 + 1j) * s, (1, 0, 0): (-3 - 1j) * s, (0, 1, 0): (-1 + 1j) * s,
               (1, 1, 0): (-1 - 1j) * s, (0, 0, 1): (3 + 1j) * s, (1, 0, 1): (3 - 1j) * s,
               (0, 1, 1): (1 + 1j) * s, (1, 1, 1): (1 - 1j) * s}
    elif nbits == 4:
        tab = {(0, 0, 0, 0): -3 + 3j, (1, 0, 0, 0): -3 + 1j, (0, 1, 0, 0): -3 - 3j, (1, 1, 0, 0): -3 - 1j,
               (0, 0, 1, 0): -1 + 3j, (1, 0, 1, 0): -1 + 1j, (0, 1, 1, 0): -1 - 3j, (1, 1, 1, 0): -1 - 1j,
               (0, 0, 0, 1): 3 + 3j, (1, 0, 0, 1): 3 + 1j, (0, 1, 0, 1): 3 - 3j, (1, 1, 0, 1): 3 - 1j,
               (0, 0, 1, 1): 1 + 3j, (1, 0, 1, 1): 1 + 1j, (0, 1, 1, 1): 1 - 3j, (1, 1, 1, 1): 1 - 1j}
    else:
        raise ValueError(nbits)
    out = np.empty(2 ** nbits, dtype=np.complex64)
    for bits, v in tab.items():
        idx = 0
        for b in bits:
            idx = (idx << 1) | b
        out[idx] = v
    return out


def v1_geometry():
    eff = np.array([k for k in range(4, 60) if k not in (31, 32)])
    base = np.arange(0, 54, 7)
    pilots, data = [], []
    for s in range(8):
        loc = np.sort((base + 3 * s) % 54)
        p = eff[loc]
        pilots.append(p)
        data.append(np.setdiff1d(eff, p))
    return eff, pilots, data


def v1_frames(nbits, snr_db, n_frames=2000, seed_base=1000):
    """-> (x float32 [n,8,80,2], bits uint8 [n, 368, nbits]) per the BASELINE.md recipe."""
    rng = np.random.default_rng(seed_base + snr_db)
    bits = rng.integers(0, 2, (n_frames, 8, 46, nbits))
    wts = 1 << np.arange(nbits - 1, -1, -1)
    sym = constellation(nbits)[(bits * wts).sum(-1)]
    _, pilots, data = v1_geometry()
    X = np.zeros((n_frames, 8, 64), dtype=np.complex128)
    for s in range(8):
        X[:, s, data[s]] = sym[:, s, :]
        X[:, s, pilots[s]] = 3 + 3j
    x = np.fft.ifft(X, axis=-1)
    x = np.concatenate([x[..., -16:], x], axis=-1)
    p = np.mean(np.abs(x) ** 2)
    n = (rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) * np.sqrt(p / 2 * 10 ** (-snr_db / 10))
    y = x + n
    out = np.stack([y.real, y.imag], axis=-1).astype(np.float32)
    return out, bits.reshape(n_frames, 8 * 46, nbits).astype(np.uint8)
