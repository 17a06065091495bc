"""dev/py/util.py surface: bit_source, ber_calc (BER arithmetic of ber_tensor)."""
from __future__ import annotations

import numpy as np


def bit_source(nbits, frame_size, msg_length, rng=None):
    """Uniform random bits [msg_length, frame_size, nbits] (dev/py/util.py:25-34)."""
    rng = np.random if rng is None else rng
    gen = rng.integers if hasattr(rng, 'integers') else rng.randint
    return gen(0, 2, (int(msg_length), int(frame_size), int(nbits)))


def ber_calc(conf_matrix):
    """(errors / total) of a 2x2 confusion matrix (dev/py/util.py:37-41, :44-48)."""
    conf_matrix = np.asarray(conf_matrix)
    assert conf_matrix.shape == (2, 2)
    total = float(conf_matrix.sum())
    return float(conf_matrix[0][1] + conf_matrix[1][0]) / total
