import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without CUDA skips the GPU parity tests instead of erroring in their fixtures."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200)')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def v1_weights(npz):
    """Expand a committed v1 fixture (live centre tap only) back to TF variable layout."""
    w = {}
    for k in npz.files:
        if not k.startswith('meta_'):
            w[k.replace('.', '/')] = npz[k]
    shape = tuple(int(v) for v in w.pop('fft_like/conv3d/kernel_shape'))
    centre = w.pop('fft_like/conv3d/kernel_center')
    full = np.zeros(shape, dtype=np.float32)
    full[0, (shape[1] - 1) // 2, 0] = centre
    w['fft_like/conv3d/kernel'] = full
    return w


dev_weights = v1_weights        # same fixture format (tools/train_fixture.py): live centre tap of fft_like only


@pytest.fixture(scope='session')
def trained_dev():
    """16-QAM receiver + equalizer_ofdm TRAINED on the GPU with the reference's two-phase schedule
    (tools/train_fixture.py: AWGN 20 dB, then EPA transfer learning from TF's zero-bias init) -- TF variable names."""
    return dev_weights(np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz'), allow_pickle=False))


@pytest.fixture(scope='session')
def libdccn():
    import __graft_entry__ as g
    g.build()
    from dl_ofdm_b200 import _lib
    return _lib.load()
