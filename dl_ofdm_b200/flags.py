"""The reference's ``tf.app.flags`` surface without TensorFlow.

Flag names, defaults and types are those of dev/py/ofdmreceiver_np_mp.py:33-58
(a superset of dev/py/ofdmreceiver_np.py:30-53); ``parse_flags(argv)`` accepts the
same ``--name=value`` strings that ``run_local_ofdm.py`` builds (dev/py/run_local_ofdm.py:74-78).
The two drivers differ in four defaults (``parse_flags(argv, driver='np')`` applies ofdmreceiver_np.py's:
SNR 3, max_epoch_num 1000, early_stop 100, load_model False -- ofdmreceiver_np.py:36-48).
Two extra flags select B200 execution: --precision, --frames (0 = the driver's own test-set size).
"""
from __future__ import annotations

import argparse

_DEFS = [
    ('save_dir', str, './output/'), ('nbits', int, 1), ('msg_length', int, 100800),
    ('batch_size', int, 512), ('max_epoch_num', int, 5000), ('seed', int, 1), ('nfft', int, 64),
    ('nsymbol', int, 7), ('npilot', int, 8), ('nguard', int, 8), ('nfilter', int, 80),
    ('SNR', float, 30.0), ('SNR2', float, 30.0), ('early_stop', int, 400), ('ofdm', bool, True),
    ('pilot', str, 'lte'), ('channel', str, 'EPA'), ('cp', bool, True), ('longcp', bool, True),
    ('load_model', bool, True), ('split', float, 1.0), ('token', str, 'OFDM'), ('opt', int, 3),
    ('mobile', bool, False), ('init_learning', float, 0.001), ('test', bool, False),
    # B200 additions
    ('precision', str, 'parity'), ('frames', int, 0),
]
# dev/py/ofdmreceiver_np.py:36-48 (the basic-receiver driver) where it differs from ofdmreceiver_np_mp.py
_NP_DEFS = {'SNR': 3.0, 'max_epoch_num': 1000, 'early_stop': 100, 'load_model': False}


def _to_bool(v):
    if isinstance(v, bool):
        return v
    return str(v).lower() in ('1', 'true', 't', 'yes', 'y')


class Flags:
    """Attribute bag with the reference defaults; keyword arguments override."""

    def __init__(self, **kw):
        for name, _, default in _DEFS:
            setattr(self, name, default)
        self.nfilter = 64                       # the launcher always passes nFFT (run_local_ofdm.py:62)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError('unknown flag %r' % k)
            setattr(self, k, v)

    def copy(self, **kw):
        f = Flags()
        f.__dict__.update(self.__dict__)
        f.__dict__.update(kw)
        return f


def parse_flags(argv=None, driver='mp'):
    p = argparse.ArgumentParser(allow_abbrev=False)
    for name, typ, default in _DEFS:
        if driver == 'np' and name in _NP_DEFS:
            default = _NP_DEFS[name]
        if typ is bool:
            p.add_argument('--' + name, type=_to_bool, default=default, nargs='?', const=True)
        else:
            p.add_argument('--' + name, type=typ, default=default)
    ns, _ = p.parse_known_args(argv)
    f = Flags()
    f.__dict__.update(vars(ns))
    return f
