"""ctypes binding of libdccn.so (include/dccn.h).  No torch types cross this boundary.

The library is built in-tree by ``dl_ofdm_b200/build.py`` (``__graft_entry__.build()``).
There is no CPU fallback: if the shared object is missing, or a call fails, this
module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DCCN_LIB') or os.path.join(HERE, 'libdccn.so')   # DCCN_LIB: debug builds (tools/)

PREC_EXACT, PREC_PARITY, PREC_FAST = 0, 1, 2
HEAD_DEV, HEAD_V1 = 0, 1
FWD_NO_NORM, FWD_EQ_ONLY, FWD_SKIP_EQ, FWD_FOLDED = 1, 2, 4, 8
PRECISIONS = {'exact': PREC_EXACT, 'parity': PREC_PARITY, 'fast': PREC_FAST}


class dccn_cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'nfft', 'cp_len', 'nsymbol', 'nfilter', 'nbits', 'use_cp', 'n_data', 'pilot_size',
        'head', 'equalizer', 'precision', 'chunk_frames', 'eq_opt')]


class dccn_train_cfg(C.Structure):
    _fields_ = [('reg_coeff', C.c_float), ('l2', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float),
                ('eps', C.c_float), ('max_batch', C.c_int64), ('mode', C.c_int32)]


TRAIN_EQ, TRAIN_RX = 0, 1      # dccn_train_cfg.mode (include/dccn.h)


class DccnError(RuntimeError):
    pass


_vp, _i64, _i32, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64

# name -> (restype, argtypes); mirrors include/dccn.h one to one
PROTOTYPES = {
    'dccn_abi_version': (C.c_int, []),
    'dccn_last_error': (C.c_char_p, []),
    'dccn_create': (C.c_int, [C.POINTER(dccn_cfg), C.POINTER(_vp)]),
    'dccn_destroy': (None, [_vp]),
    'dccn_workspace_bytes': (C.c_size_t, [_vp]),
    'dccn_set_weight': (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32]),
    'dccn_get_weight': (_i64, [_vp, C.c_char_p, _vp, _i64]),
    'dccn_commit_weights': (C.c_int, [_vp, _vp]),
    'dccn_batch_moments': (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    'dccn_forward': (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    'dccn_forward_host': (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    'dccn_forward_host_begin': (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    'dccn_forward_host_end': (C.c_int, [_vp, _i32, _vp, _vp]),
    'dccn_forward_host_begin_packed': (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    'dccn_cconv2d': (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'dccn_vconv2d': (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'dccn_chan_fir_awgn': (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _u64,
                                     _vp, _vp, _vp]),
    'dccn_chan_fading': (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _i32, C.c_double, C.c_double, _vp,
                                   _u64, _i64, _i64, _i32, _vp, _vp]),
    'dccn_chan_awgn': (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _u64, _vp, _vp]),
    'dccn_ber_accum': (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'dccn_tx_frames': (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp, _i32, _vp, C.c_float, C.c_float, _vp, _vp]),
    'dccn_tx_fade': (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp, _i32, _vp, C.c_float, C.c_float, _vp, _vp, _i32, _i32, _vp,
                               _u64, _i32, _vp, _vp, _vp]),
    'dccn_bit_source': (C.c_int, [_vp, _i64, _u64, _vp]),
    'dccn_crc32c': (C.c_uint32, [_vp, C.c_size_t, C.c_uint32]),
    'dccn_monitors': (C.c_int, [_vp, _vp, _i64, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    'dccn_forward_monitors': (C.c_int, [_vp, _vp, _vp, _i32]),
    'dccn_debug_tma_rate': (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    'dccn_debug_mma_rate': (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    'dccn_train_init': (C.c_int, [_vp, C.POINTER(dccn_train_cfg), _vp]),
    'dccn_train_step': (C.c_int, [_vp, _vp, _i64, _vp, C.c_float, _i32, _vp, _vp, _i32, _vp]),
    'dccn_train_get_grad': (_i64, [_vp, C.c_char_p, _vp, _i64]),
    'dccn_train_global_step': (_i64, [_vp]),
    'dccn_train_set_global_step': (C.c_int, [_vp, _i64]),
    'dccn_launch_count': (_i64, []),
    'dccn_profile_enable': (C.c_int, [_vp, _i32]),
    'dccn_profile_collect': (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i64), _i32]),
    'dccn_profile_slot_name': (C.c_char_p, [_i32]),
}

_lib = None


def load():
    """Load libdccn.so and attach the prototypes (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DccnError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc is not None and rc < 0:
        raise DccnError(load().dccn_last_error().decode(errors='replace'))
    return rc
