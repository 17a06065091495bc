"""CPU tests of the host side: geometry / transmitter vs reference fixtures, flags,
checkpoint I/O, and that libdccn.so exports every symbol include/dccn.h declares."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, v1_weights
from dl_ofdm_b200 import tfbundle
from dl_ofdm_b200.flags import Flags, parse_flags
from dl_ofdm_b200.ofdm import const_map, get_lte_dl_cfg, ofdm_tx


def test_const_map(golden):
    g = golden('const_map.npz')
    for o in (1, 2, 3, 4):
        assert np.array_equal(const_map(o), g['ord%d' % o])


@pytest.mark.parametrize('tag,nb,pilot,ns,lcp', [
    ('lte_1b', 1, 'lte', 7, True), ('lte_2b', 2, 'lte', 7, True), ('lte_3b', 3, 'lte', 7, True),
    ('lte_4b', 4, 'lte', 7, True), ('scattered_4b', 4, 'scattered', 8, True),
    ('lte_2b_shortcp', 2, 'lte', 7, False)])
def test_ofdm_tx_matches_reference(golden, tag, nb, pilot, ns, lcp):
    g = golden('ofdm_tx_%s.npz' % tag)
    o = ofdm_tx(Flags(nbits=nb, pilot=pilot, nsymbol=ns, longcp=lcp))
    K, CP, P, G, DC, fs, ps, nrb = (int(v) for v in g['meta'])
    assert (o.K, o.CP, o.P, o.G, o.DC, o.frame_size, o.pilot_size, o.nRB) == (K, CP, P, G, DC, fs, ps, nrb)
    assert o.Fs == float(g['Fs'])
    for name in ('dataSc', 'pilotSc', 'guardSc', 'effecCarriers', 'pilotCarriers', 'dataCarriers'):
        assert np.array_equal(getattr(o, name), g[name]), name
    cpx, real, pil = o.ofdm_tx_frame_np(g['bits'])
    assert np.array_equal(cpx, g['cpx'])
    assert np.array_equal(real.astype(np.float64), g['real'])
    assert pil.shape == (6, ns, P, 2)


def test_lte_cfg_and_errors():
    assert get_lte_dl_cfg(64) == (0.96e6, 4)
    with pytest.raises(AssertionError):
        get_lte_dl_cfg(100)
    with pytest.raises(ValueError):
        ofdm_tx(Flags(pilot='nope'))
    o = ofdm_tx(Flags())
    with pytest.raises(AssertionError):
        o.ofdm_tx_frame_np(np.zeros((2, 17, 1), dtype=np.uint8))      # wrong frame_size (ofdm.py:343)
    cpx, real, _ = o.ofdm_tx_frame_np(np.zeros((0, o.frame_size, 1), dtype=np.uint8))   # empty batch
    assert cpx.shape == (0, 7, 80) and real.shape == (0, 7, 80, 2)


def test_flags_surface():
    f = parse_flags(['--nbits=4', '--cp=False', '--channel=mixRayleigh', '--opt=0', '--mobile=True',
                     '--token=OFDM_Dense3', '--batch_size=512', '--nfilter=64', '--longcp=True'])
    assert (f.nbits, f.cp, f.channel, f.opt, f.mobile, f.token, f.nfilter) == \
        (4, False, 'mixRayleigh', 0, True, 'OFDM_Dense3', 64)
    assert f.nsymbol == 7 and f.pilot == 'lte' and f.SNR == 30.0
    with pytest.raises(AttributeError):
        Flags(bogus=1)


def test_tfbundle_roundtrip(tmp_path, golden):
    w = v1_weights(golden('v1_1mod_cpFalse.npz'))
    w['global_step'] = np.float32(1336.0).reshape(())
    prefix = str(tmp_path / 'ckpt' / 'model')
    tfbundle.write_checkpoint(prefix, w)
    idx = tfbundle.read_index(prefix)
    assert set(idx) == set(w)
    assert idx['demodulation/dense/kernel'][1] == (1024, 736)
    back = tfbundle.read_checkpoint(prefix)
    for k in w:
        assert np.array_equal(back[k], w[k]), k
    assert tfbundle.crc32c(b'123456789') == 0xE3069283


def test_libdccn_exports_header_symbols(libdccn):
    """dlopen works without a GPU and every function declared in include/dccn.h is exported."""
    from dl_ofdm_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'dccn.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(dccn_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 14
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(libdccn, name)
    assert libdccn.dccn_abi_version() == 3


def test_no_gpu_fails_loudly(libdccn):
    """No silent CPU fallback: without a CUDA device the product raises."""
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from dl_ofdm_b200.engine import DCCN, DccnError
    with pytest.raises(DccnError):
        DCCN(nbits=1)
    import ctypes as C
    from dl_ofdm_b200._lib import dccn_cfg
    cfg = dccn_cfg(nfft=64, cp_len=16, nsymbol=7, nfilter=64, nbits=1, use_cp=1, n_data=320, pilot_size=16)
    h = C.c_void_p()
    assert libdccn.dccn_create(C.byref(cfg), C.byref(h)) < 0
    assert b'no CUDA device' in libdccn.dccn_last_error()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under dl_ofdm_b200/ may import it."""
    pkg = os.path.join(ROOT, 'dl_ofdm_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_equalizer_graph_tables_agree():
    """The host's --opt wiring table (variable names, shapes, detection) == the oracle's independent copy."""
    from dl_ofdm_b200 import init
    from oracle import dccn_oracle as orc
    for opt in (0, 1, 2, 3, 4, 5, 7):
        w = init.equalizer_variables(np.random.default_rng(0), opt=opt)
        wo = {k: v for k, v in orc.glorot_weights(np.random.default_rng(0), 2, equalizer=True, eq_opt=opt).items()
              if k.startswith('Equalizer/')}
        assert {k: v.shape for k, v in w.items()} == {k: v.shape for k, v in wo.items()}, opt
        assert init.detect_eq_opt(w) == opt
    # dense_8 only exists in equalizer_dnnE (dense, dense_1, ..., dense_8; no conv3d)
    assert [n for _, n in init.eq_layer_roles(3)][-1] == 'dense_8'
    assert [n for _, n in init.eq_layer_roles(2)] == ['dense', 'conv3d', 'dense_1', 'dense_2', 'dense_3']
