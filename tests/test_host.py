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


def _meta_digest_helpers():
    import hashlib

    def node_key(n):
        attrs = ''.join('%s=%s;' % (k, n.attr[k].SerializeToString(deterministic=True).hex()) for k in sorted(n.attr.keys()))
        return '%s|%s|%s|%s|%s' % (n.name, n.op, ','.join(n.input), n.device, attrs)

    def needed(nodes, fetches):
        st = list(fetches) + [n for n in nodes if n.endswith('/Assign') and 'Adam' not in n and '_power' not in n
                              and not n.startswith('save/')]
        seen = set()
        while st:
            n = st.pop().lstrip('^').split(':')[0]
            if n not in seen:
                seen.add(n)
                st.extend(nodes[n].input)
        return sorted(seen)

    def digest(nodes, names):
        h = hashlib.sha256()
        for n in names:
            h.update(node_key(nodes[n]).encode())
            h.update(b'\n')
        return h.hexdigest()
    return node_key, needed, digest


@pytest.mark.parametrize('nb', [1, 2, 3, 4])
@pytest.mark.parametrize('cp', [True, False])
def test_meta_graph_matches_shipped_v1(nb, cp):
    """The `.meta` emitter (tfmeta / tfgraph, no TensorFlow) rebuilds the graph of every checkpoint the reference ships
    (test_v1/model/*.meta, written by TF 1.10.1): all 452 / 455 nodes reachable from the named fetches of
    dev/py/model.py:51-72 and from the model variables' initializers are identical in name, op, inputs, device and
    EVERY attribute (tests/golden/v1_meta_digest.json, made by oracle/make_meta_digest.py from the shipped files)."""
    import hashlib
    import json
    from dl_ofdm_b200 import tfmeta
    node_key, needed, digest = _meta_digest_helpers()
    gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'v1_meta_digest.json')))
    name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, 3 * nb, cp)
    g = tfmeta.build_receiver_graph(nb, 8, 80, 64, 8 * 46, 64, cp, 'v1')
    m = tfmeta.meta_graph(g, '1.10.1', 'v1.10.1-0-g4dcfddc5d1')
    nodes = {n.name: n for n in m.graph_def.node}
    names = needed(nodes, tfmeta.FETCHES)
    assert len(names) == gold[name]['nodes']
    assert hashlib.sha256('\n'.join(names).encode()).hexdigest() == gold[name]['names_sha256']
    if 'per_node' in gold[name]:
        bad = [n for n in names if hashlib.sha1(node_key(nodes[n]).encode()).hexdigest()[:12] != gold[name]['per_node'][n]]
        assert not bad, bad[:10]
    assert digest(nodes, names) == gold[name]['sha256']
    # saver + collections the reference's import_meta_graph / restore use
    assert m.saver_def.filename_tensor_name == 'save/Const:0' and m.saver_def.restore_op_name == 'save/restore_all'
    assert m.saver_def.save_tensor_name == 'save/control_dependency:0' and m.saver_def.version == 2
    from tensorboard.compat.proto import variable_pb2
    vs = []
    for b in m.collection_def['variables'].bytes_list.value:
        v = variable_pb2.VariableDef()
        v.ParseFromString(b)
        vs.append(v)
    assert [v.variable_name for v in vs] == ['fft_like/conv3d/kernel:0', 'fft_like/conv3d/bias:0', 'demodulation/dense/kernel:0',
                                             'demodulation/dense/bias:0', 'demodulation/conv2d/kernel:0', 'demodulation/conv2d/bias:0',
                                             'demodulation/conv2d_1/kernel:0', 'demodulation/conv2d_1/bias:0',
                                             'demodulation/dense_1/kernel:0', 'demodulation/dense_1/bias:0', 'global_step:0']
    assert all(v.snapshot_name == v.variable_name[:-2] + '/read:0' and v.initializer_name == v.variable_name[:-2] + '/Assign' for v in vs)
    assert len(m.collection_def['trainable_variables'].bytes_list.value) == 10
    assert list(m.collection_def['regularization_losses'].node_list.value) == [
        'receiver/demodulation/%s/Regularizer/l2_regularizer:0' % t for t in ('dense/kernel', 'dense/bias', 'dense_1/kernel', 'dense_1/bias')]


def test_save_model_writes_meta_for_the_bundle(tmp_path):
    """save_model on a basic receiver writes .index / .data AND a .meta whose saver covers exactly the bundle's tensors and
    whose graph holds every tensor name dev/py/model.py:51-72 fetches after import_meta_graph (dev architecture)."""
    from tensorboard.compat.proto import meta_graph_pb2
    from dl_ofdm_b200 import init, tfbundle
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.model import save_model
    from dl_ofdm_b200.ofdm import ofdm_tx
    for nb, cp in ((4, True), (1, False)):
        fl = Flags(nbits=nb, cp=cp)
        o = ofdm_tx(fl)
        w = init.receiver_variables(np.random.default_rng(0), nb, use_cp=cp)
        prefix = str(tmp_path / ('rx%d' % nb))
        save_model(prefix, w, global_step=123, FLAGS=fl, ofdmobj=o)
        m = meta_graph_pb2.MetaGraphDef()
        m.ParseFromString(open(prefix + '.meta', 'rb').read())
        nodes = {n.name: n for n in m.graph_def.node}
        for t in ('bits_in', 'tx_ofdm', 'input', 'output', 'cost', 'log_ber', 'linear_ber', 'conf_matrix', 'tx_power',
                  'noise_power', 'iq_rx', 'iq_tx', 'ce_mean', 'SNR'):
            assert t in nodes, t
        saved = [s.decode() for s in nodes['save/SaveV2/tensor_names'].attr['value'].tensor.string_val]
        assert saved == sorted(tfbundle.read_checkpoint(prefix).keys())
        assert list(nodes['save/SaveV2'].input[3:]) == saved
        shp = lambda n: [d.size for d in nodes[n].attr['_output_shapes'].list.shape[0].dim]      # noqa: E731
        assert shp('tx_ofdm') == [-1, 7, 80, 2] and shp('output') == [-1, 320, nb, 2] and shp('bits_in') == [-1, 320, nb]
        T = 80 if cp else 64
        assert shp('fft_like/conv3d/kernel') == [1, T, 1, T, 128] == list(w['fft_like/conv3d/kernel'].shape)
        assert shp('demodulation/dense/kernel') == [896, 640]


def test_expert_baselines_restated():
    """dl_ofdm_b200/baselines.py (NumPy restatement of the estimators of dev/m/OFDM_Benchmark_dev.m) on frames from the host
    transmitter + the oracle's Rayleigh / AWGN: the interpolation matrices reproduce their defining properties, and the
    BER ordering perfect CSI <= ideal LMMSE <= LS-spline <= LS-linear holds on a fading channel, all tending to 0 on AWGN."""
    from dl_ofdm_b200.baselines import ClassicReceiver, spline_matrix, linear_matrix
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.ofdm import ofdm_tx
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(5)
    # interpolators: exact at the nodes; the linear one reproduces any plane, the spline one any constant... at the nodes
    px, py = rng.uniform(1, 64, 16), rng.uniform(1, 7, 16)
    assert np.abs(spline_matrix(px, py, px, py) - np.eye(16)).max() < 1e-8
    assert np.abs(linear_matrix(px, py, px, py) - np.eye(16)).max() < 1e-9
    qx, qy = rng.uniform(10, 50, 40), rng.uniform(2, 6, 40)
    plane = lambda x, y: 0.3 * x - 1.7 * y + 4.0                                   # noqa: E731
    assert np.abs(linear_matrix(px, py, qx, qy) @ plane(px, py) - plane(qx, qy)).max() < 1e-9
    nb, B = 2, 1500
    o = ofdm_tx(Flags(nbits=nb))
    cr = ClassicReceiver(o, nb)
    bits = rng.integers(0, 2, (B, o.frame_size, nb)).astype(np.uint8)
    tx = o.ofdm_tx_frame_np(bits)[0].reshape(B, -1)
    alpha = np.load(os.path.join(ROOT, 'dl_ofdm_b200', 'data', 'lte_alpha.npz'))['etu']
    coeff = orc.channel_coeff('etu')
    z = (rng.standard_normal((B, len(coeff))) + 1j * rng.standard_normal((B, len(coeff)))) * np.sqrt(.5)
    faded, g = orc.rayleigh_static(tx, z, coeff, alpha)
    x, _, _ = orc.awgn(faded.reshape(B, 7, 80, 2), np.full((B, 1), 20.0), rng.standard_normal((B, 7, 80, 2)))
    ber = {e: cr.ber(x, bits, e, 20.0, g) for e in ('perfect', 'lmmse', 'ls_spline', 'ls_linear')}
    assert ber['perfect'] < ber['lmmse'] < ber['ls_spline'] < ber['ls_linear'] < 0.06, ber
    assert 0.004 < ber['perfect'] < 0.02                                          # Rayleigh QPSK at 20 dB: ~1 / (4 snr) + ISI floor
    xa, _, _ = orc.awgn(np.stack([tx.real, tx.imag], -1).reshape(B, 7, 80, 2), np.full((B, 1), 20.0), rng.standard_normal((B, 7, 80, 2)))
    g1 = np.ones((B, 1), dtype=np.complex128)
    assert all(cr.ber(xa, bits, e, 20.0, g1) == 0.0 for e in ('perfect', 'ls_spline', 'ls_linear'))
