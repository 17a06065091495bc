"""CPU tests: the oracle against the reference's own artefacts (golden fixtures)."""
import numpy as np
import pytest

from conftest import v1_weights
from oracle import dccn_oracle as orc
from oracle.v1_recipe import v1_frames, constellation


@pytest.mark.parametrize('fixture,nb,cp', [('v1_4mod_cpTrue.npz', 4, True), ('v1_1mod_cpFalse.npz', 1, False)])
def test_v1_known_answer_ber(golden, fixture, nb, cp):
    """Shipped trained checkpoints + BASELINE.md recipe -> the recorded BER (exactly)."""
    w = v1_weights(golden(fixture))
    ka = golden('v1_known_answers.npz')
    tag = '%dmod_cp%s' % (nb, cp)
    for snr in (10, 22) if nb == 4 else (0, 15):
        x, bits = v1_frames(nb, snr, 2000)
        soft = orc.basic_receiver(x, w, nb, 16, use_cp=cp, head='v1', dtype=np.float64)
        _, conf, ber, _ = orc.ber_head(soft, bits)
        expect = float(ka[tag][list(ka['snr']).index(snr)])
        assert conf.sum() == 2000 * 368 * nb
        assert abs(ber - expect) <= 1.5 / conf.sum(), (snr, ber, expect)


def test_v1_known_answers_table_matches_baseline_md(golden):
    """Spot values of BASELINE.md section 2 (computed in the survey session, independently)."""
    ka = golden('v1_known_answers.npz')
    snr = list(ka['snr'])
    assert abs(ka['4mod_cpTrue'][snr.index(10)] - 4.328e-2) < 5e-5
    assert abs(ka['4mod_cpTrue'][snr.index(15)] - 2.195e-3) < 5e-6
    assert abs(ka['1mod_cpTrue'][snr.index(0)] - 4.606e-2) < 5e-5
    assert abs(ka['2mod_cpFalse'][snr.index(5)] - 2.747e-2) < 5e-5
    assert all(ka[k][snr.index(22)] == 0 for k in ka.files if k != 'snr')


def test_fp32_oracle_close_to_fp64(golden):
    w = v1_weights(golden('v1_4mod_cpTrue.npz'))
    x, _ = v1_frames(4, 10, 300)
    s64 = orc.basic_receiver(x, w, 4, 16, head='v1', dtype=np.float64)
    s32 = orc.basic_receiver(x, w, 4, 16, head='v1', dtype=np.float32)
    err = np.abs(s32 - s64)
    assert err.max() < 1e-4 and np.quantile(err, 0.999) < 2e-5


def test_constellation_tables(golden):
    g = golden('const_map.npz')
    for o in (1, 2, 3, 4):
        assert np.array_equal(constellation(o), g['ord%d' % o])


@pytest.mark.parametrize('chan', ['epa', 'eva', 'etu', 'flat', 'custom'])
def test_rayleigh_static_matches_reference(golden, chan):
    g = golden('rayleigh_%s.npz' % chan)
    B = g['tx'].shape[0]
    assert np.allclose(orc.channel_coeff(chan), g['ch_coeff'], rtol=0, atol=0)
    rx, _ = orc.rayleigh_static(g['tx'].reshape(B, -1), g['z'], g['ch_coeff'], np.atleast_2d(g['alpha']))
    assert np.array_equal(rx.reshape(g['rx'].shape).astype(np.float64), g['rx'])


def test_awgn_matches_reference(golden):
    g = golden('awgn.npz')
    out, npw, _ = orc.awgn(g['x'], g['snr'], g['normals'])
    assert np.array_equal(out, g['out'])
    assert npw == float(g['noise_power'])


def test_packed_gemm_equals_conv():
    """SURVEY App. D packing == literal conv3d/reshape/sub formulation of complex.py."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 7, 64, 1, 2))
    k = rng.standard_normal((1, 64, 1, 1, 128)) * 0.1
    b = rng.standard_normal(128) * 0.1
    ref = orc.conv2d_complex(x, k, b, 'valid')                     # [5,7,1,64,2]
    Bp, bp = orc.pack_complex_kernel(k[0, :, 0, 0, :64], k[0, :, 0, 0, 64:], b[:64], b[64:])
    y = x.reshape(35, 128) @ Bp + bp
    assert np.abs(y.reshape(5, 7, 1, 64, 2) - ref).max() < 1e-12


def test_oracle_equals_tf_mirror():
    """NumPy restatement vs the op-for-op torch mirror (padded conv3d) on seeded weights."""
    torch = pytest.importorskip('torch')
    from oracle.tf_mirror import TFMirror
    rng = np.random.default_rng(4)
    for nb, cp, eq in ((2, True, True), (4, False, True), (1, True, False)):
        w = orc.glorot_weights(rng, nb, use_cp=cp, equalizer=eq, bias_scale=0.05, chest_bias=(0.6, -0.4))
        x = (rng.standard_normal((48, 7, 80, 2)) * 0.2).astype(np.float32)
        if eq:
            ref, _, chest = orc.equalized_receiver(x, w, nb, 64, 16, use_cp=cp, dtype=np.float64)
            good = np.abs(chest).reshape(48, -1).min(axis=1) > 2e-2
        else:
            ref = orc.basic_receiver(x, w, nb, 16, use_cp=cp, dtype=np.float64)
            good = np.ones(48, dtype=bool)
        got = TFMirror(w, nb, use_cp=cp, equalizer=eq).forward(x).numpy()
        assert np.quantile(np.abs(got[good] - ref[good]), 0.999) < 2e-4


def test_oracle_equals_tf_mirror_fp64(trained_dev):
    """The same cross-check in float64 on EVERY frame, on seeded and on the GPU-trained variables: the NumPy
    restatement (tap loops, complex arithmetic) and the torch mirror (real zero-padded conv3d + reshape / subtract,
    real arithmetic) share no code and agree to 1e-9 -- what is left unpinned is only TF's own kernels."""
    torch = pytest.importorskip('torch')
    from oracle.tf_mirror import TFMirror
    rng = np.random.default_rng(14)
    cases = [(4, True, trained_dev)]
    for nb, cp in ((2, True), (3, False)):
        cases.append((nb, cp, orc.glorot_weights(rng, nb, use_cp=cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))))
    for nb, cp, w in cases:
        x = (rng.standard_normal((24, 7, 80, 2)) * 0.3).astype(np.float32)
        ref, _, chest = orc.equalized_receiver(x, w, nb, 64, 16, use_cp=cp, dtype=np.float64)
        got = TFMirror(w, nb, use_cp=cp, equalizer=True, dtype=torch.float64).forward(x).numpy()
        # a frame whose channel estimate passes within 1e-6 of zero amplifies even fp64 rounding (no-epsilon divide)
        scale = 1.0 / max(np.abs(chest).min(), 1e-6)
        assert np.abs(got - ref).max() < 1e-12 * scale + 1e-10, (nb, cp, np.abs(got - ref).max())


def test_lean_oracle_equals_literal(trained_dev):
    """oracle/dccn_oracle_lean.py (dense-product form used for the full-size GPU parity run) == the literal op-by-op
    restatement, float64, trained and seeded variables, with and without the cyclic prefix, receiver-only too."""
    from oracle.dccn_oracle_lean import LeanModel, batch_norm_with
    rng = np.random.default_rng(15)
    x = (rng.standard_normal((40, 7, 80, 2)) * 0.3).astype(np.float32)
    z, mean, inv = orc.batch_moment_norm(x, np.float64)
    assert np.array_equal(batch_norm_with(x, mean, inv), z)
    cases = [(4, True, trained_dev)]
    for nb, cp in ((1, True), (2, False)):
        cases.append((nb, cp, orc.glorot_weights(rng, nb, use_cp=cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))))
    for nb, cp, w in cases:
        eq_ref, chest_ref = orc.equalizer_ofdm(z, w, 64, 16, use_cp=cp)
        soft_ref = orc.ofdm_dense_rx(eq_ref, w, nb, 16, use_cp=cp)
        soft, eq, chest = LeanModel(w, nb, use_cp=cp).forward(z)
        scale = 1.0 / max(np.abs(chest_ref).min(), 1e-6)
        assert np.abs(chest - chest_ref).max() < 1e-12
        assert np.abs(eq - eq_ref).max() < 1e-12 * scale and np.abs(soft - soft_ref).max() < 1e-11 * scale
        rx_ref = orc.ofdm_dense_rx(z, w, nb, 16, use_cp=cp)
        rx, _, _ = LeanModel(w, nb, use_cp=cp, equalizer=False).forward(z)
        assert np.abs(rx - rx_ref).max() < 1e-12


def test_trained_fixture_is_a_working_receiver(trained_dev):
    """The committed GPU-trained variables decode: the host transmitter + the oracle's EPA channel + AWGN at 25 dB
    through the oracle give the BER the training run logged on the GPU (tests/golden/..., meta_curve_*)."""
    import os
    from conftest import GOLDEN
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.ofdm import ofdm_tx
    from oracle.dccn_oracle_lean import LeanModel
    meta = np.load(os.path.join(GOLDEN, 'dev_4mod_eq_trained.npz'))
    curve = dict(zip(meta['meta_curve_snr'].tolist(), meta['meta_curve_ber'].tolist()))
    rng = np.random.default_rng(16)
    B, nb = 1500, 4
    o = ofdm_tx(Flags(nbits=nb))
    bits = rng.integers(0, 2, (B, o.frame_size, nb)).astype(np.uint8)
    tx = o.ofdm_tx_frame_np(bits)[0]                                     # complex [B, S*T]
    alpha = np.load(os.path.join(os.path.dirname(GOLDEN), '..', 'dl_ofdm_b200', 'data', 'lte_alpha.npz'))['epa']
    zg = (rng.standard_normal((B, 7)) + 1j * rng.standard_normal((B, 7))) * np.sqrt(0.5)
    faded, _ = orc.rayleigh_static(tx.reshape(B, -1), zg, orc.channel_coeff('epa'), alpha)
    x, _, _ = orc.awgn(faded.reshape(B, 7, 80, 2), np.full((B, 1), 25.0), rng.standard_normal((B, 7, 80, 2)))
    z, _, _ = orc.batch_moment_norm(x.astype(np.float32), np.float64)
    soft, _, chest = LeanModel(trained_dev, nb).forward(z)
    _, _, ber, _ = orc.ber_head(soft, bits)
    assert 0.5 * curve[25.0] < ber < 2.0 * curve[25.0], (ber, curve[25.0])
    assert np.abs(chest).mean() > 0.3                                    # a trained estimate, not one hovering at 0


def test_ber_head_counts():
    soft = np.array([[[[0.7, 0.3]], [[0.5, 0.5]], [[0.2, 0.8]]]])      # [1,3,1,2]
    bits = np.array([[[0], [1], [1]]])
    hard, conf, ber, ce = orc.ber_head(soft, bits)
    assert hard.reshape(-1).tolist() == [0, 0, 1]                       # tie -> index 0
    assert conf.tolist() == [[1, 0], [1, 1]]
    assert abs(ber - 1 / 3) < 1e-12


def test_equalizer_variant_template_reduces_to_equalizer_ofdm():
    """The generic --opt template of the oracle with opt = 0 is exactly equalizer_ofdm (which the TF mirror pins)."""
    import numpy as np
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(4)
    w = orc.glorot_weights(rng, 2, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    z = orc.batch_moment_norm((rng.standard_normal((5, 7, 80, 2)) * 0.3).astype(np.float32))[0]
    for cp in (True, False):
        wc = orc.glorot_weights(np.random.default_rng(5), 2, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), use_cp=cp)
        a, ca = orc.equalizer_ofdm(z, wc, 64, 16, use_cp=cp)
        b, cb = orc.equalizer_variant(z, wc, 0, 64, 16, use_cp=cp)
        assert np.array_equal(a, b) and np.array_equal(ca, cb)
    for opt in (1, 2, 3, 4, 5):
        wv = orc.glorot_weights(np.random.default_rng(opt), 2, equalizer=True, chest_bias=(0.6, -0.4), eq_opt=opt)
        out, chest = orc.equalizer_variant(z, wv, opt, 64, 16)
        assert out.shape == (5, 7, 80, 2) and chest.shape == (5, 7, 64) and np.isfinite(out).all()
        assert np.abs(chest).min() > 0.05


def test_conv2d_vector_matches_literal_conv3d():
    """layers_conv2d_vector (dev/py/complex.py:199-255) restated in the oracle vs a literal conv3d over (length, width,
    IQ) with TF's SAME padding (pad_before = (k-1)//2; the size-2 IQ axis gets 0 before / 1 after) and the reference's
    reshape / slice of the (IQ position, channel) axes."""
    import numpy as np
    import torch
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(12)
    x = rng.standard_normal((3, 7, 64, 1, 2))
    # 'same', one filter: kernel [7,64,2,1,2]
    k = rng.standard_normal((7, 64, 2, 1, 2)) * 0.1
    b = rng.standard_normal(2)
    ref = orc.conv2d_vector(x, k, b, 'same')
    xt = torch.nn.functional.pad(torch.tensor(x[:, :, :, 0, :]).unsqueeze(1), (0, 1, 31, 32, 3, 3))
    y = torch.nn.functional.conv3d(xt, torch.tensor(np.transpose(k[:, :, :, 0, :], (3, 0, 1, 2))).unsqueeze(1),
                                   torch.tensor(b)).numpy()                     # [B, ch, L, W, IQ position]
    merged = np.transpose(y, (0, 2, 3, 4, 1)).reshape(3, 7, 64, 4)             # (IQ position, channel) merged, :243
    assert np.abs(ref[:, :, :, 0, 0] - merged[..., 0]).max() < 1e-12           # conv_re = merged index 0, :245
    assert np.abs(ref[:, :, :, 0, 1] - merged[..., 1]).max() < 1e-12           # conv_im = merged index 1, :246
    # 'valid', (1,K) with K filters: kernel [1,64,2,1,128]
    k2 = rng.standard_normal((1, 64, 2, 1, 128)) * 0.1
    b2 = rng.standard_normal(128)
    ref2 = orc.conv2d_vector(x, k2, b2, 'valid')                               # [B,7,1,64,2]
    y2 = torch.nn.functional.conv3d(torch.tensor(x[:, :, :, 0, :]).unsqueeze(1),
                                    torch.tensor(np.transpose(k2[:, :, :, 0, :], (3, 0, 1, 2))).unsqueeze(1),
                                    torch.tensor(b2)).numpy()                   # [B,128,7,1,1]
    m2 = np.transpose(y2[:, :, :, 0, 0], (0, 2, 1)).reshape(3, 7, 2, 64)       # [.., 1*2, filters], :243
    assert np.abs(ref2[:, :, 0, :, 0] - m2[:, :, 0, :]).max() < 1e-12
    assert np.abs(ref2[:, :, 0, :, 1] - m2[:, :, 1, :]).max() < 1e-12


def test_fp16_split_operand_error_matches_tf32_split():
    """The staged fp16 hi/lo GEMM form (DCCN_F16X3): with the weights pre-scaled by a power of two its
    operand representation error on the shipped 16-QAM checkpoint equals the tf32 pair's (CPU emulation, exact products)."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location('acc_split_emul', os.path.join(ROOT, 'tools', 'acc_split_emul.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.run(n_frames=200, verbose=False)
    tf32, f16 = r['tf32x3'], r['fp16x3, weights x 2^k']
    assert tf32[2] == 0 and f16[2] == 0                      # no hard-bit flips against the fp64 oracle
    assert f16[0] < 1.5 * tf32[0] and f16[1] < 2.0 * tf32[1]
    assert f16[1] < 1e-5                                      # the north-star tolerance on soft outputs
    assert r['fp16x3 unscaled'][0] > 2.0 * f16[0]            # the weight scale is what keeps the lo plane normal


def test_v1_curves_match_baseline_table(golden):
    """tests/golden/v1_curves.npz (bit errors of the 8 shipped checkpoints on the seeded recipe frames, -10..29 dB; the
    known answers of tests/test_gpu_curves.py) reproduces every entry of the table in BASELINE.md section 2."""
    import os
    import re
    cur = golden('v1_curves.npz')
    cols = ['%dmod_cp%s' % (nb, cp) for nb in (1, 2, 3, 4) for cp in (True, False)]
    rows = {}
    for line in open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'BASELINE.md')):
        m = re.match(r'\| (-?\d+) \|(.*)\|', line)
        if m:
            rows[int(m.group(1))] = [float(v) for v in m.group(2).split('|')]
    assert len(rows) == 32
    for snr, vals in rows.items():
        i = int(np.where(cur['snr'] == snr)[0][0])
        for c, v in zip(cols, vals):
            b = cur[c + '_errors'][i] / float(cur[c + '_bits'])
            assert abs(b - v) <= 5.1e-4 * v + 1e-12, (snr, c, b, v)       # the table prints 4 significant digits
    for c in cols:
        assert (cur[c + '_errors'][cur['snr'] >= 22] == 0).all()
