"""GPU tests of the alternative forms of the same kernels: the fp16 hi/lo GEMM form (default since round 2) against
the tf32 hi/lo form of round 1 (DCCN_F16X3=0, still used by the training step) and against the oracle; the dynamic
operand scale that keeps the fp16 form inside fp16's range; the packed-label host entry point; the 8x8-IDFT transmitter
against the generic DFT kernel.  Same oracle, same bounds as tests/test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

from conftest import v1_weights

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _helpers():
    import test_gpu_parity as tp
    return tp._cuda, tp._check_soft


@pytest.mark.parametrize('fixture,nb,cp', [('v1_4mod_cpTrue.npz', 4, True), ('v1_1mod_cpFalse.npz', 1, False)])
@pytest.mark.parametrize('f16', ['0', '1'])
def test_gemm_forms_v1_checkpoint_receiver(libdccn, golden, monkeypatch, fixture, nb, cp, f16):
    monkeypatch.setenv('DCCN_F16X3', f16)
    _cuda, _check_soft = _helpers()
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    from oracle.v1_recipe import v1_frames
    w = v1_weights(golden(fixture))
    snr = 10 if nb == 4 else 0
    x, bits = v1_frames(nb, snr, 700)
    soft_ref = orc.basic_receiver(x, w, nb, 16, use_cp=cp, head='v1', dtype=np.float64)
    soft_ref32 = orc.basic_receiver(x, w, nb, 16, use_cp=cp, head='v1', dtype=np.float32)
    _, conf_ref, _, ce_ref = orc.ber_head(soft_ref, bits)
    m = DCCN(nbits=nb, nsymbol=8, n_data=368, use_cp=cp, head='v1', precision='parity', chunk_frames=256)
    m.load_weights(w)
    out = m.forward(_cuda(x), _cuda(bits))
    torch.cuda.synchronize()
    soft, hard = out['soft'].cpu().numpy(), out['hard'].cpu().numpy()
    flips = _check_soft(soft, soft_ref, hard, soft_ref32)
    conf = out['conf'].cpu().numpy()
    assert conf.sum() == bits.size and np.abs(conf - conf_ref).sum() <= 2 * flips
    assert abs(float(out['ce_sum'].cpu()[0]) / bits.size - ce_ref) < 1e-5
    m.close()


@pytest.mark.parametrize('cp', [True, False])
@pytest.mark.parametrize('folded', [0, 1])
def test_gemm_forms_equalizer_agree(libdccn, monkeypatch, cp, folded):
    """eq + rx on seeded weights (all twelve GEMMs: N = 32, K = 32, K = 136 tails, the banded Toeplitz operand):
    the fp16 form against the fp64 oracle with the default mode's bounds, and against the default mode itself."""
    _cuda, _ = _helpers()
    from dl_ofdm_b200 import _lib
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(78)
    nb, B = 4, 300                                    # ragged: 300 frames = 2100 symbol rows, not multiples of 128
    w = orc.glorot_weights(rng, nb, use_cp=cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.2).astype(np.float32)
    bits = rng.integers(0, 2, (B, 320, nb)).astype(np.uint8)
    soft_ref, eq_ref, chest_ref = orc.equalized_receiver(x, w, nb, 64, 16, use_cp=cp, dtype=np.float64)
    flags = _lib.FWD_FOLDED if folded else 0
    outs = {}
    for f16 in ('0', '1'):
        monkeypatch.setenv('DCCN_F16X3', f16)
        m = DCCN(nbits=nb, use_cp=cp, equalizer=True, precision='parity', chunk_frames=128)
        m.load_weights(w)
        o = m.forward(_cuda(x), _cuda(bits), flags=flags)
        torch.cuda.synchronize()
        outs[f16] = (o['soft'].cpu().numpy(), o['hard'].cpu().numpy(), o['conf'].cpu().numpy())
        m.close()
    good = np.abs(chest_ref).reshape(B, -1).min(axis=1) > 2e-2      # phase equaliser divides by |chest| (no epsilon)
    assert good.sum() > B // 4
    for f16 in ('0', '1'):
        soft, hard, conf = outs[f16]
        assert np.isfinite(soft).all() and conf.sum() == bits.size
        err = np.abs(soft[good] - soft_ref[good])
        assert np.quantile(err, 0.999) < 2e-4, (f16, np.quantile(err, 0.999))
        hard_ref = (soft_ref[..., 1] > soft_ref[..., 0]).astype(np.uint8)
        decided = (np.abs(soft_ref[..., 1] - soft_ref[..., 0]) >= 1e-3) & good[:, None, None]
        assert np.array_equal(hard[decided], hard_ref[decided]), f16
    d = np.abs(outs['1'][0][good] - outs['0'][0][good])
    assert np.quantile(d, 0.999) < 2e-4


def test_f16_form_survives_fp16_range(libdccn, monkeypatch):
    """Activations far outside fp16's range (|x| up to ~1e6, and ~1e-6) with the first layer's weights scaled the other
    way: mathematically the same network, so the outputs must equal the nominal ones -- the per-pass operand scale
    (max |activation| recorded by the producing kernel) keeps the fp16 hi/lo operands finite and normal."""
    from dl_ofdm_b200 import _lib
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(11)
    nb, B = 2, 200
    w = orc.glorot_weights(rng, nb, equalizer=False, bias_scale=0.05)
    z = (rng.standard_normal((B, 7, 80, 2)) * 0.7).astype(np.float32)
    soft_ref = orc.ofdm_dense_rx(z.astype(np.float64), w, nb, 16)
    for scale in (1.0, 2.0 ** 20, 2.0 ** -20):
        ws = dict(w)
        ws['fft_like/conv3d/kernel'] = (w['fft_like/conv3d/kernel'] / scale).astype(np.float32)   # power of two: exact
        m = DCCN(nbits=nb, equalizer=False, precision='parity')
        m.load_weights(ws)
        o = m.forward(torch.as_tensor(z * np.float32(scale)).cuda(), flags=_lib.FWD_NO_NORM)
        soft = o['soft'].cpu().numpy()
        assert np.isfinite(soft).all(), scale
        err = np.abs(soft - soft_ref)
        assert np.quantile(err, 0.999) <= 1e-5 and err.max() <= 5e-5, (scale, np.quantile(err, 0.999), err.max())
        m.close()


def test_f16_form_chunk_invariance_full_size(libdccn, monkeypatch):
    """65 536 frames in one pass == the same frames in 4 096-frame passes (the batch moments are taken over the whole
    batch first, so the internal chunking must not change a single hard bit) -- the regression test that found the
    A-slot release race of the tf32 form (test_gpu_parity.py::test_full_size_properties)."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(5)
    nb, B = 4, 65536
    w = orc.glorot_weights(rng, nb, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    x = torch.randn((B, 7, 80, 2), device='cuda') * 0.2
    bits = torch.randint(0, 2, (B, 320, nb), device='cuda', dtype=torch.uint8)
    res = []
    for chunk in (0, 4096):
        m = DCCN(nbits=nb, equalizer=True, precision='parity', chunk_frames=chunk)
        m.load_weights(w)
        o = m.forward(x, bits, want_soft=False)
        torch.cuda.synchronize()
        res.append((o['hard'].clone(), o['conf'].cpu().numpy()))
        m.close()
    assert res[0][1].sum() == B * 320 * nb
    assert torch.equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize('cp,nb,B,want_eq', [(True, 4, 900, True), (False, 2, 333, True), (True, 1, 128, False), (True, 4, 4099, False)])
def test_chained_kernels_equal_layer_by_layer(libdccn, trained_dev, monkeypatch, cp, nb, B, want_eq):
    """csrc/chain.cu keeps the layer-by-layer arithmetic (same fp32 sums, same order, bias in fp32, then the fp16 split): the
    chained schedule must reproduce the layer-by-layer one -- soft outputs to 2e-7 (in practice bit for bit; the per-row
    operand scale can only differ from the per-pass one below fp16's normal range), identical hard bits, and the dense_5
    side output (`eq`, the aux store of the tail chain) -- on ragged batches, both CP modes, several passes per call."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(21)
    if cp and nb == 4:
        w = trained_dev
        from test_gpu_parity import _config3_frames
        m0 = DCCN(nbits=nb, equalizer=True, precision='parity', chunk_frames=2048)
        x, bits = _config3_frames(m0, B, 15.0, seed=5)
        m0.close()
    else:
        w = orc.glorot_weights(rng, nb, use_cp=cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
        x = torch.as_tensor((rng.standard_normal((B, 7, 80, 2)) * 0.2).astype(np.float32)).cuda()
        bits = torch.as_tensor(rng.integers(0, 2, (B, 320, nb)).astype(np.uint8)).cuda()
    outs = {}
    for chain in ('0', '1'):
        monkeypatch.setenv('DCCN_CHAIN', chain)
        m = DCCN(nbits=nb, use_cp=cp, equalizer=True, precision='parity', chunk_frames=2048)
        m.load_weights(w)
        m.profile(True)
        o = m.forward(x, bits, want_eq=want_eq)
        torch.cuda.synchronize()
        slots = set(m.profile_collect())
        m.profile(False)
        assert ('eq_chain_tail' in slots) == (chain == '1') and ('eq_dense5' in slots) == (chain == '0'), slots
        outs[chain] = {k: o[k].clone() for k in (('soft', 'hard', 'conf', 'eq') if want_eq else ('soft', 'hard', 'conf'))}
        m.close()
    assert (outs['0']['soft'] - outs['1']['soft']).abs().max().item() <= 2e-7
    assert torch.equal(outs['0']['hard'], outs['1']['hard']) and torch.equal(outs['0']['conf'], outs['1']['conf'])
    if want_eq:
        scale = max(1.0, outs['0']['eq'].abs().max().item())
        assert (outs['0']['eq'] - outs['1']['eq']).abs().max().item() <= 2e-7 * scale


def test_run_to_run_determinism_full_size(libdccn, monkeypatch):
    """The same 65 536-frame pass twice (persistent multi-tile kernels, every pipeline ring wrapping hundreds of times):
    soft outputs, channel estimate and equaliser output must be bit-identical, with the chained per-symbol kernels and with
    the layer-by-layer schedule.  (Regression test: a register-rebalancing experiment in the GEMM kernel -- setmaxnreg, see
    gemm_tc.cuh DCCN_TC_REGBAL -- passed every parity test at small sizes and differed from run to run in 0.03 % of the
    channel estimates at this size.)"""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(11)
    nb, B = 4, 65536
    w = orc.glorot_weights(rng, nb, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    g = torch.Generator(device='cuda').manual_seed(12)
    x = torch.randn((B, 7, 80, 2), generator=g, device='cuda') * 0.2
    for chain in ('1', '0'):
        monkeypatch.setenv('DCCN_CHAIN', chain)
        m = DCCN(nbits=nb, equalizer=True, precision='parity')
        m.load_weights(w)
        outs = []
        for _ in range(2):
            o = m.forward(x, None, want_eq=True, want_chest=True)
            torch.cuda.synchronize()
            outs.append({k: o[k].clone() for k in ('soft', 'hard', 'eq', 'chest')})
        for k in ('chest', 'eq', 'soft', 'hard'):
            assert torch.equal(outs[0][k], outs[1][k]), (chain, k, int((outs[0][k] != outs[1][k]).sum()))
        m.close()


def test_packed_labels_host_entry(libdccn):
    """dccn_forward_host_begin_packed (labels 8 per byte) == dccn_forward_host_begin (one uint8 per label): same
    confusion matrix and loss, through both slots; ragged batch."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(9)
    nb, B = 4, 333
    w = orc.glorot_weights(rng, nb, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    x = torch.as_tensor((rng.standard_normal((B, 7, 80, 2)) * 0.2).astype(np.float32)).pin_memory()
    bits = rng.integers(0, 2, (B, 320, nb)).astype(np.uint8)
    packed = torch.as_tensor(np.packbits(bits.reshape(-1), bitorder='little')).pin_memory()
    bh = torch.as_tensor(bits).pin_memory()
    m = DCCN(nbits=nb, equalizer=True, precision='parity')
    m.load_weights(w)
    conf_ref, ce_ref, _ = m.forward_host(x, bh)
    assert conf_ref.sum() == bits.size
    for slot in (0, 1, 0):
        m.forward_host_begin_packed(slot, x, packed)
        conf, ce = m.forward_host_end(slot)
        assert np.array_equal(conf, conf_ref) and ce == ce_ref
    m.close()


@pytest.mark.parametrize('tag,nb,pilot,nsym', [('lte_4b', 4, 'lte', 7), ('lte_1b', 1, 'lte', 7),
                                              ('scattered_4b', 4, 'scattered', 8)])
def test_tx_v2_golden(libdccn, golden, monkeypatch, tag, nb, pilot, nsym):
    """The 8 x 8 IDFT transmitter kernel (default for nfft = 64) and the generic K-point DFT kernel (DCCN_TX_V2=0) against
    the reference transmitter's own output, twice (the subcarrier map is rebuilt per call), and against each other."""
    _cuda, _ = _helpers()
    from dl_ofdm_b200.engine import DCCN
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.ofdm import ofdm_tx, const_map
    g = golden('ofdm_tx_%s.npz' % tag)
    fl = Flags(nbits=nb, pilot=pilot, nsymbol=nsym)
    o = ofdm_tx(fl)
    bits = _cuda(g['bits'])
    res = {}
    for v2 in ('0', '1'):
        monkeypatch.setenv('DCCN_TX_V2', v2)
        m = DCCN(nbits=nb, nsymbol=nsym, n_data=o.frame_size, precision='exact')
        a = m.transmit(bits, o, const_map(nb)).cpu().numpy()
        b = m.transmit(bits, o, const_map(nb)).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.abs(a - g['real']).max() < 5e-7
        res[v2] = a
        m.close()
    assert np.abs(res['0'] - res['1']).max() < 1.2e-7          # both round an fp64 result: at most one fp32 ulp apart


@pytest.mark.parametrize('chan', ['EPA', 'ETU', 'Flat', 'AWGN'])
@pytest.mark.parametrize('nb', [4, 1])
def test_tx_fade_equals_separate_kernels(libdccn, chan, nb):
    """dccn_tx_fade (transmitter + static Rayleigh FIR in one kernel, frame kept in shared memory) == dccn_tx_frames followed
    by dccn_chan_fading, bit for bit: faded frames, the optional copy of the transmitted frames, and (through the batch power)
    the AWGN output; Philox path gains and injected ones; ragged batch."""
    from dl_ofdm_b200.engine import DCCN, bit_source_gpu
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.ofdm import const_map, ofdm_tx
    from dl_ofdm_b200.radio import rayleigh_chan_lte
    B = 1037
    fl = Flags(nbits=nb, channel=chan)
    o = ofdm_tx(fl)
    m = DCCN(nbits=nb, n_data=o.frame_size, precision='exact')
    bits = bit_source_gpu(B * o.frame_size * nb, seed=5, device=m.device).view(B, o.frame_size, nb)
    snr = torch.full((B,), 12.0, device=m.device)
    ch_a = rayleigh_chan_lte(fl, o.Fs, engine=m, seed=3)
    ch_b = rayleigh_chan_lte(fl, o.Fs, engine=m, seed=3)
    tx = m.transmit(bits, o, const_map(nb))
    rx_sep = ch_a.run(tx, snr)
    rx_fused = ch_b.run_bits(bits, o, const_map(nb), snr)
    # (after AWGN: the batch power is an fp64 sum of per-warp partials added with atomics, in an order that differs between
    #  the two kernels and from run to run -- the scale can differ in its last bit, i.e. single fp32 ulps in rx)
    assert (rx_sep - rx_fused).abs().max().item() <= 2.5e-7 * rx_sep.abs().max().item()
    assert (rx_sep != rx_fused).float().mean().item() < 0.05
    alpha, coeff = ch_a._profile(ch_a.profiles[0], m.device)
    n_taps = 0 if coeff is None else coeff.numel()
    z = torch.randn((B, max(n_taps, 1), 2), dtype=torch.float64, device=m.device) * np.sqrt(0.5)
    zz = z if n_taps else None
    faded_sep = torch.empty_like(tx)
    m.fading(tx, faded_sep, alpha, coeff, 0.0, o.Fs, zz, 0, 0, 1, True)
    faded, tx2 = m.transmit_fade(bits, o, const_map(nb), alpha, coeff, zz, seed=0, want_tx=True)
    assert torch.equal(tx2, tx) and torch.equal(faded, faded_sep)
    if n_taps == 0:
        assert torch.equal(faded, tx)
    m.close()


def test_host_entry_returns_hard_bits(libdccn, trained_dev):
    """Both host-buffer entry points with a pinned destination for the hard decisions (copied back on the library's own
    D2H stream): decisions == the device entry point's, through both slots, pipelined."""
    from dl_ofdm_b200.engine import DCCN
    rng = np.random.default_rng(19)
    nb, B = 4, 517
    m = DCCN(nbits=nb, equalizer=True, precision='parity')
    m.load_weights(trained_dev)
    xs = [torch.as_tensor((rng.standard_normal((B, 7, 80, 2)) * 0.4).astype(np.float32)).pin_memory() for _ in range(3)]
    bits = rng.integers(0, 2, (B, 320, nb)).astype(np.uint8)
    bh = torch.as_tensor(bits).pin_memory()
    packed = torch.as_tensor(np.packbits(bits.reshape(-1), bitorder='little')).pin_memory()
    refs = []
    for x in xs:
        o = m.forward(x.cuda(), bh.cuda(), want_soft=False)
        refs.append((o['hard'].cpu(), o['conf'].cpu().numpy()))
    for packed_labels in (False, True):
        hards = [torch.zeros((B, 320, nb), dtype=torch.uint8).pin_memory() for _ in range(3)]
        begin = (lambda s, x, h: m.forward_host_begin_packed(s, x, packed, h)) if packed_labels else \
                (lambda s, x, h: m.forward_host_begin(s, x, bh, h))
        begin(0, xs[0], hards[0])
        for i in range(3):
            if i + 1 < 3:
                begin((i + 1) & 1, xs[i + 1], hards[i + 1])
            conf, _ = m.forward_host_end(i & 1)
            assert np.array_equal(conf, refs[i][1])
            assert torch.equal(hards[i], refs[i][0])
    m.close()


def test_monitor_fetches(libdccn, trained_dev):
    """The monitor tensors the reference's drivers fetch next to the BER (dev/py/ofdmreceiver_np.py:80,172-183) and
    equalizer_ofdm's snr_db (dev/py/model.py:464-475) against NumPy restatements."""
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.model import Session, equalizer_ofdm
    from dl_ofdm_b200.ofdm import ofdm_tx
    from oracle import dccn_oracle as orc
    from oracle.dccn_oracle_lean import LeanModel
    rng = np.random.default_rng(23)
    B = 300
    fl = Flags(nbits=4)
    o = ofdm_tx(fl)
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.7 + 0.1).astype(np.float32)
    x[5, 3, 10] = (90.0, -70.0)                                   # an outlier that the PAPR clip (norm 8) catches
    bits = rng.integers(0, 2, (B, 320, 4)).astype(np.uint8)
    snr = np.full((B, 1), 10.0, dtype=np.float32)
    s = Session(fl, o, trained_dev, precision='parity')
    conf, cost, txp, npw, inp, iq_tx, iq_rx, ce = s.run(
        ['conf_matrix', 'cost', 'tx_power', 'noise_power', 'input', 'iq_tx', 'iq_rx', 'ce_mean'],
        {'tx_ofdm': x, 'bits_in': bits, 'SNR': snr})
    z, _, _ = orc.batch_moment_norm(x, np.float32)
    assert np.abs(inp.cpu().numpy() - z).max() < 5e-7 * np.abs(z).max()            # fp32 op order as TF; the oracle takes its moments in fp32, the kernel in fp64
    nrm = np.sqrt((z.astype(np.float64) ** 2).sum(-1, keepdims=True))
    clipped = z * 8.0 / np.maximum(nrm, 8.0)                       # tf.clip_by_norm(x, 8, axes=[-1])
    assert nrm.max() > 8.0
    assert abs(float(txp) - (clipped ** 2).sum(-1).mean()) < 1e-5 * float(txp)
    assert np.abs(iq_tx.float().cpu().numpy() - clipped.reshape(-1, 2)).max() < 4e-3          # fp16 rounding
    assert abs(float(npw) - 0.5 * 10 ** (-1.0)) < 0.02 * 0.05                                  # E|noise|^2 = level^2
    d = iq_rx.float().cpu().numpy() - clipped.reshape(-1, 2)
    assert abs((d ** 2).sum(-1).mean() - float(npw)) < 0.03 * float(npw)
    ber = (conf[0, 1] + conf[1, 0]) / conf.sum()
    reg = sum(0.01 * float((trained_dev[k].astype(np.float64) ** 2).sum()) for k in
              ('demodulation/dense/kernel', 'demodulation/dense/bias', 'demodulation/dense_1/kernel', 'demodulation/dense_1/bias'))
    assert abs(float(cost) - (float(ce) + ber * 1e-4 * reg + np.log(ber))) < 1e-5
    s.close()
    # snr_db: oracle = the literal definition on the lean oracle's phase-equalised frequency-domain frame
    zc = torch.as_tensor(z).cuda()
    eq, snr_db, chest = equalizer_ofdm(zc, fl, o, weights=trained_dev)
    lm = LeanModel(trained_dev, 4)
    xx = orc.layer_norm(z.astype(np.float64)).reshape(B, 7, 160)
    f = ((xx @ lm.g1[0] + lm.g1[1]) @ lm.g2[0] + lm.g2[1]).reshape(B, 7, 64, 2)
    _, ch = lm.equalizer(z.astype(np.float64))
    eqf = (f[..., 0] + 1j * f[..., 1]) * (np.conj(ch) / np.abs(ch))
    p = np.abs(eqf[:, :, o.pilotCarriers]).reshape(B, -1) ** 2
    ref = np.log10(np.clip(p.mean(1) / p.var(1), 1e-3, 1e4))
    assert np.abs(snr_db.cpu().numpy().reshape(-1) - ref).max() < 2e-3
