"""GPU parity tests of the transfer-learning step (BASELINE config 4) through the C ABI against
oracle/dccn_train_oracle.py (fp64 NumPy backward, pinned against torch autograd on the TF mirror).

Tolerances: a gradient tensor g must satisfy  max|g - g64| <= 2e-4 * max|g64| + 1e-9  (fp32-class
arithmetic over contractions of up to 28672 terms; measured ~1e-5); Adam is checked exactly (1e-6)
on the GPU's own gradients, and end to end over a few steps where |g| is not at the noise floor
(Adam's first steps are ~ lr * sign(g), so a variable whose gradient is numerically zero is compared
by the gradient test, not by the weight test).
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

GRAD_RTOL = 2e-4


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda().contiguous()


def _case(seed, B, nbits, use_cp=True):
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(seed)
    w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (B, 320, nbits)).astype(np.uint8)
    return w, x, bits


def _engine(w, nbits, precision, max_batch, use_cp=True):
    from dl_ofdm_b200.engine import DCCN
    m = DCCN(nbits=nbits, equalizer=True, precision=precision, use_cp=use_cp)
    m.load_weights(w)
    m.train_init(max_batch)
    return m


@pytest.mark.parametrize('precision,nbits,use_cp,B', [
    ('exact', 2, True, 96), ('parity', 2, True, 96), ('parity', 4, True, 200), ('parity', 1, False, 64),
    ('exact', 4, False, 33)])
def test_gradients_match_oracle(libdccn, precision, nbits, use_cp, B):
    from oracle import dccn_train_oracle as tro
    w, x, bits = _case(20 + nbits, B, nbits, use_cp)
    ce, _, g64, _ = tro.loss_and_grads(x, bits, w, nbits, use_cp=use_cp)
    m = _engine(w, nbits, precision, B, use_cp)
    out = m.train_step(_cuda(x), _cuda(bits), 1e-3, apply_update=False)
    torch.cuda.synchronize()
    ce_gpu = float(out['ce_sum'][0]) / out['n_bits']
    assert abs(ce_gpu - ce) < 5e-6, (ce_gpu, ce)
    worst = {}
    for name in tro.trainable_names():
        g = m.get_grad(name).astype(np.float64).reshape(g64[name].shape)
        scale = np.abs(g64[name]).max()
        err = np.abs(g - g64[name]).max()
        worst[name] = err / max(scale, 1e-30)
        assert err <= GRAD_RTOL * scale + 1e-9, (name, err, scale)
    # weights untouched by apply_update=False
    assert np.array_equal(m.get_weight('Equalizer/dense_3/kernel'), w['Equalizer/dense_3/kernel'].ravel())
    assert m.global_step == 0
    m.close()


def test_adam_update_exact_on_own_gradients(libdccn):
    """w' = Adam(w, g_gpu): isolates the optimiser kernel (TF-1.15 formulation) from gradient noise."""
    from oracle import dccn_train_oracle as tro
    w, x, bits = _case(31, 64, 2)
    m = _engine(w, 2, 'parity', 64)
    xg, bg = _cuda(x), _cuda(bits)
    opt = tro.Adam(tro.trainable_names(), w, dtype=np.float64)
    wr = {k: np.array(v, dtype=np.float64) for k, v in w.items()}
    for step in range(3):
        lr = tro.learning_rate(1e-3, step)
        m.train_step(xg, bg, lr, apply_update=True)
        grads = {n: m.get_grad(n).astype(np.float64).reshape(w[n].shape) for n in tro.trainable_names()}
        opt.step(wr, grads, lr)
        for n in tro.trainable_names():
            got = m.get_weight(n).reshape(w[n].shape)
            assert np.abs(got - wr[n]).max() <= 2e-6, (step, n, np.abs(got - wr[n]).max())
            wr[n] = got.astype(np.float64)       # follow the fp32 trajectory
    assert m.global_step == 3
    # frozen receiver
    assert np.array_equal(m.get_weight('demodulation/dense/kernel'), w['demodulation/dense/kernel'].ravel())
    m.close()


@pytest.mark.parametrize('precision', ['exact', 'parity'])
def test_training_steps_match_oracle(libdccn, precision):
    """Three reference training steps on three different minibatches vs the fp64 oracle trajectory."""
    from oracle import dccn_train_oracle as tro
    w, x, bits = _case(41, 3 * 80, 2)
    xs = [x[i * 80:(i + 1) * 80] for i in range(3)]
    bs = [bits[i * 80:(i + 1) * 80] for i in range(3)]
    w_ref, losses_ref = tro.train_steps(xs, bs, w, 2)
    _, _, g0, _ = tro.loss_and_grads(xs[0], bs[0], w, 2)
    m = _engine(w, 2, precision, 80)
    losses = []
    for i in range(3):
        out = m.train_step(_cuda(xs[i]), _cuda(bs[i]), tro.learning_rate(1e-3, i))
        losses.append(float(out['ce_sum'][0]) / out['n_bits'])
    assert np.allclose(losses, losses_ref, rtol=0, atol=5e-6), (losses, losses_ref)
    for n in tro.trainable_names():
        got = m.get_weight(n).reshape(w[n].shape).astype(np.float64)
        ref = np.asarray(w_ref[n], dtype=np.float64)
        solid = np.abs(g0[n]) > 1e-3 * np.abs(g0[n]).max()       # gradient well above the fp32 noise floor
        assert solid.any()
        err = np.abs(got - ref)[solid].max()
        assert err <= 3e-5, (n, err)
        # everywhere: a step is bounded by ~lr per update
        assert np.abs(got - np.asarray(w[n], dtype=np.float64)).max() <= 3.5e-3
    m.close()


def test_training_reduces_loss_config4(libdccn):
    """Config-4 sized property test: QPSK, B = 4096 frames per step, loss goes down, receiver stays frozen,
    and the inference path sees the updated equalizer."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(7)
    nbits, B = 2, 4096
    w = orc.glorot_weights(rng, nbits, equalizer=True, chest_bias=(0.6, -0.4))
    w['demodulation/dense_1/kernel'] *= 30      # a confident (trained-like) head: the loss has something to gain
    x = torch.randn((B, 7, 80, 2), device='cuda') * 0.3
    bits = torch.randint(0, 2, (B, 320, nbits), device='cuda', dtype=torch.uint8)
    m = DCCN(nbits=nbits, equalizer=True, precision='parity')
    m.load_weights(w)
    m.train_init(B)
    before = m.forward(x, bits)
    losses = []
    for i in range(8):
        out = m.train_step(x, bits, 1e-3)
        losses.append(float(out['ce_sum'][0]) / out['n_bits'])
    after = m.forward(x, bits)
    torch.cuda.synchronize()
    assert abs(losses[0] - float(before['ce_sum'][0]) / before['n_bits']) < 1e-6
    assert losses[-1] < losses[0] - 2e-4, losses
    assert float(after['ce_sum'][0]) < float(before['ce_sum'][0])
    assert abs(float(after['ce_sum'][0]) / after['n_bits'] - losses[-1]) < 5e-3
    assert np.array_equal(m.get_weight('fft_like/conv3d/bias'), w['fft_like/conv3d/bias'].ravel())
    assert m.global_step == 8
    m.close()


def test_train_errors(libdccn):
    from dl_ofdm_b200.engine import DCCN
    from dl_ofdm_b200._lib import DccnError
    from oracle import dccn_oracle as orc
    rng = np.random.default_rng(1)
    w = orc.glorot_weights(rng, 2, equalizer=False)
    m = DCCN(nbits=2, equalizer=False, precision='exact')
    m.load_weights(w)
    with pytest.raises(DccnError):
        m.train_init(16, mode='eq')            # no equalizer -> no Equalizer/* variables to train
    m.close()
    w = orc.glorot_weights(rng, 2, equalizer=True, chest_bias=(0.6, -0.4))
    m = DCCN(nbits=2, equalizer=True, precision='exact')
    m.load_weights(w)
    x = torch.zeros((8, 7, 80, 2), device='cuda')
    bits = torch.zeros((8, 320, 2), device='cuda', dtype=torch.uint8)
    with pytest.raises(DccnError):
        m.train_step(x, bits, 1e-3)            # train_init not called
    with pytest.raises(DccnError):
        m.train_init(4, mode='rx')             # receiver training is defined for the handle without equalizer
    m.train_init(4)
    with pytest.raises(DccnError):
        m.train_step(x, bits, 1e-3)            # batch larger than max_batch
    m.close()


def test_train_equalizer_driver(libdccn, tmp_path):
    """Host mirror of the epoch loop (dev/py/ofdmreceiver_np_mp.py:394-466): data generation on the GPU,
    minibatch steps, checkpoint in the reference's bundle format and file name, reload gives the same model."""
    from dl_ofdm_b200 import tfbundle
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.init import receiver_variables
    from dl_ofdm_b200.model import load_model_np
    from dl_ofdm_b200.ofdm import ofdm_tx
    from dl_ofdm_b200.ofdmreceiver_np_mp import TRAINABLE, train_equalizer
    FLAGS = Flags(nbits=2, channel='EPA', batch_size=7 * 256, msg_length=7 * 1024, opt=0, token='T',
                  save_dir=str(tmp_path) + '/', precision='parity', early_stop=400)
    ofdmobj = ofdm_tx(FLAGS)
    rx = receiver_variables(np.random.default_rng(3), 2)
    session, hist = train_equalizer(FLAGS, ofdmobj, rx, max_epoch_num=2, log=lambda *a: None)
    assert len(hist) == 2 and hist[-1]['global_step'] == 8 and session.engine.global_step == 8
    assert all(np.isfinite(h['train_loss']) and np.isfinite(h['test_loss']) and 0.0 <= h['test_ber'] <= 1.0 for h in hist)
    path = str(tmp_path) + '/T_Equalizer_EPA'
    ck = tfbundle.read_checkpoint(path)
    # the checkpoint is the best-train-loss epoch; if that is the last one it equals the live weights
    if hist[1]['train_loss'] < hist[0]['train_loss']:
        for n in TRAINABLE:
            assert np.array_equal(ck[n].ravel(), session.engine.get_weight(n)), n
    assert np.array_equal(ck['demodulation/dense/kernel'], rx['demodulation/dense/kernel'])
    s2 = load_model_np(path, FLAGS=FLAGS, ofdmobj=ofdmobj, precision='parity')
    x = torch.randn((64, 7, 80, 2), device='cuda') * 0.3
    if hist[1]['train_loss'] < hist[0]['train_loss']:
        a = session.engine.forward(x)['soft']
        b = s2.engine.forward(x)['soft']
        assert (a - b).abs().max().item() < 2e-6      # tf32 hi/lo form (training handle) vs fp16 hi/lo form (reloaded)
    s2.close()
    session.close()


# ---------------------------------------------------------------------------------------------
# training of the basic receiver itself (DCCN_TRAIN_RX; dev/py/ofdmreceiver_np.py:154-189)
# ---------------------------------------------------------------------------------------------
def _rx_case(seed, B, nbits, use_cp=True):
    """Seeded receiver case whose leaky-ReLU inputs all stay away from the kink: a pre-activation within fp32 rounding of 0
    (1.5 M units per case; the first draw of one case had |v| = 1.2e-8) takes slope 0.2 or 1 depending on the last bit of
    the forward arithmetic, which changes the GRADIENT by a finite amount in any fp32 implementation -- such a draw tests
    the coin, not the kernels, so the next seed is taken instead."""
    from oracle import dccn_oracle as orc
    for attempt in range(20):
        rng = np.random.default_rng(seed + 1000 * attempt)
        w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=False, bias_scale=0.05)
        x = (rng.standard_normal((B, 7, 80, 2)) * 0.3).astype(np.float32)
        bits = rng.integers(0, 2, (B, 320, nbits)).astype(np.uint8)
        z, _, _ = orc.batch_moment_norm(x, np.float64)
        _, inter = orc.ofdm_dense_rx(z, w, nbits, 16, use_cp=use_cp, return_intermediate=True)
        oiq = inter['out_iq']
        h = oiq @ w['demodulation/conv2d/kernel'].reshape(2, -1).astype(np.float64) + w['demodulation/conv2d/bias']
        lg = np.concatenate([np.maximum(0.2 * h, h), oiq], -1) @ w['demodulation/dense_1/kernel'].astype(np.float64) + \
            w['demodulation/dense_1/bias']
        if min(np.abs(h).min(), np.abs(lg).min()) > 4e-8:     # these values are O(0.1): fp32 forward error ~1e-8
            return w, x, bits
    raise AssertionError('no kink-free draw')


def _rx_engine(w, nbits, precision, max_batch, use_cp=True):
    from dl_ofdm_b200.engine import DCCN
    m = DCCN(nbits=nbits, equalizer=False, precision=precision, use_cp=use_cp)
    m.load_weights(w)
    m.train_init(max_batch)          # no equalizer -> mode 'rx', REG_COEFF 1e-4
    return m


@pytest.mark.parametrize('precision,nbits,use_cp,B', [
    ('exact', 2, True, 96), ('parity', 1, True, 64), ('parity', 4, True, 200), ('parity', 2, False, 72), ('exact', 3, False, 33)])
def test_rx_gradients_match_oracle(libdccn, precision, nbits, use_cp, B):
    from oracle import dccn_train_oracle as tro
    w, x, bits = _rx_case(60 + nbits, B, nbits, use_cp)
    ce, _, berl, g64, _ = tro.rx_loss_and_grads(x, bits, w, nbits, use_cp=use_cp)
    m = _rx_engine(w, nbits, precision, B, use_cp)
    out = m.train_step(_cuda(x), _cuda(bits), 1e-3, apply_update=False)
    torch.cuda.synchronize()
    assert abs(float(out['ce_sum'][0]) / out['n_bits'] - ce) < 5e-6
    conf = out['conf'].cpu().numpy()
    assert abs(float(conf[0, 1] + conf[1, 0]) / conf.sum() - berl) < 1e-12        # the berlin that scales the regulariser
    for name in tro.rx_trainable_names():
        g = m.get_grad(name).astype(np.float64).reshape(g64[name].shape)
        scale = np.abs(g64[name]).max()
        err = np.abs(g - g64[name]).max()
        assert err <= GRAD_RTOL * scale + 1e-9, (name, err, scale)
    assert np.array_equal(m.get_weight('demodulation/dense/kernel'), w['demodulation/dense/kernel'].ravel())
    m.close()


def test_rx_training_steps_match_oracle(libdccn):
    """Three training steps of the basic receiver on three minibatches: the loss trajectory follows the fp64 oracle,
    every update is TF's Adam applied to the GPU's own gradients (which test_rx_gradients_match_oracle pins; Adam's first
    steps are ~ lr * sign(g), so weights are not compared where a gradient sits at the fp32 noise floor), and the inference
    path afterwards runs on the updated variables (GEMM operands and head weights)."""
    from oracle import dccn_oracle as orc
    from oracle import dccn_train_oracle as tro
    w, x, bits = _rx_case(71, 3 * 80, 2)
    xs = [x[i * 80:(i + 1) * 80] for i in range(3)]
    bs = [bits[i * 80:(i + 1) * 80] for i in range(3)]
    w_ref, losses_ref, _ = tro.rx_train_steps(xs, bs, w, 2)
    names = tro.rx_trainable_names()
    m = _rx_engine(w, 2, 'parity', 80)
    opt = tro.Adam(names, w, dtype=np.float64)
    wr = {k: np.array(v, dtype=np.float64) for k, v in w.items()}
    losses = []
    for i in range(3):
        lr = tro.learning_rate(1e-3, i)
        out = m.train_step(_cuda(xs[i]), _cuda(bs[i]), lr)
        losses.append(float(out['ce_sum'][0]) / out['n_bits'])
        grads = {n: m.get_grad(n).astype(np.float64).reshape(w[n].shape) for n in names}
        opt.step(wr, grads, lr)
        for n in names:
            got = m.get_weight(n).reshape(w[n].shape)
            assert np.abs(got - wr[n]).max() <= 2e-6, (i, n, np.abs(got - wr[n]).max())
            wr[n] = got.astype(np.float64)
    assert np.allclose(losses, losses_ref, rtol=0, atol=5e-6), (losses, losses_ref)
    assert m.global_step == 3
    w_gpu = dict(w)
    for n in names:
        w_gpu[n] = m.get_weight(n).reshape(w[n].shape)
        ref = np.asarray(w_ref[n], dtype=np.float64)
        # the bulk of every variable follows the oracle trajectory closely
        assert np.median(np.abs(w_gpu[n] - ref)) <= 3e-5, n
    # dead taps of the (1,T) 'same' kernel are untouched
    k0, k1 = w['fft_like/conv3d/kernel'], w_gpu['fft_like/conv3d/kernel']
    dead = np.ones(80, dtype=bool)
    dead[39] = False
    assert np.array_equal(k0[0, dead], k1[0, dead]) and not np.array_equal(k0[0, 39], k1[0, 39])
    # forward on the trained variables == oracle on the same (GPU-trained) variables
    o = m.forward(_cuda(xs[0]), _cuda(bs[0]))
    soft_ref = orc.basic_receiver(xs[0], w_gpu, 2, 16)
    assert np.abs(o['soft'].cpu().numpy() - soft_ref).max() < 2e-5
    m.close()


def test_train_receiver_driver(libdccn, tmp_path):
    """Host mirror of the epoch loop of dev/py/ofdmreceiver_np.py:193-274 (BASELINE config 1's launcher phase,
    run_local_ofdm.py --awgn=True): BPSK over AWGN at 5 dB, the receiver learns from glorot-uniform variables -- the test
    BER drops well below the untrained 0.5 within a few epochs -- and the checkpoint written as save_dir/token reloads
    into the same model."""
    from dl_ofdm_b200 import tfbundle
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.model import load_model_np
    from dl_ofdm_b200.ofdm import ofdm_tx
    from dl_ofdm_b200.ofdmreceiver_np import RX_TRAINABLE, train_receiver
    FLAGS = Flags(nbits=1, channel='AWGN', SNR=5.0, batch_size=512, msg_length=7 * 2048, token='OFDM_T_1mod',
                  save_dir=str(tmp_path) + '/', precision='parity', early_stop=200)
    ofdmobj = ofdm_tx(FLAGS)
    session, hist = train_receiver(FLAGS, ofdmobj, max_epoch_num=12, log=lambda *a: None)
    assert len(hist) == 12 and all(np.isfinite(h['train_loss']) for h in hist)
    assert hist[0]['global_step'] == 2048 // (512 // 7)
    assert hist[-1]['train_loss'] < hist[0]['train_loss']
    assert hist[-1]['test_ber'] < 0.25, [h['test_ber'] for h in hist]
    path = str(tmp_path) + '/OFDM_T_1mod'
    ck = tfbundle.read_checkpoint(path)
    best = min(range(len(hist)), key=lambda i: hist[i]['train_loss'])
    if best == len(hist) - 1:
        for n in RX_TRAINABLE:
            assert np.array_equal(ck[n].ravel(), session.engine.get_weight(n)), n
        s2 = load_model_np(path, FLAGS=FLAGS, ofdmobj=ofdmobj, precision='parity')
        x = torch.randn((64, 7, 80, 2), device='cuda') * 0.3
        # (the training handle keeps the tf32 hi/lo GEMM form, the reloaded one runs the fp16 hi/lo form: the same
        #  fp32-class results, not the same bits)
        assert (session.engine.forward(x)['soft'] - s2.engine.forward(x)['soft']).abs().max().item() < 2e-6
        s2.close()
    session.close()


# ---------------------------------------------------------------------------------------------
# transfer learning of the ablation equalizers (--opt 1..5) in front of the frozen receiver
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('opt,precision,nbits,use_cp,B', [
    (1, 'parity', 2, True, 96), (2, 'parity', 2, True, 96), (3, 'parity', 1, False, 64), (4, 'parity', 4, True, 120),
    (5, 'parity', 2, False, 72), (3, 'exact', 2, True, 40), (5, 'exact', 2, True, 40)])
def test_variant_gradients_match_oracle(libdccn, opt, precision, nbits, use_cp, B):
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    from oracle import dccn_train_oracle as tro
    rng = np.random.default_rng(300 + opt)
    w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), eq_opt=opt)
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (B, 320, nbits)).astype(np.uint8)
    ce, _, g64, _ = tro.variant_loss_and_grads(x, bits, w, nbits, opt, use_cp=use_cp)
    m = DCCN(nbits=nbits, equalizer=True, precision=precision, use_cp=use_cp, eq_opt=opt)
    m.load_weights(w)
    m.train_init(B)
    out = m.train_step(_cuda(x), _cuda(bits), 1e-3, apply_update=False)
    torch.cuda.synchronize()
    assert abs(float(out['ce_sum'][0]) / out['n_bits'] - ce) < 5e-6
    for name in tro.variant_trainable_names(opt):
        g = m.get_grad(name).astype(np.float64).reshape(g64[name].shape)
        scale = np.abs(g64[name]).max()
        err = np.abs(g - g64[name]).max()
        assert err <= GRAD_RTOL * scale + 1e-9, (name, err, scale)
    m.close()


def test_variant_training_steps(libdccn):
    """--opt 5 (tf.ifft tail, tanh-fused channel estimate): three steps, Adam exact on the GPU's own gradients, loss
    trajectory follows the fp64 oracle, the receiver stays frozen, inference sees the update."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    from oracle import dccn_train_oracle as tro
    opt, nbits = 5, 2
    rng = np.random.default_rng(77)
    w = orc.glorot_weights(rng, nbits, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), eq_opt=opt)
    x = (rng.standard_normal((3 * 64, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (3 * 64, 320, nbits)).astype(np.uint8)
    xs = [x[i * 64:(i + 1) * 64] for i in range(3)]
    bs = [bits[i * 64:(i + 1) * 64] for i in range(3)]
    _, losses_ref = tro.variant_train_steps(xs, bs, w, nbits, opt)
    names = tro.variant_trainable_names(opt)
    m = DCCN(nbits=nbits, equalizer=True, precision='parity', eq_opt=opt)
    m.load_weights(w)
    m.train_init(64)
    optm = tro.Adam(names, w, dtype=np.float64)
    wr = {k: np.array(v, dtype=np.float64) for k, v in w.items()}
    losses = []
    for i in range(3):
        lr = tro.learning_rate(1e-3, i)
        out = m.train_step(_cuda(xs[i]), _cuda(bs[i]), lr)
        losses.append(float(out['ce_sum'][0]) / out['n_bits'])
        grads = {n: m.get_grad(n).astype(np.float64).reshape(w[n].shape) for n in names}
        optm.step(wr, grads, lr)
        for n in names:
            got = m.get_weight(n).reshape(w[n].shape)
            assert np.abs(got - wr[n]).max() <= 2e-6, (i, n)
            wr[n] = got.astype(np.float64)
    assert np.allclose(losses, losses_ref, rtol=0, atol=1e-5), (losses, losses_ref)
    assert np.array_equal(m.get_weight('demodulation/dense/kernel'), w['demodulation/dense/kernel'].ravel())
    w_gpu = dict(w)
    for n in names:
        w_gpu[n] = m.get_weight(n).reshape(w[n].shape)
    o = m.forward(_cuda(xs[0]), _cuda(bs[0]))
    soft_ref, _, chest_ref = orc.equalized_receiver(xs[0], w_gpu, nbits, 64, 16, opt=opt)
    good = np.abs(chest_ref).reshape(64, -1).min(axis=1) > 2e-2
    assert good.sum() > 16
    assert np.quantile(np.abs(o['soft'].cpu().numpy()[good] - soft_ref[good]), 0.999) < 2e-4
    m.close()


def test_train_equalizer_driver_ablation(libdccn, tmp_path):
    """train_equalizer with --opt=3 (equalizer_dnnE): checkpoint under the reference's name, reload picks the graph."""
    from dl_ofdm_b200 import tfbundle
    from dl_ofdm_b200.flags import Flags
    from dl_ofdm_b200.init import receiver_variables
    from dl_ofdm_b200.model import load_model_np
    from dl_ofdm_b200.ofdm import ofdm_tx
    from dl_ofdm_b200.ofdmreceiver_np_mp import train_equalizer
    FLAGS = Flags(nbits=2, channel='EPA', batch_size=7 * 256, msg_length=7 * 1024, opt=3, token='T',
                  save_dir=str(tmp_path) + '/', precision='parity', early_stop=400)
    ofdmobj = ofdm_tx(FLAGS)
    rx = receiver_variables(np.random.default_rng(3), 2)
    session, hist = train_equalizer(FLAGS, ofdmobj, rx, max_epoch_num=2, log=lambda *a: None)
    assert len(hist) == 2 and session.engine.global_step == 8 and session.engine.eq_opt == 3
    assert all(np.isfinite(h['train_loss']) for h in hist)
    ck = tfbundle.read_checkpoint(str(tmp_path) + '/T_Equalizer3_EPA')
    assert 'Equalizer/dense_8/kernel' in ck and 'Equalizer/conv3d/kernel' not in ck
    s2 = load_model_np(str(tmp_path) + '/T_Equalizer3_EPA', FLAGS=FLAGS, ofdmobj=ofdmobj, precision='parity')
    assert s2.engine.eq_opt == 3
    s2.close()
    session.close()


@pytest.mark.parametrize('precision,B', [('parity', 96), ('exact', 40)])
def test_separateIQ_gradients_match_oracle(libdccn, precision, B):
    """--opt 7 (equalizer_separateIQ): layers_conv2d_vector layers are other packers for the same operands, so the
    config-4 backward applies with two extra tanh's; gradients of all 20 variables vs the fp64 oracle."""
    from dl_ofdm_b200.engine import DCCN
    from oracle import dccn_oracle as orc
    from oracle import dccn_train_oracle as tro
    nbits = 2
    rng = np.random.default_rng(407)
    w = orc.glorot_weights(rng, nbits, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), eq_opt=7)
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (B, 320, nbits)).astype(np.uint8)
    ce, _, g64, _ = tro.loss_and_grads(x, bits, w, nbits, opt=7)
    m = DCCN(nbits=nbits, equalizer=True, precision=precision, eq_opt=7)
    m.load_weights(w)
    m.train_init(B)
    out = m.train_step(_cuda(x), _cuda(bits), 1e-3, apply_update=False)
    torch.cuda.synchronize()
    assert abs(float(out['ce_sum'][0]) / out['n_bits'] - ce) < 5e-6
    for name in tro.trainable_names():
        g = m.get_grad(name).astype(np.float64).reshape(g64[name].shape)
        scale = np.abs(g64[name]).max()
        err = np.abs(g - g64[name]).max()
        assert err <= GRAD_RTOL * scale + 1e-9, (name, err, scale)
    # one update, then the inference path runs on the re-packed operands (bias kinds 3 / 4)
    m.train_step(_cuda(x), _cuda(bits), 1e-3)
    w_gpu = dict(w)
    for n in tro.trainable_names():
        w_gpu[n] = m.get_weight(n).reshape(w[n].shape)
    o = m.forward(_cuda(x), _cuda(bits))
    soft_ref, _, chest_ref = orc.equalized_receiver(x, w_gpu, nbits, 64, 16, opt=7)
    good = np.abs(chest_ref).reshape(B, -1).min(axis=1) > 2e-2
    assert good.sum() > B // 4
    assert np.quantile(np.abs(o['soft'].cpu().numpy()[good] - soft_ref[good]), 0.999) < 2e-4
    m.close()
