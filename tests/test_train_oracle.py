"""Pins the hand-derived backward pass of oracle/dccn_train_oracle.py (BASELINE config 4) against
torch autograd through the independent op-for-op mirror of the TF graph (oracle/tf_mirror.py)."""
import numpy as np
import pytest
import torch

from oracle import dccn_oracle as orc
from oracle import dccn_train_oracle as tro
from oracle.tf_mirror import TFMirror


def _case(seed, B, nbits, use_cp=True):
    rng = np.random.default_rng(seed)
    w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4))
    x = (rng.standard_normal((B, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (B, 320, nbits)).astype(np.uint8)
    return w, x, bits


def _autograd(w, x, bits, nbits, use_cp=True):
    m = TFMirror(w, nbits, use_cp=use_cp, equalizer=True)
    m.w = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=k.startswith('Equalizer/'))
           for k, v in w.items()}
    z = m.norm(torch.tensor(x, dtype=torch.float64))
    soft = m.dense_rx(m.equalizer(z)).reshape(-1, 2)
    y = torch.tensor(bits.reshape(-1).astype(np.int64))
    ce = torch.nn.functional.cross_entropy(soft, y)          # softmax-xent ON the softmax outputs, mean
    reg = sum(tro.L2_L * (m.w['Equalizer/' + n + s] ** 2).sum() for n in tro.DENSE_NAMES for s in ('/kernel', '/bias'))
    total = ce + tro.REG_COEFF * reg
    total.backward()
    return float(ce.detach()), float(reg.detach()), {k: v.grad.numpy() for k, v in m.w.items() if v.requires_grad}


@pytest.mark.parametrize('nbits,use_cp', [(2, True), (4, True), (1, False)])
def test_backward_matches_autograd(nbits, use_cp):
    w, x, bits = _case(3 + nbits, 24, nbits, use_cp)
    ce, reg, g, _ = tro.loss_and_grads(x, bits, w, nbits, use_cp=use_cp)
    ce_t, reg_t, g_t = _autograd(w, x, bits, nbits, use_cp)
    assert abs(ce - ce_t) < 1e-12
    assert abs(reg - reg_t) < 1e-9 * max(1.0, reg_t)
    assert set(g) == set(g_t) == set(tro.trainable_names())
    for k in g:
        ref = g_t[k]
        err = np.abs(g[k] - ref).max()
        assert err <= 1e-9 * max(np.abs(ref).max(), 1e-12) + 1e-15, (k, err, np.abs(ref).max())


def test_forward_matches_restatement():
    w, x, bits = _case(11, 16, 2)
    _, _, _, aux = tro.loss_and_grads(x, bits, w, 2)
    soft, eq, _ = orc.equalized_receiver(x, w, 2, 64, 16)
    assert np.abs(aux['soft'] - soft).max() < 1e-12
    assert np.abs(aux['oeq'] - eq).max() < 1e-10


def test_adam_and_schedule():
    # closed form of the first TF Adam step: w -= lr * g / (|g| + eps*sqrt(1-b2)) (to first order: lr*sign(g))
    wts = {'a': np.array([1.0, -2.0, 3.0])}
    opt = tro.Adam(['a'], wts)
    g = {'a': np.array([0.5, -1e-3, 0.0])}
    opt.step(wts, g, 1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = np.array([1.0, -2.0, 3.0]) - lr_t * (0.1 * g['a']) / (np.sqrt(0.001 * g['a'] ** 2) + 1e-8)
    assert np.allclose(wts['a'], exp, rtol=0, atol=1e-15)
    assert wts['a'][2] == 3.0
    assert tro.learning_rate(1e-3, 499) == 1e-3
    assert abs(tro.learning_rate(1e-3, 500) - 0.98e-3) < 1e-18
    assert abs(tro.learning_rate(1e-3, 1700) - 1e-3 * 0.98 ** 3) < 1e-18


def test_training_reduces_loss():
    w, x, bits = _case(5, 32, 2)
    w2, losses = tro.train_steps([x] * 6, [bits] * 6, w, 2)
    assert losses[-1] < losses[0]
    assert any(np.abs(w2[k] - w[k]).max() > 0 for k in tro.trainable_names())
    assert np.array_equal(w2['demodulation/dense/kernel'], w['demodulation/dense/kernel'])   # receiver frozen


# ---- training of the basic receiver (dev/py/ofdmreceiver_np.py:154-198) ---------------------------------------
def _rx_autograd(w, x, bits, nbits, use_cp=True):
    m = TFMirror(w, nbits, use_cp=use_cp, equalizer=False)
    m.w = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in w.items()}
    z = m.norm(torch.tensor(x, dtype=torch.float64))
    soft = m.dense_rx(z).reshape(-1, 2)
    y = torch.tensor(bits.reshape(-1).astype(np.int64))
    ce = torch.nn.functional.cross_entropy(soft, y)
    berlin = float((soft.argmax(1) != y).double().mean())             # constant for the gradient (no path through argmax)
    reg = sum(tro.L2_L * (m.w[n + s] ** 2).sum() for n in tro.RX_DENSE for s in ('/kernel', '/bias'))
    total = ce + berlin * tro.REG_COEFF_RX * reg + berlin
    total.backward()
    return float(ce.detach()), float(reg.detach()), berlin, {k: v.grad.numpy() for k, v in m.w.items() if v.grad is not None}


@pytest.mark.parametrize('nbits,use_cp', [(1, True), (4, True), (2, False)])
def test_rx_backward_matches_autograd(nbits, use_cp):
    rng = np.random.default_rng(40 + nbits)
    w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=False, bias_scale=0.05)
    x = (rng.standard_normal((24, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (24, 320, nbits)).astype(np.uint8)
    ce, reg, berl, g, _ = tro.rx_loss_and_grads(x, bits, w, nbits, use_cp=use_cp)
    ce_t, reg_t, berl_t, g_t = _rx_autograd(w, x, bits, nbits, use_cp)
    assert abs(ce - ce_t) < 1e-12 and abs(berl - berl_t) < 1e-15
    assert abs(reg - reg_t) < 1e-9 * max(1.0, reg_t)
    assert set(g) == set(tro.rx_trainable_names()) and set(g_t) >= set(g)
    for k in g:
        ref = g_t[k]
        err = np.abs(g[k] - ref).max()
        assert err <= 1e-9 * max(np.abs(ref).max(), 1e-12) + 1e-15, (k, err, np.abs(ref).max())
    # dead taps of the (1,T) 'same' kernel never move
    T = 80 if use_cp else 64
    dead = np.ones(T, dtype=bool)
    dead[(T - 1) // 2] = False
    assert np.all(g_t['fft_like/conv3d/kernel'][0, dead] == 0)


def test_rx_training_reduces_loss():
    rng = np.random.default_rng(8)
    w = orc.glorot_weights(rng, 2, equalizer=False, bias_scale=0.05)
    x = (rng.standard_normal((32, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (32, 320, 2)).astype(np.uint8)
    w2, losses, bers = tro.rx_train_steps([x] * 6, [bits] * 6, w, 2)
    assert losses[-1] < losses[0]
    assert all(np.abs(w2[k] - w[k]).max() > 0 for k in tro.rx_trainable_names())


# ---- transfer learning of the ablation equalizers (--opt 1..5) ---------------------------------------------------------
def _variant_forward_torch(m, z, opt):
    """equalizer_<variant> op by op in torch (complex arithmetic, torch.fft.ifft, padded conv3d for the complex convs):
    independent of the oracle's packed-GEMM backward."""
    from oracle.tf_mirror import conv2d_complex_tf
    sp = orc.EQ_SPECS[opt]
    names = dict(orc.eq_layer_names(opt))
    W = lambda role, s: m.w['Equalizer/' + names[role] + '/' + s]
    B, S, T, _ = z.shape
    K = 64
    mu = z.reshape(B, -1).mean(1).reshape(B, 1, 1, 1)
    var = ((z - mu) ** 2).reshape(B, -1).mean(1).reshape(B, 1, 1, 1)
    c = (z - mu) / torch.sqrt(var + orc.LN_EPS)
    c = c.reshape(B, S, T * 2) if m.use_cp else c[:, :, m.CP:m.CP + K, :].reshape(B, S, K * 2)
    c = c @ W('front1', 'kernel') + W('front1', 'bias')
    if sp['front2'] == 'cconv':
        f = conv2d_complex_tf(c.reshape(B, S, K, 1, 2), W('front2', 'kernel'), W('front2', 'bias'), 'valid')
        f = f.permute(0, 1, 3, 2, 4)[:, :, :, 0, :]
    else:
        f = (c @ W('front2', 'kernel') + W('front2', 'bias')).reshape(B, S, K, 2)
    inputs_c = torch.complex(f[..., 0], f[..., 1])
    c = f.reshape(B, S * K * 2) @ W('pilot', 'kernel') + W('pilot', 'bias')
    for i, act in enumerate(sp['chain']):
        c = c @ W('chain%d' % i, 'kernel') + W('chain%d' % i, 'bias')
        c = torch.tanh(c) if act else c
    c5 = c.reshape(B, S, K, 1, 2)
    if sp['toeplitz']:
        c5 = conv2d_complex_tf(c5, W('toeplitz', 'kernel'), W('toeplitz', 'bias'), 'same')
    chest = torch.complex(c5[..., 0], c5[..., 1])[:, :, :, 0]
    ab = torch.abs(chest)
    eq = inputs_c * torch.complex(chest.real / ab, -chest.imag / ab)
    if sp['tail'] == 'dense2':
        t = torch.stack([eq.real, eq.imag], -1).reshape(B, S, K * 2) @ W('tail1', 'kernel') + W('tail1', 'bias')
    else:
        e = torch.fft.ifft(eq, dim=-1)
        t = torch.stack([e.real, e.imag], -1).reshape(B, S, K * 2)
    return (t @ W('tail2', 'kernel') + W('tail2', 'bias')).reshape(B, S, T, 2)


@pytest.mark.parametrize('opt,nbits,use_cp', [(1, 2, True), (2, 2, True), (3, 1, False), (4, 4, True), (5, 2, False)])
def test_variant_backward_matches_autograd(opt, nbits, use_cp):
    rng = np.random.default_rng(90 + opt)
    w = orc.glorot_weights(rng, nbits, use_cp=use_cp, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), eq_opt=opt)
    x = (rng.standard_normal((20, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (20, 320, nbits)).astype(np.uint8)
    ce, reg, g, aux = tro.variant_loss_and_grads(x, bits, w, nbits, opt, use_cp=use_cp)
    m = TFMirror(w, nbits, use_cp=use_cp, equalizer=False)
    m.w = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=k.startswith('Equalizer/')) for k, v in w.items()}
    z = m.norm(torch.tensor(x, dtype=torch.float64))
    oeq = _variant_forward_torch(m, z, opt)
    soft = m.dense_rx(oeq).reshape(-1, 2)
    ce_t = torch.nn.functional.cross_entropy(soft, torch.tensor(bits.reshape(-1).astype(np.int64)))
    reg_t = sum(tro.L2_L * (m.w['Equalizer/' + n + s] ** 2).sum() for _, n in orc.eq_layer_names(opt) if n.startswith('dense')
                for s in ('/kernel', '/bias'))
    (ce_t + tro.REG_COEFF * reg_t).backward()
    assert abs(ce - float(ce_t.detach())) < 1e-12
    assert np.abs(aux['oeq'] - oeq.detach().numpy()).max() < 1e-10
    assert set(g) == set(tro.variant_trainable_names(opt))
    for k in g:
        ref = m.w[k].grad.numpy()
        err = np.abs(g[k] - ref).max()
        assert err <= 1e-9 * max(np.abs(ref).max(), 1e-12) + 1e-15, (k, err, np.abs(ref).max())


# ---- --opt 7: equalizer_separateIQ (layers_conv2d_vector) -----------------------------------------------------------
def _vconv_torch(x5, kernel, bias, padding):
    """layers_conv2d_vector as the literal conv3d over (length, width, IQ) with TF's SAME padding (complex.py:199-255)."""
    import torch.nn.functional as Fn
    xt = x5[:, :, :, 0, :].unsqueeze(1)                                  # [B,1,L,W,2]
    kl, kw = kernel.shape[0], kernel.shape[1]
    wt = kernel[:, :, :, 0, :].permute(3, 0, 1, 2).unsqueeze(1)           # [2F,1,kl,kw,2]
    if padding == 'same':
        xt = Fn.pad(xt, (0, 1, (kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kl - 1) // 2, kl - 1 - (kl - 1) // 2))
    y = Fn.conv3d(xt, wt, bias)[..., 0]                                   # IQ position 0: [B,2F,L',W']
    F_ = kernel.shape[4] // 2
    return torch.stack([y[:, :F_], y[:, F_:]], -1).permute(0, 2, 3, 1, 4)  # [B,L',W',F,2]


def _separateIQ_forward_torch(m, z):
    W = lambda n: m.w['Equalizer/' + n]
    B, S, T, _ = z.shape
    K = 64
    mu = z.reshape(B, -1).mean(1).reshape(B, 1, 1, 1)
    var = ((z - mu) ** 2).reshape(B, -1).mean(1).reshape(B, 1, 1, 1)
    c = ((z - mu) / torch.sqrt(var + orc.LN_EPS)).reshape(B, S, T * 2)
    c = c @ W('dense/kernel') + W('dense/bias')
    f = _vconv_torch(c.reshape(B, S, K, 1, 2), W('conv3d/kernel'), W('conv3d/bias'), 'valid')     # [B,S,1,K,2]
    f = f.permute(0, 1, 3, 2, 4)[:, :, :, 0, :]
    inputs_c = torch.complex(f[..., 0], f[..., 1])
    c = f.reshape(B, S * K * 2) @ W('dense_1/kernel') + W('dense_1/bias')
    for n in ('dense_2', 'dense_3', 'dense_4'):
        c = torch.tanh(c @ W(n + '/kernel') + W(n + '/bias'))
    c5 = _vconv_torch(c.reshape(B, S, K, 1, 2), W('conv3d_1/kernel'), W('conv3d_1/bias'), 'same')
    chest = torch.complex(c5[..., 0], c5[..., 1])[:, :, :, 0]
    ab = torch.abs(chest)
    eq = inputs_c * torch.complex(chest.real / ab, -chest.imag / ab)
    corr = eq * torch.conj(eq)

    def vc(v, n):
        o = _vconv_torch(torch.stack([v.real, v.imag], -1).reshape(B, S, K, 1, 2), W(n + '/kernel'), W(n + '/bias'), 'valid')
        return o.permute(0, 1, 3, 2, 4)[:, :, :, 0, :]
    cat = torch.cat([vc(eq, 'conv3d_3'), vc(corr, 'conv3d_2')], -1).reshape(B, S, K * 4)
    return (cat @ W('dense_5/kernel') + W('dense_5/bias')).reshape(B, S, T, 2)


def test_separateIQ_backward_matches_autograd():
    nbits = 2
    rng = np.random.default_rng(97)
    w = orc.glorot_weights(rng, nbits, equalizer=True, bias_scale=0.05, chest_bias=(0.6, -0.4), eq_opt=7)
    x = (rng.standard_normal((16, 7, 80, 2)) * 0.3).astype(np.float32)
    bits = rng.integers(0, 2, (16, 320, nbits)).astype(np.uint8)
    ce, reg, g, aux = tro.loss_and_grads(x, bits, w, nbits, opt=7)
    # forward of the GEMM-form oracle == the op-by-op restatement (conv2d_vector pinned against a literal conv3d)
    _, eq_ref, _ = orc.equalized_receiver(x, w, nbits, 64, 16, opt=7)
    assert np.abs(aux['oeq'] - eq_ref).max() < 1e-10
    m = TFMirror(w, nbits, equalizer=False)
    m.w = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=k.startswith('Equalizer/')) for k, v in w.items()}
    oeq = _separateIQ_forward_torch(m, m.norm(torch.tensor(x, dtype=torch.float64)))
    assert np.abs(aux['oeq'] - oeq.detach().numpy()).max() < 1e-10
    soft = m.dense_rx(oeq).reshape(-1, 2)
    ce_t = torch.nn.functional.cross_entropy(soft, torch.tensor(bits.reshape(-1).astype(np.int64)))
    reg_t = sum(tro.L2_L * (m.w['Equalizer/' + n + s] ** 2).sum() for n in tro.DENSE_NAMES for s in ('/kernel', '/bias'))
    (ce_t + tro.REG_COEFF * reg_t).backward()
    assert abs(ce - float(ce_t.detach())) < 1e-12
    assert set(g) == set(tro.trainable_names())
    for k in g:
        ref = m.w[k].grad.numpy()
        err = np.abs(g[k] - ref).max()
        assert err <= 1e-9 * max(np.abs(ref).max(), 1e-12) + 1e-15, (k, err, np.abs(ref).max())
