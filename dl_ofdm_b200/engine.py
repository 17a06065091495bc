"""Python host over the C ABI: one ``DCCN`` object per (device, model).

PyTorch is used for device storage and streams only; every computation on the
path is a hand-written CUDA kernel inside libdccn.so.  The class plays the role
of the reference's ``tf.Session`` + imported graph: feeds ``tx_ofdm`` /
``bits_in`` and fetches ``output`` / ``conf_matrix`` / ``linear_ber`` /
``ce_mean`` (reference: dev/py/ofdmreceiver_np.py:80, dev/py/ofdmreceiver_np_mp.py:89).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DccnError, dccn_cfg, dccn_train_cfg


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DCCN:
    """Handle on the B200 implementation of [equalizer_ofdm ->] ofdm_dense_rx.

    Parameters mirror the reference FLAGS / ofdm_tx attributes; ``precision`` is
    'exact' (fp32 CUDA cores), 'parity' (tcgen05 3xTF32, fp32-equivalent) or
    'fast' (tcgen05 single-pass TF32, reduced precision).
    """

    def __init__(self, nbits, nfft=64, cp_len=16, nsymbol=7, nfilter=64, n_data=320, pilot_size=16,
                 use_cp=True, head='dev', equalizer=False, precision='parity', chunk_frames=0,
                 device=None, eq_opt=0):
        if not torch.cuda.is_available():
            raise DccnError('dl_ofdm_b200 needs a CUDA device (no CPU fallback)')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.cfg = dccn_cfg(nfft=nfft, cp_len=cp_len, nsymbol=nsymbol, nfilter=nfilter, nbits=nbits,
                            use_cp=int(bool(use_cp)), n_data=n_data, pilot_size=pilot_size,
                            head=_lib.HEAD_V1 if head == 'v1' else _lib.HEAD_DEV,
                            equalizer=int(bool(equalizer)), precision=_lib.PRECISIONS[precision],
                            chunk_frames=chunk_frames,
                            eq_opt=(0 if int(eq_opt) in (9, 10) else int(eq_opt)) if equalizer else 0)   # --opt 9 / 10 build equalizer_ofdm
        self.eq_opt = int(self.cfg.eq_opt)
        self.nbits, self.S, self.K, self.T, self.D = nbits, nsymbol, nfft, nfft + cp_len, n_data
        self.equalizer = bool(equalizer)
        self.precision = precision
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self._scratch = {}

    # -- lifecycle -----------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None):
            self.lib.dccn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_ofdm(cls, FLAGS, ofdmobj, equalizer=False, precision='parity', head='dev', **kw):
        """Build from the reference's FLAGS + ofdm_tx object (same fields the TF builders read).  ``eq_opt`` (keyword)
        selects the equalizer graph: 0 equalizer_ofdm (default), 1 nocconv, 2 noresdl, 3 dnnE, 4 noresdl2, 5 noresdl4."""
        return cls(nbits=FLAGS.nbits, nfft=ofdmobj.K, cp_len=ofdmobj.CP, nsymbol=ofdmobj.nSymbol,
                   nfilter=FLAGS.nfilter, n_data=ofdmobj.frame_size, pilot_size=ofdmobj.pilot_size,
                   use_cp=FLAGS.cp, head=head, equalizer=equalizer, precision=precision, **kw)

    # -- weights ---------------------------------------------------------------------
    def load_weights(self, weights):
        """``weights``: TF variable name -> array in the reference layout (see dccn.h)."""
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            if a.ndim == 0:
                continue
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.dccn_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p),
                                                shape, a.ndim))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_commit_weights(self._h, _stream()))

    def workspace_bytes(self):
        return int(self.lib.dccn_workspace_bytes(self._h))

    # -- measurement hooks -----------------------------------------------------------------
    def profile(self, on=True):
        _lib.check(self.lib.dccn_profile_enable(self._h, int(on)))

    def profile_collect(self):
        """-> {slot name: (total ms, launches)} since the last collect (synchronises)."""
        ms = (C.c_double * 32)()
        cnt = (C.c_int64 * 32)()
        n = _lib.check(self.lib.dccn_profile_collect(self._h, ms, cnt, 32))
        return {self.lib.dccn_profile_slot_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}

    # -- the pass ----------------------------------------------------------------------
    def forward(self, x, bits=None, want_soft=True, want_hard=True, want_eq=False, want_chest=False,
                flags=0, snr_pilot_carriers=None):
        """x: float32 CUDA tensor [B,S,T,2]; bits: uint8 CUDA tensor [B,D,nbits] or None.

        Returns dict(soft, hard, eq, chest, conf (int64 [2,2] tensor), ce_sum (float64 [1]), n_bits).
        All outputs are CUDA tensors; nothing is synchronised.
        """
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (self.S, self.T, 2), x.shape
        dev = x.device
        eq_only = bool(flags & _lib.FWD_EQ_ONLY)
        out = {}
        soft = hard = eq = chest = conf = ce = None
        if want_soft and not eq_only:
            soft = torch.empty((B, self.D, self.nbits, 2), dtype=torch.float32, device=dev)
        if want_hard and not eq_only:
            hard = torch.empty((B, self.D, self.nbits), dtype=torch.uint8, device=dev)
        if want_eq:
            eq = torch.empty((B, self.S, self.T, 2), dtype=torch.float32, device=dev)
        if want_chest:
            chest = torch.empty((B, self.S, self.K, 2), dtype=torch.float32, device=dev)
        if bits is not None and not eq_only:
            assert bits.is_cuda and bits.dtype == torch.uint8 and bits.is_contiguous()
            assert tuple(bits.shape) == (B, self.D, self.nbits), bits.shape
            conf = torch.zeros((2, 2), dtype=torch.int64, device=dev)
            ce = torch.zeros((1,), dtype=torch.float64, device=dev)
        snr_db = None
        with torch.cuda.device(dev):
            if snr_pilot_carriers is not None:         # equalizer_ofdm's snr_db monitor (model.py:464-475), one-shot request
                pc = torch.as_tensor(np.asarray(snr_pilot_carriers, dtype=np.int32), device=dev)
                snr_db = torch.empty((B, 1), dtype=torch.float32, device=dev)
                self._scratch['snr_pc'] = pc           # keep alive until the pass has run
                _lib.check(self.lib.dccn_forward_monitors(self._h, _ptr(snr_db), _ptr(pc), pc.numel()))
            _lib.check(self.lib.dccn_forward(self._h, _ptr(x), B, _ptr(bits), _ptr(soft), _ptr(hard), _ptr(eq),
                                             _ptr(chest), _ptr(conf), _ptr(ce), int(flags), _stream()))
        out.update(soft=soft, hard=hard, eq=eq, chest=chest, conf=conf, ce_sum=ce, snr_db=snr_db,
                   n_bits=B * self.D * self.nbits)
        return out

    def forward_host(self, x_host, bits_host=None, want_hard=False):
        """End-to-end call with HOST (ideally pinned) tensors: H2D + pass + D2H, synchronous.

        Returns (conf int64[2,2] numpy, ce_sum float, hard uint8 tensor or None).
        """
        assert not x_host.is_cuda and x_host.dtype == torch.float32 and x_host.is_contiguous()
        B = x_host.shape[0]
        key = ('host', B, want_hard)
        if key not in self._scratch:
            self._scratch[key] = (torch.zeros(4, dtype=torch.int64).pin_memory(),
                                  torch.zeros(1, dtype=torch.float64).pin_memory(),
                                  torch.empty((B, self.D, self.nbits), dtype=torch.uint8).pin_memory()
                                  if want_hard else None)
        conf, ce, hard = self._scratch[key]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_forward_host(self._h, _ptr(x_host), B, _ptr(bits_host), _ptr(hard),
                                                  _ptr(conf), _ptr(ce), _stream()))
        return conf.numpy().reshape(2, 2).copy(), float(ce[0]), hard

    def forward_host_begin(self, slot, x_host, bits_host=None, hard_host=None):
        """Queue H2D + pass + D2H for one batch on `slot` (0/1) and return immediately.  ``hard_host``: optional pinned
        uint8 [B,D,nbits] destination of the hard decisions (valid after ``forward_host_end(slot)``)."""
        assert not x_host.is_cuda and x_host.dtype == torch.float32 and x_host.is_contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_forward_host_begin(self._h, int(slot), _ptr(x_host), x_host.shape[0],
                                                        _ptr(bits_host), _ptr(hard_host), _stream()))

    def forward_host_begin_packed(self, slot, x_host, bits_packed_host, hard_host=None):
        """forward_host_begin with labels packed 8 per byte (``np.packbits(bits.reshape(-1), bitorder='little')``)."""
        assert not x_host.is_cuda and x_host.dtype == torch.float32 and x_host.is_contiguous()
        assert not bits_packed_host.is_cuda and bits_packed_host.dtype == torch.uint8
        assert bits_packed_host.numel() * 8 == x_host.shape[0] * self.D * self.nbits
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_forward_host_begin_packed(self._h, int(slot), _ptr(x_host), x_host.shape[0],
                                                               _ptr(bits_packed_host), _ptr(hard_host), _stream()))

    def forward_host_end(self, slot):
        """Block until the batch on `slot` is done; -> (conf int64[2,2] numpy, ce_sum float)."""
        conf = (C.c_int64 * 4)()
        ce = C.c_double()
        _lib.check(self.lib.dccn_forward_host_end(self._h, int(slot), conf, C.byref(ce)))
        return np.array(conf[:], dtype=np.int64).reshape(2, 2), float(ce.value)

    def monitors(self, x, snr_db=None, seed=0, want_input=False, want_iq=False):
        """The reference graph's monitor tensors for a batch x [B,S,T,2] (dev/py/ofdmreceiver_np.py:125-149,172-183):
        dict(tx_power, noise_power (None without snr_db), input, iq_tx, iq_rx).  Synchronises (two scalars come back)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B = x.shape[0]
        n = B * self.S * self.T
        sums = torch.zeros(2, dtype=torch.float64, device=x.device)
        inp = torch.empty_like(x) if want_input else None
        iq_tx = torch.empty((n, 2), dtype=torch.float16, device=x.device) if want_iq else None
        iq_rx = torch.empty((n, 2), dtype=torch.float16, device=x.device) if (want_iq and snr_db is not None) else None
        if snr_db is not None:
            snr_db = torch.as_tensor(snr_db, dtype=torch.float32, device=x.device).reshape(-1).contiguous()
            assert snr_db.numel() == B
        with torch.cuda.device(x.device):
            _lib.check(self.lib.dccn_monitors(self._h, _ptr(x), B, _ptr(snr_db), int(seed), _ptr(sums), _ptr(inp), _ptr(iq_tx),
                                              _ptr(iq_rx), _stream()))
        s = sums.cpu().numpy()
        return dict(tx_power=np.float32(s[0] / n), noise_power=np.float32(s[1] / n) if snr_db is not None else None,
                    input=inp, iq_tx=iq_tx, iq_rx=iq_rx)

    def batch_moments(self, x):
        P = self.S * self.T * 2
        mean = torch.empty(P, dtype=torch.float32, device=x.device)
        rstd = torch.empty(P, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.dccn_batch_moments(self._h, _ptr(x), x.shape[0], _ptr(mean), _ptr(rstd), _stream()))
        return mean, rstd

    # -- training (BASELINE config 4) -------------------------------------------------------
    def get_weight(self, name):
        """Variable in the reference layout (the trained value once train_init was called)."""
        n = int(_lib.check(self.lib.dccn_get_weight(self._h, name.encode(), None, 0)))
        out = np.empty(n, dtype=np.float32)
        _lib.check(self.lib.dccn_get_weight(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n))
        return out

    def train_init(self, max_batch, reg_coeff=None, l2=0.01, beta1=0.9, beta2=0.999, eps=1e-8, mode=None):
        """Optimiser state for a training graph of the reference.

        mode 'eq' (default when the model has an equalizer): ``optimizer.minimize(total_loss, var_list=Equalizer vars)``
        in front of the frozen receiver, REG_COEFF 0.001 (dev/py/ofdmreceiver_np_mp.py:335-347).
        mode 'rx' (default otherwise): training of the basic receiver itself, every ofdm_dense_rx variable trainable,
        ``total_loss = ce_mean + berlin * 1e-4 * sum(reg) + ber`` (dev/py/ofdmreceiver_np.py:154-189)."""
        if mode is None:
            mode = 'eq' if self.equalizer else 'rx'
        assert mode in ('eq', 'rx'), mode
        if reg_coeff is None:
            reg_coeff = 0.001 if mode == 'eq' else 0.0001
        cfg = dccn_train_cfg(reg_coeff=reg_coeff, l2=l2, beta1=beta1, beta2=beta2, eps=eps, max_batch=int(max_batch),
                             mode=_lib.TRAIN_EQ if mode == 'eq' else _lib.TRAIN_RX)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dccn_train_init(self._h, C.byref(cfg), _stream()))

    def train_step(self, x, bits, learning_rate, apply_update=True, flags=0):
        """One minibatch of ``session.run([train_op, ce_mean, conf_matrix], {x, y})`` (_mp.py:419).

        Returns dict(conf int64 [2,2] CUDA tensor, ce_sum float64 [1] CUDA tensor, n_bits); nothing is synchronised.
        """
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert bits.is_cuda and bits.dtype == torch.uint8 and bits.is_contiguous()
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (self.S, self.T, 2) and tuple(bits.shape) == (B, self.D, self.nbits)
        conf = torch.zeros((2, 2), dtype=torch.int64, device=x.device)
        ce = torch.zeros((1,), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.dccn_train_step(self._h, _ptr(x), B, _ptr(bits), float(learning_rate),
                                                int(bool(apply_update)), _ptr(conf), _ptr(ce), int(flags), _stream()))
        return dict(conf=conf, ce_sum=ce, n_bits=B * self.D * self.nbits)

    def get_grad(self, name):
        """d total_loss / d var of the last train_step, reference layout (flat float32)."""
        n = int(_lib.check(self.lib.dccn_train_get_grad(self._h, name.encode(), None, 0)))
        out = np.empty(n, dtype=np.float32)
        _lib.check(self.lib.dccn_train_get_grad(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n))
        return out

    @property
    def global_step(self):
        return int(self.lib.dccn_train_global_step(self._h))

    @global_step.setter
    def global_step(self, step):
        _lib.check(self.lib.dccn_train_set_global_step(self._h, int(step)))

    # -- channel ------------------------------------------------------------------------
    def channel(self, tx, snr_db, alpha=None, coeff=None, z=None, normals=None, seed=0, want_faded=False):
        """rayleigh static FIR (optional) + AWGN on the GPU.

        tx float32 [B,S,T,2]; snr_db float32 [B]; alpha float64 [n_taps,n_fir]; coeff float64 [n_taps];
        z float64 [B,n_taps,2] (optional injected path draws); normals float64 [B,S,T,2] (optional).
        """
        assert tx.is_cuda and tx.dtype == torch.float32 and tx.is_contiguous()
        B = tx.shape[0]
        n_samp = tx.numel() // (B * 2)
        n_taps = 0 if coeff is None else int(coeff.numel())
        n_fir = 1 if alpha is None else int(alpha.shape[1])
        rx = torch.empty_like(tx)
        faded = torch.empty_like(tx) if want_faded else None
        with torch.cuda.device(tx.device):
            _lib.check(self.lib.dccn_chan_fir_awgn(self._h, _ptr(tx), B, n_samp, _ptr(alpha), _ptr(coeff), n_taps,
                                                   n_fir, _ptr(z), _ptr(snr_db), _ptr(normals), int(seed),
                                                   _ptr(rx), _ptr(faded), _stream()))
        return (rx, faded) if want_faded else rx

    def fading(self, tx, faded, alpha=None, coeff=None, doppler_hz=0.0, sample_rate=0.96e6, draws=None, seed=0,
               frame0=0, fstride=1, reset_power=True):
        """Fading only, into `faded`, for frames frame0, frame0+fstride, ... (Doppler if doppler_hz > 0)."""
        B, S, T = tx.shape[0], tx.shape[1], tx.shape[2]
        n_taps = 0 if coeff is None else int(coeff.numel())
        n_fir = 1 if alpha is None else int(alpha.shape[1])
        with torch.cuda.device(tx.device):
            _lib.check(self.lib.dccn_chan_fading(self._h, _ptr(tx), B, S, T, _ptr(alpha), _ptr(coeff), n_taps, n_fir,
                                                 float(doppler_hz), float(sample_rate), _ptr(draws), int(seed),
                                                 int(frame0), int(fstride), int(bool(reset_power)), _ptr(faded),
                                                 _stream()))

    def awgn(self, faded, snr_db, normals=None, seed=0):
        rx = torch.empty_like(faded)
        B = faded.shape[0]
        with torch.cuda.device(faded.device):
            _lib.check(self.lib.dccn_chan_awgn(self._h, _ptr(faded), B, faded.numel() // (2 * B), _ptr(snr_db),
                                               _ptr(normals), int(seed), _ptr(rx), _stream()))
        return rx

    # -- transmitter ---------------------------------------------------------------------
    def _tx_tables(self, ofdmobj, constellation, dev):
        key = ('tx', id(ofdmobj))
        if key not in self._scratch:
            for sc in (ofdmobj.dataSc, ofdmobj.pilotSc):        # the device-side map builder ignores bad indices
                sc = np.asarray(sc)
                if sc.size and (sc.min() < 0 or sc.max() >= self.S * self.K):
                    raise DccnError('subcarrier index out of range [0, %d)' % (self.S * self.K))
            # (the ofdm_tx object is kept in the entry: id() of a collected object may be handed out again)
            self._scratch[key] = (torch.as_tensor(np.asarray(ofdmobj.dataSc, dtype=np.int32), device=dev),
                                  torch.as_tensor(np.asarray(ofdmobj.pilotSc, dtype=np.int32), device=dev),
                                  torch.as_tensor(np.stack([constellation.real, constellation.imag], -1)
                                                  .astype(np.float32), device=dev), ofdmobj)
        return self._scratch[key][:3]

    def transmit(self, bits, ofdmobj, constellation):
        """GPU OFDM transmitter: bits uint8 [B,D,nbits] -> float32 [B,S,T,2]."""
        dev = bits.device
        dsc, psc, const = self._tx_tables(ofdmobj, constellation, dev)
        B = bits.shape[0]
        tx = torch.empty((B, self.S, self.T, 2), dtype=torch.float32, device=dev)
        pv = complex(ofdmobj.pilotValue)
        with torch.cuda.device(dev):
            _lib.check(self.lib.dccn_tx_frames(self._h, _ptr(bits), B, _ptr(dsc), dsc.numel(), _ptr(psc),
                                               psc.numel(), _ptr(const), pv.real, pv.imag, _ptr(tx), _stream()))
        return tx

    def can_tx_fade(self):
        return self.K == 64 and 0 < (self.T - self.K) <= 16 and self.S <= 8

    def transmit_fade(self, bits, ofdmobj, constellation, alpha=None, coeff=None, z=None, seed=0, want_tx=False,
                      reset_power=True):
        """Fused feeder (nfft = 64): bits uint8 [B,D,nbits] -> faded float32 [B,S,T,2] (static Rayleigh FIR; alpha / coeff
        None = no fading, i.e. the 'AWGN' channel) without the transmitted frames ever reaching HBM.  Bit-identical to
        ``fading(transmit(bits))``.  Returns (faded, tx or None); the batch power is left for ``awgn``."""
        dev = bits.device
        dsc, psc, const = self._tx_tables(ofdmobj, constellation, dev)
        B = bits.shape[0]
        faded = torch.empty((B, self.S, self.T, 2), dtype=torch.float32, device=dev)
        tx = torch.empty_like(faded) if want_tx else None
        n_taps = 0 if coeff is None else int(coeff.numel())
        n_fir = 1 if alpha is None else int(alpha.shape[1])
        pv = complex(ofdmobj.pilotValue)
        with torch.cuda.device(dev):
            _lib.check(self.lib.dccn_tx_fade(self._h, _ptr(bits), B, _ptr(dsc), dsc.numel(), _ptr(psc), psc.numel(),
                                             _ptr(const), pv.real, pv.imag, _ptr(alpha), _ptr(coeff), n_taps, n_fir,
                                             _ptr(z), int(seed), int(bool(reset_power)), _ptr(tx), _ptr(faded), _stream()))
        return faded, tx


def launch_count():
    """Number of kernels libdccn has launched in this process."""
    return int(_lib.load().dccn_launch_count())


def bit_source_gpu(n, seed, device='cuda'):
    """util.bit_source on the GPU (Philox): uint8 tensor of n uniform bits."""
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.dccn_bit_source(_ptr(out), n, int(seed), _stream()))
    return out


def cconv2d(x, kernel, bias, filters, kernal, padding='valid', vector=False):
    """Op-level layers_conv2d_complex (dev/py/complex.py:140) or, with ``vector``, layers_conv2d_vector (:199) on CUDA
    tensors."""
    lib = _lib.load()
    assert x.is_cuda and x.dim() == 5 and x.shape[-1] == 2
    B, L, W, Cc, _ = x.shape
    kl, kw = (kernal, kernal) if isinstance(kernal, int) else kernal
    pad = 1 if padding.lower() == 'same' else 0
    Lo, Wo = (L, W) if pad else (L - kl + 1, W - kw + 1)
    y = torch.empty((B, Lo, Wo, filters, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        fn = lib.dccn_vconv2d if vector else lib.dccn_cconv2d
        _lib.check(fn(_ptr(x.contiguous()), B, L, W, Cc, _ptr(kernel.contiguous()),
                      _ptr(bias.contiguous()), filters, kl, kw, pad, _ptr(y), _stream()))
    return y


def ber_accum(hard, bits, conf=None):
    lib = _lib.load()
    if conf is None:
        conf = torch.zeros((2, 2), dtype=torch.int64, device=hard.device)
    with torch.cuda.device(hard.device):
        _lib.check(lib.dccn_ber_accum(_ptr(hard), _ptr(bits), hard.numel(), _ptr(conf), _stream()))
    return conf
