"""Op-level mirror of the reference's complex-layer library (dev/py/complex.py).

``layers_conv2d_complex`` is the layer of the DCCN path; ``layers_conv2d_vector`` is the one of the
``--opt 7`` ablation (equalizer_separateIQ).  The 1-D, transpose and streams variants are unused by every
reachable graph.  TF creates the variables inside the call; here they are passed in (``kernel``
[kl,kw,1,C,2*filters] / [kl,kw,2,C,2*filters] for the vector layer, ``bias`` [2*filters]).
"""
from __future__ import annotations

import torch

from .engine import cconv2d


def layers_conv2d_vector(inputs, filters, kernal, strides=1, padding='valid', kernel=None, bias=None):
    """dev/py/complex.py:199-255: one real conv3d over (length, width, IQ), kernel depth 2 across IQ, no complex
    recombination; same shapes in and out as layers_conv2d_complex."""
    return layers_conv2d_complex(inputs, filters, kernal, strides, padding, kernel, bias, _vector=True)


def layers_conv2d_complex(inputs, filters, kernal, strides=1, padding='valid', kernel=None, bias=None, _vector=False):
    """[batch, length, width, channel, IQ(2)] real or [batch, length, width, channel] complex64
    CUDA tensor -> same rank, `filters` output channels (dev/py/complex.py:140-196)."""
    if strides not in (1, (1, 1)):
        raise NotImplementedError('strides != 1 is never used on the DCCN path')
    if isinstance(kernal, int):
        kernal = (kernal, kernal)
    elif not (isinstance(kernal, tuple) and len(kernal) == 2):
        raise NameError('Unacceptable Kernal Size')
    complex_flag = False
    if inputs.dim() == 4 and inputs.is_complex():
        inputs = torch.view_as_real(inputs.contiguous())
        complex_flag = True
    elif not (inputs.dim() == 5 and inputs.shape[-1] == 2):
        raise TypeError('Check input tensor dtypes or shape')
    if kernel is None or bias is None:
        raise ValueError('kernel / bias tensors are required (TF would create them here)')
    out = cconv2d(inputs.float().contiguous(), kernel, bias, filters, kernal, padding, vector=_vector)
    return torch.view_as_complex(out) if complex_flag else out
