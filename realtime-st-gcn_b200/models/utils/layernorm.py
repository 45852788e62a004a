"""Drop-in for the reference ``models/utils/layernorm.py`` ``LayerNorm``.

Custom LayerNorm over every non-singleton dim of ``normalized_shape`` with
*unbiased* variance (reference layernorm.py:19,22-28).  Every use on the hot
path is ``LayerNorm([C, 1, V])`` on ``(N, C, T, V)`` input (stgcn.py:46,152,160,
171; rtstgcn.py:101,320,331), i.e. statistics over (C, V) per (n, t); that is the
shape the CUDA kernel implements (C ABI ``stgcn_layernorm_forward``).
"""
import torch
import torch.nn as nn

from ... import _lib


class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-05, elementwise_affine=True, bias=True,
                 device=None, dtype=None):
        super().__init__()
        normalized_shape = list(normalized_shape)
        if len(normalized_shape) != 3 or normalized_shape[1] != 1:
            raise NotImplementedError("B200 LayerNorm supports normalized_shape [C, 1, V] "
                                      "(got %r)" % (normalized_shape,))
        self.weight = nn.Parameter(torch.ones(normalized_shape, device=device, dtype=dtype))
        if bias:
            self.bias = nn.Parameter(torch.zeros(normalized_shape, device=device, dtype=dtype))
        else:
            self.register_buffer('bias', torch.zeros(normalized_shape, device=device, dtype=dtype),
                                 persistent=False)
        self.eps = eps
        self.dim = [1, 3]

    @torch.no_grad()
    def forward(self, x):
        n, c, t, v = x.shape
        if (c, v) != (self.weight.shape[0], self.weight.shape[2]):
            raise RuntimeError("LayerNorm expects (N, %d, T, %d), got %s"
                               % (self.weight.shape[0], self.weight.shape[2], tuple(x.shape)))
        x = x.contiguous()
        dev = _lib.require_cuda(x, self.weight, self.bias)
        y = torch.empty_like(x)
        _lib.check(_lib.load().stgcn_layernorm_forward(
            _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(y),
            n, c, t, v, float(self.eps), _lib.stream_ptr(dev)))
        return y
