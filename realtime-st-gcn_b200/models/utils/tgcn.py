"""Drop-in for the reference ``models/utils/tgcn.py`` ``ConvTemporalGraphical``.

``forward(x, A)``: 1x1 conv C_in -> K*C_out (channel index k*C_out + c), then
``z[n,c,t,w] = sum_k sum_v y[n,k*C_out+c,t,v] * A[k,v,w]`` (einsum nkctv,kvw,
reference tgcn.py:70-79).  ``A`` may be ``(K,V,V)`` or per-sample ``(N,K,V,V)``
(AA-GCN passes the latter, reference models/aagcn/aagcn.py:148).
"""
import torch
import torch.nn as nn

from ... import _lib
from .conv import Conv2d


class ConvTemporalGraphical(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, partitions, t_kernel_size=1,
                 t_stride=1, t_padding=0, t_dilation=1, bias=True):
        super().__init__()
        if (t_kernel_size, t_stride, t_padding, t_dilation) != (1, 1, 0, 1):
            raise NotImplementedError("only the 1x1 feature transform used by the reference is built")
        self.out_channels = out_channels
        self.partitions = partitions
        self.kernel_size = kernel_size
        self.conv = Conv2d(in_channels, out_channels * partitions, kernel_size=(1, 1), bias=bias)
        self._ws = _lib.Workspace()

    @torch.no_grad()
    def forward(self, x, A):
        n, c, t, v = x.shape
        x = x.contiguous()
        A = A.contiguous()
        dev = _lib.require_cuda(x, A, self.conv.weight, self.conv.bias)
        per_sample = 1 if A.dim() == 4 else 0
        k = self.partitions
        if tuple(A.shape[-3:]) != (k, v, v) or (per_sample and A.shape[0] != n):
            raise RuntimeError("adjacency must be (K,V,V) or (N,K,V,V), got %s" % (tuple(A.shape),))
        lib = _lib.load()
        ws = self._ws.get(lib.stgcn_graphconv_workspace_bytes(n, c, self.out_channels, k, t, v), dev)
        y = torch.empty((n, self.out_channels, t, v), device=dev, dtype=torch.float32)
        _lib.check(lib.stgcn_graphconv_forward(
            _lib.ptr(x), _lib.ptr(self.conv.weight), _lib.ptr(self.conv.bias), _lib.ptr(A), per_sample,
            _lib.ptr(y), n, c, self.out_channels, k, t, v, _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(dev)))
        return y
