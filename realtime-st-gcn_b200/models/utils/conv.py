"""``nn.Conv2d`` stand-in for the (Gamma x 1) / (1 x 1) convolutions on the hot path.

Subclasses ``nn.Conv2d`` so constructor arguments, default initialisation and
state_dict keys are the reference's (stgcn.py:49,74,154-159,166-170; tgcn.py:48-55;
rtstgcn.py:104,131,317,330), but ``forward`` runs the sm_100a implicit-GEMM kernel
(C ABI ``stgcn_conv_forward``) instead of cuDNN.
"""
import torch
import torch.nn as nn

from ... import _lib


class Conv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        kh, kw = self.kernel_size
        if kw != 1 or self.stride[1] != 1 or self.padding != ((kh - 1) // 2, 0) \
                or self.dilation != (1, 1) or self.groups != 1 or kh % 2 != 1:
            raise NotImplementedError(
                "B200 Conv2d covers (G,1) kernels with 'same' temporal padding, "
                "stride (s,1), no dilation/groups -- the shapes the reference uses")
        self._ws = _lib.Workspace()

    @torch.no_grad()
    def forward(self, x):
        n, c, t, v = x.shape
        x = x.contiguous()
        dev = _lib.require_cuda(x, self.weight, self.bias)
        lib = _lib.load()
        co, g, s = self.out_channels, self.kernel_size[0], self.stride[0]
        t_out = (t - 1) // s + 1
        ws = self._ws.get(lib.stgcn_conv_workspace_bytes(n, c, co, t, v, g, s), dev)
        y = torch.empty((n, co, t_out, v), device=dev, dtype=torch.float32)
        _lib.check(lib.stgcn_conv_forward(
            _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(y),
            n, c, co, t, v, g, s, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return y
