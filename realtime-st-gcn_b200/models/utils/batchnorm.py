"""Drop-ins for the reference batch-statistics normalisations.

``BatchNorm1d``: reference ``models/utils/batchnorm.py`` -- one feature per
(v, c), statistics over (N, T) of the *current call* (``track_running_stats=
False``: eval mode still uses batch statistics).
``BatchNorm2d``: the ``nn.BatchNorm2d(C, track_running_stats=False)`` the
reference places in every st_gcn block (stgcn.py:152,160,171) -- per channel over
(N, T, V).  Both keep the reference state_dict keys (``norm.weight`` /
``weight``).
"""
import torch
import torch.nn as nn

from ... import _lib


class _BatchStat(nn.Module):
    """Affine parameters of a batch-statistics BatchNorm (no running stats)."""

    def __init__(self, features, eps=1e-5, track_running_stats=False):
        super().__init__()
        if track_running_stats:
            raise NotImplementedError("the reference only builds track_running_stats=False norms")
        self.weight = nn.Parameter(torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.eps = eps
        self._ws = _lib.Workspace()

    def _run(self, x, mode):
        n, c, t, v = x.shape
        x = x.contiguous()
        dev = _lib.require_cuda(x, self.weight, self.bias)
        lib = _lib.load()
        ws = self._ws.get(lib.stgcn_batchnorm_workspace_bytes(c, v, mode), dev)
        y = torch.empty_like(x)
        _lib.check(lib.stgcn_batchnorm_forward(
            _lib.ptr(x), _lib.ptr(self.weight), _lib.ptr(self.bias), _lib.ptr(y), n, c, t, v,
            float(self.eps), mode, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return y


class BatchNorm2d(_BatchStat):
    def __init__(self, num_features, track_running_stats=False):
        super().__init__(num_features, track_running_stats=track_running_stats)

    @torch.no_grad()
    def forward(self, x):
        if x.shape[1] != self.weight.numel():
            raise RuntimeError("BatchNorm2d expects %d channels" % self.weight.numel())
        return self._run(x, 0)


class BatchNorm1d(nn.Module):
    def __init__(self, features, track_running_stats=False):
        super().__init__()
        self.norm = _BatchStat(features, track_running_stats=track_running_stats)

    @torch.no_grad()
    def forward(self, x):
        n, c, t, v = x.shape
        if c * v != self.norm.weight.numel():
            raise RuntimeError("BatchNorm1d expects V*C == %d" % self.norm.weight.numel())
        return self.norm._run(x, 1)
