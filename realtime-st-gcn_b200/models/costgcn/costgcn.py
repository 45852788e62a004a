"""Drop-in for the reference ``models/costgcn/costgcn.py`` (``Model``, ``StgcnLayer``): CoST-GCN, the
continual ST-GCN the reference's README compares RT-ST-GCN against.

Frame-by-frame model: every layer keeps a FIFO of its last ``F = stride*(kernel-1)+1`` graph-convolved
frames and applies the LEARNABLE ``kernel x 1`` temporal convolution (dilation = stride) to it, plus a
residual delayed by ``kernel // 2`` frames (costgcn.py:190-211).  Same constructor kwargs (the
``'st-gcn'`` config group plus a per-layer ``'dilation'`` list), forward signature and state_dict keys as
the reference; the forward body is ``costgcn_step`` behind the C ABI.  Differences that do not change
results: the FIFOs live in one device buffer for ``B`` concurrent streams (the reference hard-wires
batch 1 with CPU-only plain tensors, costgcn.py:153-154) and hold the normalised frames (LayerNorm is per
frame, so normalising a frame once when it enters is the same as re-normalising the whole FIFO every
step); ``reset_streams`` restarts streams.  LayerNorm only (a single frame has no batch statistics).
"""
import ctypes

import torch
import torch.nn as nn

from ... import _lib
from ..utils import BatchNorm1d, BatchNorm2d, Conv2d, ConvTemporalGraphical, Graph, LayerNorm


def _norm(normalization, channels, joints):
    if normalization == 'LayerNorm':
        return LayerNorm([channels, 1, joints])
    return BatchNorm2d(channels, track_running_stats=False)


class StgcnLayer(nn.Module):
    """Parameter holder of one CoST-GCN block (costgcn.py:125-188); the step runs in ``Model``."""

    def __init__(self, in_channels, out_channels, kernel_size, partitions, num_joints, stride=1, dilation=1,
                 dropout=0, residual=True, normalization='LayerNorm'):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.gamma, self.stride, self.dilation = kernel_size[0], stride, dilation
        self.fifo_size = stride * (self.gamma - 1) + 1
        self.partitions, self.num_joints = partitions, num_joints
        self.normalization = normalization
        self.dropout_p = dropout
        self.is_residual = residual
        self.is_residual_conv = residual and not ((in_channels == out_channels) and (stride == 1))
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1], partitions)
        self.tcn = nn.Sequential(
            _norm(normalization, out_channels, num_joints),
            nn.ReLU(inplace=True),
            # parameter holder only ('valid' padding, dilation = stride, costgcn.py:165-170); the convolution
            # over the FIFO runs inside costgcn_step
            nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), dilation=(stride, 1)),
            _norm(normalization, out_channels, num_joints),
            nn.Dropout(dropout, inplace=True))
        if self.is_residual_conv:
            self.residual = nn.Sequential(Conv2d(in_channels, out_channels, kernel_size=1),
                                          _norm(normalization, out_channels, num_joints))
        else:
            self.residual = nn.Identity()
        self.relu = nn.ReLU(inplace=True)

    def _fill_desc(self, d, a_eff):
        d.c_in, d.c_out = self.in_channels, self.out_channels
        d.kernel, d.stride = self.gamma, self.stride
        d.residual = (_lib.RES_NONE if not self.is_residual else
                      _lib.RES_CONV if self.is_residual_conv else _lib.RES_IDENTITY)
        d.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        d.rt = 2
        d.a_per_sample = 0
        d.gcn_w, d.gcn_b = self.gcn.conv.weight.data_ptr(), self.gcn.conv.bias.data_ptr()
        d.a_eff = a_eff.data_ptr()
        d.n1_w, d.n1_b = self.tcn[0].weight.data_ptr(), self.tcn[0].bias.data_ptr()
        d.tcn_w, d.tcn_b = self.tcn[2].weight.data_ptr(), self.tcn[2].bias.data_ptr()
        d.n2_w, d.n2_b = self.tcn[3].weight.data_ptr(), self.tcn[3].bias.data_ptr()
        if self.is_residual_conv:
            d.res_w, d.res_b = self.residual[0].weight.data_ptr(), self.residual[0].bias.data_ptr()
            d.nr_w, d.nr_b = self.residual[1].weight.data_ptr(), self.residual[1].bias.data_ptr()


class Model(nn.Module):
    """``forward(x)``: ``x (B, in_feat, 1, V)`` -> ``(B, num_classes, 1)`` (costgcn.py:81-99); longer
    inputs ``(B, in_feat, L, V)`` are fed frame by frame -> ``(B, num_classes, L)``."""

    def __init__(self, **kwargs):
        super().__init__()
        conf = kwargs['st-gcn']
        self.graph = Graph(strategy=kwargs['strategy'], **kwargs['graph'])
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        kernel_size = (conf['kernel'], kwargs['graph']['num_node'])
        dilation = conf.get('dilation', [1] * conf['layers'])
        self.dilation = dilation[-1]
        self.normalization = kwargs['normalization']
        self.math = kwargs.get('math', 'bf16x3')
        if self.normalization == 'LayerNorm':
            self.norm_in = LayerNorm([kwargs['in_feat'], 1, A.size(1)])
        else:
            self.norm_in = BatchNorm1d(kwargs['in_feat'] * A.size(1), track_running_stats=False)
        self.fcn_in = Conv2d(in_channels=conf['in_feat'], out_channels=conf['in_ch'][0], kernel_size=1)
        self.gcn_networks = nn.ModuleList([
            StgcnLayer(in_channels=conf['in_ch'][i], out_channels=conf['out_ch'][i], kernel_size=kernel_size,
                       partitions=A.size(0), num_joints=A.size(1), stride=conf['stride'][i], dilation=dilation[i],
                       residual=not not conf['residual'][i], dropout=conf['dropout'][i],
                       normalization=kwargs['normalization'])
            for i in range(conf['layers'])])
        if conf['importance']:
            self.edge_importance = nn.ParameterList(
                [nn.Parameter(torch.ones(self.A.size())) for _ in self.gcn_networks])
        else:
            self.edge_importance = [1] * len(self.gcn_networks)
        self.fcn_out = Conv2d(conf['out_ch'][-1], out_channels=kwargs['num_classes'], kernel_size=1)
        self.num_classes = kwargs['num_classes']
        self._ws = _lib.Workspace()
        self._desc = None
        self._state, self._streams, self._t = None, 0, 0

    def prepare_benchmark(self, arch_conf):
        return arch_conf

    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.A.data_ptr(), self.math)

    def _descriptor(self):
        fp = self._fingerprint()
        if self._desc is not None and self._desc[0] == fp:
            return self._desc[1]
        keep = []
        layers = (_lib.LayerDesc * len(self.gcn_networks))()
        for i, (gcn, imp) in enumerate(zip(self.gcn_networks, self.edge_importance)):
            a_eff = (self.A * imp).contiguous()           # costgcn.py:90
            keep.append(a_eff)
            gcn._fill_desc(layers[i], a_eff)
        m = _lib.ModelDesc()
        m.in_feat = self.fcn_in.in_channels
        m.num_joints, m.partitions = self.A.size(1), self.A.size(0)
        m.num_classes = self.num_classes
        m.num_layers = len(self.gcn_networks)
        m.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        m.math = _lib.MATH_NAMES[self.math]
        if int((self.graph.A != 0).sum()) <= 6 * self.A.size(1):
            m.reserved |= 2
        nin = self.norm_in if self.normalization == 'LayerNorm' else self.norm_in.norm
        m.norm_in_w, m.norm_in_b = nin.weight.data_ptr(), nin.bias.data_ptr()
        m.fcn_in_w, m.fcn_in_b = self.fcn_in.weight.data_ptr(), self.fcn_in.bias.data_ptr()
        m.fcn_out_w, m.fcn_out_b = self.fcn_out.weight.data_ptr(), self.fcn_out.bias.data_ptr()
        m.layers = ctypes.cast(layers, ctypes.POINTER(_lib.LayerDesc))
        keep.append(layers)
        if self.fcn_in.weight.is_cuda:
            keep.append(_lib.prepare_model(m, self.fcn_in.weight.device))
        self._desc = (fp, (m, keep))
        self._state = None
        return self._desc[1]

    def reset_streams(self, first=0, count=None):
        """Restart streams [first, first+count): their FIFOs return to the reference's initial state."""
        if self._state is None:
            return
        count = self._streams - first if count is None else count
        m, _ = self._descriptor()
        _lib.check(_lib.load().costgcn_state_reset(ctypes.byref(m), _lib.ptr(self._state), self._streams, first,
                                                   count, _lib.stream_ptr(self._state.device)))
        if first == 0 and count == self._streams:
            self._t = 0

    @torch.no_grad()
    def step(self, frame):
        b, c, l, v = frame.shape
        if l != 1:
            raise RuntimeError("step() takes exactly one frame per stream")
        if self.normalization != 'LayerNorm':
            raise RuntimeError("Expected more than 1 value per channel: continual inference needs LayerNorm")
        frame = frame.contiguous()
        dev = _lib.require_cuda(frame, self.A, self.fcn_in.weight)
        lib = _lib.load()
        m, _ = self._descriptor()
        if self._state is None or self._streams != b or self._state.device != dev:
            nbytes = lib.costgcn_state_bytes(ctypes.byref(m), b)
            if nbytes == 0:
                _lib.check(1)
            self._state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._streams, self._t = b, 0
            _lib.check(lib.costgcn_state_reset(ctypes.byref(m), _lib.ptr(self._state), b, 0, b, _lib.stream_ptr(dev)))
        ws = self._ws.get(lib.costgcn_step_workspace_bytes(ctypes.byref(m), b), dev)
        logits = torch.empty((b, self.num_classes), device=dev, dtype=torch.float32)
        _lib.check(lib.costgcn_step(ctypes.byref(m), _lib.ptr(frame), _lib.ptr(self._state), self._t, _lib.ptr(logits),
                                    b, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        self._t += 1
        return logits.unsqueeze(-1)

    @torch.no_grad()
    def forward(self, x):
        if x.shape[2] == 1:
            return self.step(x)
        return torch.cat([self.step(x[:, :, t:t + 1]) for t in range(x.shape[2])], dim=2)
