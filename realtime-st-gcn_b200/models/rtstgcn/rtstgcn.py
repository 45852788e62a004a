"""Drop-in for the reference ``models/rtstgcn/rtstgcn.py`` (``Model``, ``OfflineLayer``,
``OnlineLayer``, ``AggregateStgcn``).

RT-ST-GCN is not ST-GCN with a FIFO: its temporal stage is a plain *sum* of the
last Gamma graph-convolved frames, with no learnable temporal kernel and no
temporal down-sampling.  The continual (online) path keeps, per
layer and per stream, a ring FIFO of ``F = stride*(kernel-1)+1`` frames and
``stride`` running accumulators and follows the reference recurrence exactly
(rtstgcn.py:611-625):  ``acc <- (acc + z_t) + (-fifo[fi])``, output ``acc``,
``fifo[fi] <- z_t``.

Differences from the reference that do not change results:
  * state lives in one device buffer for ``B`` concurrent streams (the reference
    hard-wires batch 1 and keeps CPU-only plain tensors, rtstgcn.py:576-579);
    ``Model.reset_streams`` clears FIFOs (README TODO "Clear FIFOs after each trial");
  * the whole per-frame model step is one C-ABI call (``rtstgcn_step``).
The INT8 FX-quantisation classes (Observed/QAggregateStgcn) are out of scope.
"""
import ctypes

import torch
import torch.nn as nn

from ... import _lib
from ..utils import BatchNorm1d, BatchNorm2d, Conv2d, Graph, LayerNorm


def _norm(normalization, channels, joints):
    if normalization == 'LayerNorm':
        return LayerNorm([channels, 1, joints])
    return BatchNorm2d(channels, track_running_stats=False)


class Model(nn.Module):
    """Input ``(N, C_in, L, V)`` -> logits ``(N, num_classes, L)`` (rtstgcn.py:137-157)."""

    def __init__(self, rank=None, **kwargs):
        super().__init__()
        self.conf = kwargs['rt-st-gcn']
        self.graph = Graph(strategy=kwargs['strategy'], **kwargs['graph'])
        A = torch.tensor(self.graph.A, dtype=torch.float32, device=rank, requires_grad=False)
        self.register_buffer('A', A)
        self.normalization = kwargs['normalization']
        self.math = kwargs.get('math', 'bf16x3')
        self.num_classes = kwargs['num_classes']
        if self.normalization == 'LayerNorm':
            self.norm_in = LayerNorm([kwargs['in_feat'], 1, A.size(1)])
        else:
            self.norm_in = BatchNorm1d(kwargs['in_feat'] * A.size(1), track_running_stats=False)
        self.fcn_in = Conv2d(in_channels=self.conf['in_feat'], out_channels=self.conf['in_ch'][0],
                             kernel_size=1)
        self.st_gcn = nn.ModuleList([
            OfflineLayer(**self._layer_kwargs(i, kwargs['graph']['num_node']))
            for i in range(self.conf['layers'])])
        self.avg_pool = nn.AvgPool2d(kernel_size=(1, kwargs['graph']['num_node']))
        self.fcn_out = Conv2d(in_channels=self.conf['out_ch'][-1], out_channels=kwargs['num_classes'],
                              kernel_size=1)
        self._ws = _lib.Workspace()
        self._desc = None
        self._state = None
        self._state_streams = 0
        self._graph_on = False
        self._graph = None
        # few streams (<= 14): run the whole step as ONE cluster kernel (csrc/kernels_rt_small.cuh)
        self.small_batch_kernel = kwargs.get('small_batch_kernel', True)

    def _layer_kwargs(self, i, num_joints):
        c = self.conf
        return dict(num_joints=num_joints, in_channels=c['in_ch'][i], out_channels=c['out_ch'][i],
                    kernel_size=c['kernel'], stride=c['stride'][i], num_partitions=self.A.shape[0],
                    residual=not not c['residual'][i], dropout=c['dropout'][i],
                    importance=c['importance'], graph=self.A, normalization=self.normalization)

    # ------------------------------------------------------------------ #
    def _swap_layers_for_inference(self):
        """Replace the trainable OfflineLayers by OnlineLayers carrying the same weights
        (rtstgcn.py:160-187)."""
        new = nn.ModuleList([OnlineLayer(**self._layer_kwargs(i, self.A.shape[-1]))
                             for i in range(self.conf['layers'])])
        new.to(self.A.device)
        new.load_state_dict(self.st_gcn.state_dict(), strict=False)
        self.st_gcn = new
        self._desc = None
        self._state = None

    def prepare_benchmark(self, arch_conf):
        """Swap to online layers and bake edge importance into each layer's adjacency.
        (The reference body is broken at HEAD -- rtstgcn.py:193 calls a method that does not
        exist -- this does what it was meant to; the INT8 dicts are out of scope.)"""
        self._swap_layers_for_inference()
        for module in self.st_gcn:
            module.eval_()
        return arch_conf

    @property
    def is_online(self):
        return len(self.st_gcn) > 0 and isinstance(self.st_gcn[0], OnlineLayer)

    # ------------------------------------------------------------------ #
    def _fingerprint(self):
        extra = tuple((l.aggregate.A.data_ptr(), l.aggregate.A._version) for l in self.st_gcn) \
            if self.is_online else ()
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + extra + (self.math, self.small_batch_kernel)

    def _descriptor(self):
        fp = self._fingerprint()
        if self._desc is not None and self._desc[0] == fp:
            return self._desc[1]
        keep = []
        layers = (_lib.LayerDesc * len(self.st_gcn))()
        for i, layer in enumerate(self.st_gcn):
            a_eff = layer.aggregate.A.contiguous()        # baked by eval_() (rtstgcn.py:522-525)
            keep.append(a_eff)
            layer._fill_desc(layers[i], a_eff)
        m = _lib.ModelDesc()
        m.in_feat = self.fcn_in.in_channels
        m.num_joints = self.A.size(1)
        m.partitions = self.A.size(0)
        m.num_classes = self.num_classes
        m.num_layers = len(self.st_gcn)
        m.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        m.math = _lib.MATH_NAMES[self.math]
        m.reserved = 0 if self.small_batch_kernel else 1
        if int((self.graph.A != 0).sum()) <= 6 * self.A.size(1):      # sparse adjacency: per-joint-weight GEMM allowed
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
        return self._desc[1]

    def _ensure_state(self, streams, dev):
        lib = _lib.load()
        m, _ = self._descriptor()
        if self._state is None or self._state_streams != streams or self._state.device != dev:
            nbytes = lib.rtstgcn_state_bytes(ctypes.byref(m), streams)
            self._state = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self._state_streams = streams
        return self._state

    def reset_streams(self, first=0, count=None):
        """Zero the FIFOs / accumulators / frame counters of streams [first, first+count)."""
        if self._state is None:
            return
        count = self._state_streams - first if count is None else count
        m, _ = self._descriptor()
        _lib.check(_lib.load().rtstgcn_state_reset(
            ctypes.byref(m), _lib.ptr(self._state), self._state_streams, first, count,
            _lib.stream_ptr(self._state.device)))

    def _launch_step(self, frame, logits, b, dev, top5=None):
        lib = _lib.load()
        m, _ = self._descriptor()
        state = self._ensure_state(b, dev)
        ws = self._ws.get(lib.rtstgcn_step_workspace_bytes(ctypes.byref(m), b), dev)
        if top5 is not None:
            _lib.check(lib.rtstgcn_step_top5(ctypes.byref(m), _lib.ptr(frame), _lib.ptr(state), _lib.ptr(logits),
                                             _lib.ptr(top5), b, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
            return
        _lib.check(lib.rtstgcn_step(ctypes.byref(m), _lib.ptr(frame), _lib.ptr(state), _lib.ptr(logits), b,
                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))

    @torch.no_grad()
    def step_top5(self, frame):
        """One continual step that also ranks the classes on the device: returns ``(logits (B, classes, 1),
        top5 (B, 5) int32)`` -- ``Statistics`` (utils/statistics.py:4-16) without a separate top-k pass."""
        b, c, l, v = frame.shape
        if l != 1:
            raise RuntimeError("step_top5() takes exactly one frame per stream")
        frame = frame.contiguous()
        dev = _lib.require_cuda(frame, self.A, self.fcn_in.weight)
        logits = torch.empty((b, self.num_classes), device=dev, dtype=torch.float32)
        top5 = torch.empty((b, 5), device=dev, dtype=torch.int32)
        self._launch_step(frame, logits, b, dev, top5)
        return logits.unsqueeze(-1), top5

    @torch.no_grad()
    def step(self, frame):
        """One continual step: ``frame (B, C_in, 1, V)`` -> logits ``(B, num_classes, 1)``.

        With ``cuda_graph=True`` (set by ``enable_cuda_graph``) the step's kernel sequence is
        captured once per stream count and replayed: the per-frame cost at batch 1 is launch
        latency, and a graph replay issues the whole sequence with one driver call."""
        b, c, l, v = frame.shape
        if l != 1:
            raise RuntimeError("step() takes exactly one frame per stream")
        frame = frame.contiguous()
        dev = _lib.require_cuda(frame, self.A, self.fcn_in.weight)
        if self._graph_on:
            g = self._graph
            if g is None or g['b'] != b or g['fp'] != self._fingerprint() or g['dev'] != dev:
                g = self._capture(frame, b, dev)
            g['x'].copy_(frame)
            g['graph'].replay()
            return g['logits'].clone().unsqueeze(-1)
        logits = torch.empty((b, self.num_classes), device=dev, dtype=torch.float32)
        self._launch_step(frame, logits, b, dev)
        return logits.unsqueeze(-1)

    def enable_cuda_graph(self, on=True):
        """Replay the continual step from a captured CUDA graph (state, workspace and the static
        input/output buffers stay owned by this module)."""
        self._graph_on = bool(on)
        self._graph = None
        return self

    def _capture(self, frame, b, dev):
        x = torch.zeros_like(frame)
        logits = torch.zeros((b, self.num_classes), device=dev, dtype=torch.float32)
        self._descriptor()                       # prepared operands are built outside the capture
        state = self._ensure_state(b, dev)
        saved = state.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up launches (function attributes, tensor maps)
            self._launch_step(x, logits, b, dev)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._launch_step(x, logits, b, dev)
        state.copy_(saved)                       # the warm-up step must not advance the streams
        self._graph = dict(graph=graph, x=x, logits=logits, b=b, dev=dev, fp=self._fingerprint())
        return self._graph

    @torch.no_grad()
    def forward(self, x):
        if not self.is_online:
            # training-time definition on a whole sequence (rtstgcn.py:137-157 with OfflineLayers):
            # (N, C, L, V) -> (N, classes, L); every stage is a C-ABI call
            h = self.fcn_in(self.norm_in(x.contiguous()))
            for layer in self.st_gcn:
                h = layer(h, self.A)
            n, c, l, v = h.shape
            pooled = torch.empty((n, c, l, 1), device=h.device, dtype=torch.float32)
            _lib.check(_lib.load().stgcn_mean_joints_forward(_lib.ptr(h), _lib.ptr(pooled), n * c * l, v,
                                                             _lib.stream_ptr(h.device)))
            return self.fcn_out(pooled).squeeze(-1)
        if x.shape[2] == 1:
            return self.step(x)
        # buffered realtime: feed the frames one by one (state carries across calls)
        return torch.cat([self.step(x[:, :, t:t + 1]) for t in range(x.shape[2])], dim=2)


class _LayerBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, num_joints, stride, num_partitions,
                 dropout, residual, importance, graph, normalization='LayerNorm'):
        super().__init__()
        assert kernel_size % 2 == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_partitions, self.num_joints = num_partitions, num_joints
        self.stride, self.kernel_size = stride, kernel_size
        self.normalization = normalization
        self.dropout_p = dropout
        self.is_residual = residual
        self.is_residual_conv = residual and not ((in_channels == out_channels) and (stride == 1))
        if importance:
            self.edge_importance = nn.Parameter(torch.ones(num_partitions, num_joints, num_joints),
                                                requires_grad=self._importance_grad)
        else:
            self.edge_importance = 1
        self.conv = Conv2d(in_channels, out_channels * num_partitions, kernel_size=1)
        self.bn_relu = nn.Sequential(_norm(normalization, out_channels, num_joints), nn.ReLU())
        if self.is_residual_conv:
            self.residual = nn.Sequential(
                Conv2d(in_channels, out_channels, kernel_size=1, bias=False),
                _norm(normalization, out_channels, num_joints))
        else:
            self.residual = nn.Identity()
        if not residual:
            self.do = nn.Dropout(dropout)
        else:
            self.do = nn.Sequential(nn.ReLU(), nn.Dropout(dropout))
        self._ws = _lib.Workspace()

    def _fill_desc(self, d, a_eff, rt=1):
        d.c_in, d.c_out = self.in_channels, self.out_channels
        d.kernel, d.stride = self.kernel_size, self.stride
        d.residual = (_lib.RES_NONE if not self.is_residual else
                      _lib.RES_CONV if self.is_residual_conv else _lib.RES_IDENTITY)
        d.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        d.rt = rt
        d.a_per_sample = 0
        d.gcn_w, d.gcn_b = self.conv.weight.data_ptr(), self.conv.bias.data_ptr()
        d.a_eff = a_eff.data_ptr()
        d.n1_w, d.n1_b = self.bn_relu[0].weight.data_ptr(), self.bn_relu[0].bias.data_ptr()
        if self.is_residual_conv:
            d.res_w, d.res_b = self.residual[0].weight.data_ptr(), None
            d.nr_w, d.nr_b = self.residual[1].weight.data_ptr(), self.residual[1].bias.data_ptr()


class AggregateStgcn(nn.Module):
    """Holder of the per-layer adjacency and FIFO geometry (rtstgcn.py:556-588); the
    aggregation itself runs inside the fused step kernel."""

    def __init__(self, graph, fifo_size, kernel_size, out_channels, stride):
        super().__init__()
        self.out_channels = out_channels
        self.num_joints = graph.shape[1]
        self.stride = stride
        self.fifo_size = fifo_size
        self.kernel_size = kernel_size
        self.register_buffer('A', graph.clone().detach(), persistent=False)


class OnlineLayer(_LayerBase):
    """[Inference only] one frame per call, per-stream FIFO state (rtstgcn.py:392-553)."""
    _importance_grad = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        graph = kwargs['graph'] if 'graph' in kwargs else args[9]
        fifo_size = self.stride * (self.kernel_size - 1) + 1
        self.aggregate = AggregateStgcn(graph, fifo_size, self.kernel_size, self.out_channels, self.stride)
        self._state = None
        self._counter = None

    def eval_(self):
        # bakes the learned edge importance into the layer's adjacency (rtstgcn.py:522-525)
        with torch.no_grad():
            self.aggregate.A *= self.edge_importance
        return

    def reset(self):
        self._state = None
        self._counter = None

    @torch.no_grad()
    def forward(self, x, A=None):
        """``x (B, C_in, 1, V)`` -> ``(B, C_out, 1, V)``.  Like the reference, the ``A``
        argument is ignored: the layer uses ``self.aggregate.A`` (rtstgcn.py:528-545)."""
        b, c, l, v = x.shape
        if l != 1:
            raise RuntimeError("OnlineLayer processes one frame per call")
        x = x.contiguous()
        dev = _lib.require_cuda(x, self.aggregate.A, self.conv.weight)
        lib = _lib.load()
        k = self.num_partitions
        d = _lib.LayerDesc()
        self._fill_desc(d, self.aggregate.A, rt=1)
        if self._state is None or self._counter.numel() != b or self._state.device != dev:
            self._state = torch.zeros(lib.rtstgcn_layer_state_bytes(ctypes.byref(d), v, b),
                                      dtype=torch.uint8, device=dev)
            self._counter = torch.zeros(b, dtype=torch.int32, device=dev)
        ws = self._ws.get(lib.rtstgcn_layer_workspace_bytes(ctypes.byref(d), k, v, b), dev)
        y = torch.empty((b, self.out_channels, 1, v), device=dev, dtype=torch.float32)
        _lib.check(lib.rtstgcn_layer_step(
            ctypes.byref(d), k, v, _lib.MATH_FP32, _lib.ptr(x), _lib.ptr(y), _lib.ptr(self._state),
            _lib.ptr(self._counter), b, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return y


class OfflineLayer(_LayerBase):
    """[Training-time definition] whole-sequence layer (rtstgcn.py:220-389): causal sum of
    ``kernel // stride`` taps spaced ``stride`` (the reference builds an L x L band matrix per
    call; here it is a windowed sum).  Holds the trainable parameters; ``forward`` is inference
    of that definition (no autograd) through ``rtstgcn_offline_layer_forward``."""
    _importance_grad = True

    @torch.no_grad()
    def forward(self, x, A):
        n, c, l, v = x.shape
        x = x.contiguous()
        a_eff = (A * self.edge_importance).contiguous()        # rtstgcn.py:364
        dev = _lib.require_cuda(x, a_eff, self.conv.weight)
        lib = _lib.load()
        k = self.num_partitions
        d = _lib.LayerDesc()
        self._fill_desc(d, a_eff, rt=1)
        ws = self._ws.get(lib.rtstgcn_offline_layer_workspace_bytes(ctypes.byref(d), k, v, n, l), dev)
        y = torch.empty((n, self.out_channels, l, v), device=dev, dtype=torch.float32)
        _lib.check(lib.rtstgcn_offline_layer_forward(ctypes.byref(d), k, v, _lib.ptr(x), _lib.ptr(y), n, l,
                                                     _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return y
