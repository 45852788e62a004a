"""Drop-in for the reference ``models/stgcn/stgcn.py`` (``Model``, ``StgcnLayer``).

Same constructor kwargs, forward signatures and state_dict keys as the reference
(SURVEY.md §8b); the forward bodies are replaced by the sm_100a kernels behind the
C ABI (``stgcn_model_forward`` / ``stgcn_layer_forward``).  Inference only:
forward runs under ``torch.no_grad()`` and dropout must be inactive, as in the
reference's test/benchmark routines.
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


class Model(nn.Module):
    """ST-GCN classifier: norm_in -> fcn_in -> L x StgcnLayer -> mean(T,V) -> fcn_out.

    Input ``(N, in_feat, T, V)``; output ``(N, num_classes, 1)`` (reference
    stgcn.py:80-97).  Extra, B200-only knobs (not in the reference): ``math``
    selects the GEMM arithmetic ('fp32' | 'bf16x3' | 'bf16').
    """

    def __init__(self, **kwargs):
        super().__init__()
        conf = kwargs['st-gcn']
        self.graph = Graph(strategy=kwargs['strategy'], **kwargs['graph'])
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)

        kernel_size = (conf['kernel'], kwargs['graph']['num_node'])
        self.normalization = kwargs['normalization']
        self.math = kwargs.get('math', 'bf16x3')
        if self.normalization == 'LayerNorm':
            self.norm_in = LayerNorm([kwargs['in_feat'], 1, A.size(1)])
        else:
            self.norm_in = BatchNorm1d(kwargs['in_feat'] * A.size(1), track_running_stats=False)
        self.fcn_in = Conv2d(in_channels=conf['in_feat'], out_channels=conf['in_ch'][0], kernel_size=1)
        self.gcn_networks = nn.ModuleList([
            StgcnLayer(in_channels=conf['in_ch'][i], out_channels=conf['out_ch'][i],
                       kernel_size=kernel_size, partitions=A.size(0), num_joints=A.size(1),
                       stride=conf['stride'][i], residual=not not conf['residual'][i],
                       dropout=conf['dropout'][i], normalization=kwargs['normalization'])
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

    # ------------------------------------------------------------------ #
    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.A.data_ptr(), self.math)

    def _descriptor(self):
        """(ModelDesc, keep-alive list); rebuilt when parameters change or move."""
        fp = self._fingerprint()
        if self._desc is not None and self._desc[0] == fp:
            return self._desc[1]
        keep = []
        layers = (_lib.LayerDesc * len(self.gcn_networks))()
        for i, (gcn, imp) in enumerate(zip(self.gcn_networks, self.edge_importance)):
            a_eff = (self.A * imp).contiguous()           # stgcn.py:89
            keep.append(a_eff)
            gcn._fill_desc(layers[i], a_eff)
        m = _lib.ModelDesc()
        m.in_feat = self.fcn_in.in_channels
        m.num_joints = self.A.size(1)
        m.partitions = self.A.size(0)
        m.num_classes = self.num_classes
        m.num_layers = len(self.gcn_networks)
        m.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        m.math = _lib.MATH_NAMES[self.math]
        nin = self.norm_in if self.normalization == 'LayerNorm' else self.norm_in.norm
        m.norm_in_w, m.norm_in_b = nin.weight.data_ptr(), nin.bias.data_ptr()
        m.fcn_in_w, m.fcn_in_b = self.fcn_in.weight.data_ptr(), self.fcn_in.bias.data_ptr()
        m.fcn_out_w, m.fcn_out_b = self.fcn_out.weight.data_ptr(), self.fcn_out.bias.data_ptr()
        m.layers = ctypes.cast(layers, ctypes.POINTER(_lib.LayerDesc))
        keep.append(layers)
        # host-known sparsity of the partitioned adjacency (the edge importance can only remove entries):
        # bit 1 allows the kernels that hold one pre-scaled weight copy per edge (at most 6 V edges)
        if int((self.graph.A != 0).sum()) <= 6 * self.A.size(1):
            m.reserved |= 2
        if self.fcn_in.weight.is_cuda:
            keep.append(_lib.prepare_model(m, self.fcn_in.weight.device))
        self._desc = (fp, (m, keep))
        return self._desc[1]

    @torch.no_grad()
    def forward(self, x, return_features=False):
        n, c, t, v = x.shape
        x = x.contiguous()
        dev = _lib.require_cuda(x, self.A, self.fcn_in.weight)
        if self.training and any(g.dropout_p > 0 for g in self.gcn_networks):
            raise RuntimeError("B200 ST-GCN path is inference-only (dropout active)")
        if getattr(self, '_graph_on', False) and not return_features:
            g = getattr(self, '_graph', None)
            key = (n, c, t, v, dev, self._fingerprint())
            if g is None or g['key'] != key:
                g = self._capture(x, key)
            g['x'].copy_(x)
            g['graph'].replay()
            return g['logits'].clone().unsqueeze(-1)
        lib = _lib.load()
        m, _ = self._descriptor()
        ws = self._ws.get(lib.stgcn_model_workspace_bytes(ctypes.byref(m), n, t), dev)
        logits = torch.empty((n, self.num_classes), device=dev, dtype=torch.float32)
        feats = None
        if return_features:
            tf = t
            for g in self.gcn_networks:
                tf = (tf - 1) // g.stride + 1
            feats = torch.empty((n, self.gcn_networks[-1].out_channels, tf, v), device=dev,
                                dtype=torch.float32)
        _lib.check(lib.stgcn_model_forward(
            ctypes.byref(m), _lib.ptr(x), _lib.ptr(logits), _lib.ptr(feats), n, t,
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        out = logits.unsqueeze(-1)                         # (N, classes, 1) like stgcn.py:97
        return (out, feats) if return_features else out

    def prepare_benchmark(self, arch_conf):
        return arch_conf

    def enable_cuda_graph(self, on=True):
        """Replay ``forward`` from a captured CUDA graph, one capture per input shape (the static input /
        output buffers and the workspace stay owned by this module).  Short trials (BASELINE config 1:
        N=1, T=300) are launch-bound -- about thirty launches of a few microseconds each -- and a replay
        issues them with one driver call.  ``return_features=True`` calls stay eager."""
        self._graph_on = bool(on)
        self._graph = None
        return self

    def _capture(self, x, key):
        n, c, t, v, dev, _ = key
        lib = _lib.load()
        m, _keep = self._descriptor()                         # prepared operands are built outside the capture
        ws = torch.empty(lib.stgcn_model_workspace_bytes(ctypes.byref(m), n, t), dtype=torch.uint8, device=dev)
        xs = torch.zeros_like(x)
        logits = torch.zeros((n, self.num_classes), device=dev, dtype=torch.float32)

        def launch():
            _lib.check(lib.stgcn_model_forward(ctypes.byref(m), _lib.ptr(xs), _lib.ptr(logits), None, n, t,
                                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                         # warm-up launch (function attributes, tensor maps)
            launch()
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            launch()
        self._graph = dict(graph=graph, x=xs, logits=logits, ws=ws, key=key, keep=_keep)
        return self._graph

    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def forward_windows(self, captures, receptive_field):
        """Sliding-window inference of one trial, the reference's continual use of ST-GCN
        (``WindowSegment``, utils/segment_generator.py:109-154 + processor.py:374-384): frame i is
        classified from the ``receptive_field`` frames ending at i (zeros before the start).
        ``captures (1, in_feat, L, V)`` -> ``(1, num_classes, L)``.  The window batch is read in
        place through overlapping strides instead of being materialised."""
        n, c, length, v = captures.shape
        if n != 1:
            raise RuntimeError("forward_windows takes one trial at a time (like the reference's 'dir' datasets)")
        if self.normalization != 'LayerNorm':
            raise RuntimeError("forward_windows needs LayerNorm (windows must be independent trials)")
        w = int(receptive_field)
        padded = torch.nn.functional.pad(captures, (0, 0, w - 1, 0)).contiguous()   # pad_sequence: W-1 in front
        dev = _lib.require_cuda(padded, self.A, self.fcn_in.weight)
        lib = _lib.load()
        m, _ = self._descriptor()
        ws = self._ws.get(lib.stgcn_model_workspace_bytes(ctypes.byref(m), length, w), dev)
        logits = torch.empty((length, self.num_classes), device=dev, dtype=torch.float32)
        _lib.check(lib.stgcn_model_forward_windows(ctypes.byref(m), _lib.ptr(padded), _lib.ptr(logits), length, w,
                                                   length + w - 1, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return logits.t().unsqueeze(0)                     # mask_segment: (N', C', 1) -> (1, C', L)

    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def forward_tsplit(self, x_local, total_frames, exchange):
        """T-split forward (BASELINE config 4): ``x_local (N, in_feat, T_local, V)`` is this rank's
        contiguous chunk of a ``total_frames``-frame trial (chunking: ``tsplit.chunk_bounds``);
        ``exchange`` is a ``tsplit.DistExchange`` (or compatible) that swaps the per-layer halo
        frames with the ring neighbours and all-reduces the pooled sums.  Returns the full-trial
        logits ``(N, num_classes, 1)`` on every rank."""
        from ... import tsplit
        n, c, t, v = x_local.shape
        x_local = x_local.contiguous()
        dev = _lib.require_cuda(x_local, self.A, self.fcn_in.weight)
        if self.normalization != 'LayerNorm' or self.math == 'fp32':
            raise RuntimeError("T-split needs LayerNorm and math in {'bf16x3','bf16'}: batch statistics "
                               "would span ranks, and the halo lives in the tensor-core operand layout")
        # same check on every rank, before anything is launched (a rank-local failure would leave the
        # other ranks blocked in the halo exchange)
        tsplit.validate_chunks(total_frames, getattr(exchange, 'world', 1), [g.stride for g in self.gcn_networks])
        lib = _lib.load()
        m, _ = self._descriptor()
        need = lib.stgcn_model_halo_bytes(ctypes.byref(m), n)
        if exchange.capacity < need:
            raise RuntimeError("halo staging buffers too small: %d < %d bytes" % (exchange.capacity, need))
        hd, keep = exchange.descriptor()
        ws = self._ws.get(lib.stgcn_model_tsplit_workspace_bytes(ctypes.byref(m), n, t), dev)
        c_last = self.gcn_networks[-1].out_channels
        sums = torch.empty((n, c_last), device=dev, dtype=torch.float32)
        exchange.error = None
        rc = lib.stgcn_model_forward_tsplit(ctypes.byref(m), _lib.ptr(x_local), _lib.ptr(sums), n, t,
                                            ctypes.byref(hd), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        if rc != 0 and getattr(exchange, 'error', None) is not None:
            raise exchange.error
        _lib.check(rc)
        del keep
        exchange.all_reduce_sum(sums)
        t_final = tsplit.frames_after(total_frames, [g.stride for g in self.gcn_networks])
        pooled = (sums / float(t_final * v)).view(n, c_last, 1, 1)
        return self.fcn_out(pooled).squeeze(-1)          # (N, classes, 1), stgcn.py:95-97


class StgcnLayer(nn.Module):
    """One st_gcn block: ``relu(norm2(conv_Gx1(relu(norm1(gcn(x, A))))) + res(x))``
    (reference stgcn.py:125-193).  ``forward(x, A)`` takes ``(N, C_in, T, V)`` and
    ``A`` as ``(K, V, V)`` or ``(N, K, V, V)``.
    """

    def __init__(self, in_channels, out_channels, kernel_size, partitions, num_joints, stride=1,
                 dropout=0, residual=True, normalization='LayerNorm'):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        padding = ((kernel_size[0] - 1) // 2, 0)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.temporal_kernel, self.stride = kernel_size[0], stride
        self.partitions, self.num_joints = partitions, num_joints
        self.normalization = normalization
        self.dropout_p = dropout
        self.is_residual = residual
        self.is_residual_conv = residual and not ((in_channels == out_channels) and (stride == 1))

        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1], partitions)
        self.tcn = nn.Sequential(
            _norm(normalization, out_channels, num_joints),
            nn.ReLU(inplace=True),
            Conv2d(out_channels, out_channels, (kernel_size[0], 1), stride=(stride, 1), padding=padding),
            _norm(normalization, out_channels, num_joints),
            nn.Dropout(dropout, inplace=True))
        if self.is_residual_conv:
            self.residual = nn.Sequential(
                Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                _norm(normalization, out_channels, num_joints))
        else:
            self.residual = nn.Identity()
        self.relu = nn.ReLU(inplace=True)
        self._ws = _lib.Workspace()

    def _fill_desc(self, d, a_eff, per_sample=0):
        d.c_in, d.c_out = self.in_channels, self.out_channels
        d.kernel, d.stride = self.temporal_kernel, self.stride
        d.residual = (_lib.RES_NONE if not self.is_residual else
                      _lib.RES_CONV if self.is_residual_conv else _lib.RES_IDENTITY)
        d.norm = _lib.NORM_LAYERNORM if self.normalization == 'LayerNorm' else _lib.NORM_BATCHNORM
        d.rt = 0
        d.a_per_sample = per_sample
        d.gcn_w, d.gcn_b = self.gcn.conv.weight.data_ptr(), self.gcn.conv.bias.data_ptr()
        d.a_eff = a_eff.data_ptr()
        d.n1_w, d.n1_b = self.tcn[0].weight.data_ptr(), self.tcn[0].bias.data_ptr()
        d.tcn_w, d.tcn_b = self.tcn[2].weight.data_ptr(), self.tcn[2].bias.data_ptr()
        d.n2_w, d.n2_b = self.tcn[3].weight.data_ptr(), self.tcn[3].bias.data_ptr()
        if self.is_residual_conv:
            d.res_w, d.res_b = self.residual[0].weight.data_ptr(), self.residual[0].bias.data_ptr()
            d.nr_w, d.nr_b = self.residual[1].weight.data_ptr(), self.residual[1].bias.data_ptr()

    @torch.no_grad()
    def forward(self, x, A, math='bf16x3'):
        n, c, t, v = x.shape
        x = x.contiguous()
        A = A.contiguous()
        dev = _lib.require_cuda(x, A, self.gcn.conv.weight)
        if self.training and self.dropout_p > 0:
            raise RuntimeError("B200 ST-GCN path is inference-only (dropout active)")
        k = self.partitions
        per_sample = 1 if A.dim() == 4 else 0
        if c != self.in_channels or v != self.num_joints or tuple(A.shape[-3:]) != (k, v, v) \
                or (per_sample and A.shape[0] != n):
            raise RuntimeError("StgcnLayer: bad input/adjacency shape %s / %s"
                               % (tuple(x.shape), tuple(A.shape)))
        lib = _lib.load()
        d = _lib.LayerDesc()
        self._fill_desc(d, A, per_sample)
        ws = self._ws.get(lib.stgcn_layer_workspace_bytes(ctypes.byref(d), k, v, n, t), dev)
        t_out = (t - 1) // self.stride + 1
        y = torch.empty((n, self.out_channels, t_out, v), device=dev, dtype=torch.float32)
        _lib.check(lib.stgcn_layer_forward(
            ctypes.byref(d), k, v, _lib.MATH_NAMES[math], _lib.ptr(x), _lib.ptr(y), n, t,
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return y
