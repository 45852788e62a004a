"""Oracle restatement of the ST-GCN / RT-ST-GCN forward pass (CPU, torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional restatement: every function takes plain tensors / a flat
``state_dict``-style mapping keyed exactly like the reference checkpoints
 and returns tensors in the reference layout ``(N, C, T, V)``.
All floating-point work goes through the same ATen CPU ops the reference uses
(``conv2d``, ``matmul``, ``mean``/``var``), so on the same host it reproduces
the reference to rounding; ``dtype=torch.float64`` gives a ground truth for
error budgeting.

Reference lines followed (paths relative to the reference root):
  models/utils/layernorm.py:22-28     -> layer_norm_cv
  models/utils/batchnorm.py:13-23     -> batch_norm_input
  nn.BatchNorm2d(track_running_stats=False) as used in
  models/stgcn/stgcn.py:152,160,171   -> batch_norm_channels
  models/utils/tgcn.py:58-79          -> graph_conv
  models/stgcn/stgcn.py:125-193       -> stgcn_layer
  models/stgcn/stgcn.py:80-97         -> stgcn_model
  models/rtstgcn/rtstgcn.py:528-553,
  models/rtstgcn/rtstgcn.py:591-627   -> rt_layer_step
  models/rtstgcn/rtstgcn.py:137-157   -> rt_model_step
  models/rtstgcn/rtstgcn.py:343-389   -> rt_offline_layer (with the one-token
                                         `self.toeplitz`->`toeplitz` fix)
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


# --------------------------------------------------------------------------- #
# normalisations
# --------------------------------------------------------------------------- #
def layer_norm_cv(x, weight, bias, eps=EPS):
    """Custom LayerNorm over (C, V) per (n, t) with *unbiased* variance.

    reference models/utils/layernorm.py:22-28; ``weight``/``bias`` are
    ``(C, 1, V)``.
    """
    mean = x.mean(dim=(1, 3), keepdim=True)
    var = x.var(dim=(1, 3), keepdim=True)          # unbiased (divisor C*V-1)
    y = (x - mean) / torch.sqrt(var + eps)
    return weight * y + bias


def batch_norm_channels(x, weight, bias, eps=EPS):
    """Batch-statistics BN per channel over (N, T, V), biased variance.

    ``nn.BatchNorm2d(track_running_stats=False)`` uses batch statistics in
    train *and* eval.
    """
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), keepdim=True, unbiased=False)
    y = (x - mean) / torch.sqrt(var + eps)
    return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def batch_norm_input(x, weight, bias, eps=EPS):
    """Input BN: one feature per (v, c), statistics over (N, T).

    reference models/utils/batchnorm.py:13-23; ``weight``/``bias`` are
    ``(V*C,)`` with feature index ``v*C + c``.
    """
    n, c, t, v = x.shape
    mean = x.mean(dim=(0, 2), keepdim=True)
    var = x.var(dim=(0, 2), keepdim=True, unbiased=False)
    y = (x - mean) / torch.sqrt(var + eps)
    w = weight.view(v, c).t().reshape(1, c, 1, v)
    b = bias.view(v, c).t().reshape(1, c, 1, v)
    return y * w + b


def _norm(kind, x, weight, bias):
    if kind == 'LayerNorm':
        return layer_norm_cv(x, weight, bias)
    return batch_norm_channels(x, weight, bias)


# --------------------------------------------------------------------------- #
# graph convolution (tgcn.py:58-79)
# --------------------------------------------------------------------------- #
def graph_conv(x, conv_w, conv_b, A):
    """1x1 conv C_in -> K*C_out, then contraction with ``A`` and sum over K.

    ``A`` is ``(K, V, V)`` or ``(N, K, V, V)``; output ``(N, C_out, T, V)``.
    """
    n, _, t, v = x.shape
    k = A.shape[-3]
    y = F.conv2d(x, conv_w, conv_b)
    c_out = y.shape[1] // k
    y = y.view(n, k, c_out * t, v)
    z = torch.matmul(y, A)
    return z.sum(dim=1).view(n, c_out, t, v)


# --------------------------------------------------------------------------- #
# ST-GCN layer / model (stgcn.py)
# --------------------------------------------------------------------------- #
def stgcn_layer(x, A, sd, prefix='', stride=1, residual=True,
                normalization='LayerNorm'):
    """One ``StgcnLayer.forward`` (stgcn.py:181-193).

    ``sd`` holds the layer's tensors under ``prefix`` with the reference key
    names: gcn.conv.{weight,bias}, tcn.0.*, tcn.2.*, tcn.3.*, residual.0.*,
    residual.1.*.
    """
    g = lambda k: sd[prefix + k]
    c_in = x.shape[1]
    c_out = g('tcn.2.weight').shape[0]
    gamma = g('tcn.2.weight').shape[2]

    if not residual:
        res = 0.0
    elif c_in == c_out and stride == 1:
        res = x
    else:
        res = F.conv2d(x, g('residual.0.weight'), g('residual.0.bias'),
                       stride=(stride, 1))
        res = _norm(normalization, res, g('residual.1.weight'),
                    g('residual.1.bias'))

    z = graph_conv(x, g('gcn.conv.weight'), g('gcn.conv.bias'), A)
    u = torch.relu(_norm(normalization, z, g('tcn.0.weight'), g('tcn.0.bias')))
    q = F.conv2d(u, g('tcn.2.weight'), g('tcn.2.bias'), stride=(stride, 1),
                 padding=((gamma - 1) // 2, 0))
    q = _norm(normalization, q, g('tcn.3.weight'), g('tcn.3.bias'))
    return torch.relu(q + res)


def stgcn_model(x, sd, cfg, return_features=False):
    """``models.stgcn.Model.forward`` (stgcn.py:80-97).

    cfg: dict(layers, stride[], residual[], normalization, importance).
    Returns logits ``(N, classes, 1)`` (and the pre-pool trunk features).
    """
    norm = cfg['normalization']
    if norm == 'LayerNorm':
        h = layer_norm_cv(x, sd['norm_in.weight'], sd['norm_in.bias'])
    else:
        h = batch_norm_input(x, sd['norm_in.norm.weight'],
                             sd['norm_in.norm.bias'])
    h = F.conv2d(h, sd['fcn_in.weight'], sd['fcn_in.bias'])
    A = sd['A']
    for i in range(cfg['layers']):
        imp = sd['edge_importance.%d' % i] if cfg.get('importance', True) else 1
        h = stgcn_layer(h, A * imp, sd, 'gcn_networks.%d.' % i,
                        stride=cfg['stride'][i],
                        residual=bool(cfg['residual'][i]),
                        normalization=norm)
    pooled = F.avg_pool2d(h, h.shape[2:])
    logits = F.conv2d(pooled, sd['fcn_out.weight'], sd['fcn_out.bias'])
    logits = logits.squeeze(-1)
    return (logits, h) if return_features else logits


# --------------------------------------------------------------------------- #
# RT-ST-GCN continual (online) path (rtstgcn.py OnlineLayer + AggregateStgcn)
# --------------------------------------------------------------------------- #
def rt_state_init(cfg, sd, batch, dtype=torch.float32):
    """Zero FIFO / accumulator state per layer (rtstgcn.py:576-579), batched.

    fifo[l]: (B, C_out, F, V) with F = stride*(kernel-1)+1; acc[l]:
    (B, C_out, stride, V); one (fifo_idx, acc_idx) pair per layer.
    """
    v = sd['A'].shape[-1]
    state = []
    for i in range(cfg['layers']):
        s = cfg['stride'][i]
        f = s * (cfg['kernel'] - 1) + 1
        c = cfg['out_ch'][i]
        state.append({
            'fifo': torch.zeros(batch, c, f, v, dtype=dtype),
            'acc': torch.zeros(batch, c, s, v, dtype=dtype),
            'fi': 0, 'ai': 0})
    return state


def rt_layer_step(x, A_eff, sd, prefix, st, c_out, stride, residual=True):
    """``OnlineLayer.forward`` for one frame ``x (B, C_in, 1, V)``.

    A_eff is ``aggregate.A`` after ``eval_()`` (A * edge_importance,
    rtstgcn.py:522-525).  Update order follows rtstgcn.py:611-625 exactly:
    ``acc <- (acc + z) + (-fifo[fi])``; output ``acc``; ``fifo[fi] <- z``.
    LayerNorm only (BatchNorm cannot process one frame).
    """
    g = lambda k: sd[prefix + k]
    b, c_in, _, v = x.shape
    k = A_eff.shape[0]
    if not residual:
        res = None
    elif c_in == c_out and stride == 1:
        res = x
    else:
        res = F.conv2d(x, g('residual.0.weight'))          # no bias, no stride
        res = layer_norm_cv(res, g('residual.1.weight'), g('residual.1.bias'))

    y = F.conv2d(x, g('conv.weight'), g('conv.bias'))        # (B, K*C, 1, V)
    y = y.view(b, k, c_out, v)                               # split over partitions
    z = torch.matmul(y, A_eff).sum(dim=1)                    # (B, C, V)

    ai, fi = st['ai'], st['fi']
    a = st['acc'][:, :, ai] + z
    a = a + (-st['fifo'][:, :, fi])
    st['acc'][:, :, ai] = a
    st['fifo'][:, :, fi] = z
    st['ai'] = (ai + 1) % st['acc'].shape[2]
    st['fi'] = (fi + 1) % st['fifo'].shape[2]

    o = a.unsqueeze(2)                                       # (B, C, 1, V)
    o = torch.relu(layer_norm_cv(o, g('bn_relu.0.weight'), g('bn_relu.0.bias')))
    if res is None:
        return o
    return torch.relu(o + res)


def rt_model_step(x, sd, cfg, state):
    """``models.rtstgcn.Model.forward`` on one frame ``(B, C, 1, V)`` with
    online layers (rtstgcn.py:137-157).  Returns logits ``(B, classes, 1)``.
    """
    h = layer_norm_cv(x, sd['norm_in.weight'], sd['norm_in.bias'])
    h = F.conv2d(h, sd['fcn_in.weight'], sd['fcn_in.bias'])
    A = sd['A']
    for i in range(cfg['layers']):
        p = 'st_gcn.%d.' % i
        imp = sd[p + 'edge_importance'] if cfg.get('importance', True) else 1
        h = rt_layer_step(h, A * imp, sd, p, state[i], cfg['out_ch'][i],
                          cfg['stride'][i], bool(cfg['residual'][i]))
    pooled = h.mean(dim=3, keepdim=True)                     # AvgPool2d((1, V))
    logits = F.conv2d(pooled, sd['fcn_out.weight'], sd['fcn_out.bias'])
    return logits.squeeze(-1)


def rt_model_run(x_seq, sd, cfg):
    """Feed ``x_seq (B, C, L, V)`` frame by frame; logits ``(B, classes, L)``."""
    state = rt_state_init(cfg, sd, x_seq.shape[0], x_seq.dtype)
    outs = [rt_model_step(x_seq[:, :, t:t + 1], sd, cfg, state)
            for t in range(x_seq.shape[2])]
    return torch.cat(outs, dim=2)


def rt_offline_layer(x, A_eff, sd, prefix, c_out, kernel, stride,
                     residual=True):
    """``OfflineLayer.forward`` (rtstgcn.py:343-389) with the local-variable
    fix; causal band-matrix sum of ``kernel // stride`` taps spaced ``stride``.
    Used only to check online == offline for stride 1.
    """
    g = lambda k: sd[prefix + k]
    n, c_in, L, v = x.shape
    k = A_eff.shape[0]
    if not residual:
        res = 0.0
    elif c_in == c_out and stride == 1:
        res = x
    else:
        res = F.conv2d(x, g('residual.0.weight'))
        res = layer_norm_cv(res, g('residual.1.weight'), g('residual.1.bias'))
    y = F.conv2d(x, g('conv.weight'), g('conv.bias')).view(n, k, c_out, L, v)
    z = torch.einsum('nkclv,kvw->nclw', y, A_eff)
    band = torch.zeros(L, L, dtype=x.dtype)
    for i in range(kernel // stride):
        band += torch.diag(torch.ones(L - stride * i, dtype=x.dtype), stride * i)
    o = torch.einsum('nclw,lm->ncmw', z, band)
    o = torch.relu(layer_norm_cv(o, g('bn_relu.0.weight'), g('bn_relu.0.bias')))
    if not residual:
        return o
    return torch.relu(o + res)
