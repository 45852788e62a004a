"""Oracle restatement of the CoST-GCN continual step (CPU, torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference lines followed (paths relative to the reference root):
  models/costgcn/costgcn.py:81-99     -> cost_model_step  (norm_in, fcn_in, layers, avg_pool2d over
                                         the single frame's (1, V), fcn_out)
  models/costgcn/costgcn.py:125-211   -> cost_layer_step  (per layer: delayed residual FIFO, FIFO of
                                         graph-convolved frames, LayerNorm + ReLU applied to the WHOLE
                                         FIFO each step, Gamma x 1 convolution with dilation = stride
                                         and 'valid' padding -> one output frame, second LayerNorm,
                                         + residual of Gamma // 2 frames ago, ReLU)

Two properties of the reference that the restatement keeps (and the CUDA path must reproduce):
  * the FIFOs start as zeros and the first LayerNorm runs over FIFO frames that were never written:
    LN(0) = bias, so an empty slot contributes relu(tcn.0.bias), not zero;
  * the convolution is a cross-correlation over a newest-first FIFO: tap j multiplies the frame
    j * stride steps in the PAST (tap 0 = the current frame).
The reference is batch-1 with plain-tensor state; here the state carries a batch dimension (streams are
independent because every normalisation is per frame).
"""
import torch
import torch.nn.functional as F

from .stgcn_oracle import graph_conv, layer_norm_cv


def cost_state_init(cfg, sd, batch, dtype=torch.float32):
    """Per layer: ``fifo`` and ``fifo_res``, both ``(B, C_out, F, V)`` zeros, F = stride*(kernel-1)+1
    (costgcn.py:150-154)."""
    v = sd['A'].shape[-1]
    state = []
    for i in range(cfg['layers']):
        c_out = sd['gcn_networks.%d.tcn.2.weight' % i].shape[0]
        f = cfg['stride'][i] * (cfg['kernel'] - 1) + 1
        state.append({'fifo': torch.zeros(batch, c_out, f, v, dtype=dtype),
                      'fifo_res': torch.zeros(batch, c_out, f, v, dtype=dtype)})
    return state


def cost_layer_step(x, A_eff, sd, prefix, st, stride, residual=True):
    """One ``StgcnLayer.forward`` of costgcn.py:190-211 for a frame ``x (B, C_in, 1, V)``."""
    g = lambda k: sd[prefix + k]
    c_in = x.shape[1]
    c_out = g('tcn.2.weight').shape[0]
    gamma = g('tcn.2.weight').shape[2]
    if not residual:
        res = x * 0.0 if c_in == c_out else torch.zeros(x.shape[0], c_out, 1, x.shape[3], dtype=x.dtype)
    elif c_in == c_out and stride == 1:
        res = x
    else:
        res = F.conv2d(x, g('residual.0.weight'), g('residual.0.bias'))
        res = layer_norm_cv(res, g('residual.1.weight'), g('residual.1.bias'))
    st['fifo_res'] = torch.cat((res, st['fifo_res'][:, :, :-1]), dim=2)
    z = graph_conv(x, g('gcn.conv.weight'), g('gcn.conv.bias'), A_eff)
    st['fifo'] = torch.cat((z, st['fifo'][:, :, :-1]), dim=2)
    u = torch.relu(layer_norm_cv(st['fifo'], g('tcn.0.weight'), g('tcn.0.bias')))
    q = F.conv2d(u, g('tcn.2.weight'), g('tcn.2.bias'), dilation=(stride, 1))      # 'valid': one frame
    q = layer_norm_cv(q, g('tcn.3.weight'), g('tcn.3.bias'))
    return torch.relu(q + st['fifo_res'][:, :, gamma // 2:gamma // 2 + 1])


def cost_model_step(x, sd, cfg, state):
    """``Model.forward`` (costgcn.py:81-99) on one frame ``(B, C, 1, V)`` -> ``(B, classes, 1)``."""
    h = layer_norm_cv(x, sd['norm_in.weight'], sd['norm_in.bias'])
    h = F.conv2d(h, sd['fcn_in.weight'], sd['fcn_in.bias'])
    for i in range(cfg['layers']):
        a = sd['A'] * sd['edge_importance.%d' % i] if cfg.get('importance', True) else sd['A']
        h = cost_layer_step(h, a, sd, 'gcn_networks.%d.' % i, state[i], cfg['stride'][i],
                            bool(cfg['residual'][i]))
    h = F.avg_pool2d(h, h.shape[2:])
    return F.conv2d(h, sd['fcn_out.weight'], sd['fcn_out.bias']).squeeze(-1)


def cost_model_run(x_seq, sd, cfg):
    """Feed ``x_seq (B, C, L, V)`` frame by frame; logits ``(B, classes, L)``."""
    state = cost_state_init(cfg, sd, x_seq.shape[0], x_seq.dtype)
    return torch.cat([cost_model_step(x_seq[:, :, t:t + 1], sd, cfg, state) for t in range(x_seq.shape[2])],
                     dim=2)
