"""Oracle restatement of the T-split forward (CPU, torch).  TEST INFRASTRUCTURE ONLY.

Same arithmetic as ``stgcn_oracle.stgcn_model`` (reference stgcn.py:80-97, 181-193), evaluated on
one rank's contiguous chunk of frames.  Only the Gamma x 1 temporal convolution looks across the
chunk boundary: its input ``u`` is extended by ``halo`` frames from the ring neighbours (zeros at
the sequence ends -- the conv's own zero padding) and convolved without padding, which selects
exactly the taps ``s*tau + j - pad`` of the full sequence.  ``exchange(send_left, send_right)``
returns ``(recv_left, recv_right)`` (None where there is no neighbour); ``reduce_sum`` all-reduces
the pooled sums.  With a single rank and no neighbours this reproduces ``stgcn_model`` exactly.
"""
import torch
import torch.nn.functional as F

from . import stgcn_oracle as O


def stgcn_layer_tsplit(x, A, sd, prefix, stride, residual, exchange, halo=4):
    g = lambda k: sd[prefix + k]                                             # noqa: E731
    c_in = x.shape[1]
    c_out, gamma = g('tcn.2.weight').shape[0], g('tcn.2.weight').shape[2]
    pad = (gamma - 1) // 2
    assert pad <= halo
    if not residual:
        res = 0.0
    elif c_in == c_out and stride == 1:
        res = x
    else:
        res = F.conv2d(x, g('residual.0.weight'), g('residual.0.bias'), stride=(stride, 1))
        res = O.layer_norm_cv(res, g('residual.1.weight'), g('residual.1.bias'))
    z = O.graph_conv(x, g('gcn.conv.weight'), g('gcn.conv.bias'), A)
    u = torch.relu(O.layer_norm_cv(z, g('tcn.0.weight'), g('tcn.0.bias')))
    left, right = exchange(u[:, :, :halo].contiguous(), u[:, :, -halo:].contiguous())
    zeros = torch.zeros_like(u[:, :, :halo])
    u_pad = torch.cat([zeros if left is None else left, u, zeros if right is None else right], dim=2)
    t_out = (u.shape[2] - 1) // stride + 1
    # u_pad index = original index + halo; tap j of output tau reads original s*tau + j - pad
    q = F.conv2d(u_pad[:, :, halo - pad:], g('tcn.2.weight'), g('tcn.2.bias'), stride=(stride, 1))[:, :, :t_out]
    q = O.layer_norm_cv(q, g('tcn.3.weight'), g('tcn.3.bias'))
    return torch.relu(q + res)


def stgcn_model_tsplit(x_local, sd, cfg, exchange, reduce_sum, total_frames):
    """LayerNorm only.  Returns the full-trial logits (N, classes, 1)."""
    assert cfg['normalization'] == 'LayerNorm'
    h = O.layer_norm_cv(x_local, sd['norm_in.weight'], sd['norm_in.bias'])
    h = F.conv2d(h, sd['fcn_in.weight'], sd['fcn_in.bias'])
    A = sd['A']
    for i in range(cfg['layers']):
        imp = sd['edge_importance.%d' % i] if cfg.get('importance', True) else 1
        h = stgcn_layer_tsplit(h, A * imp, sd, 'gcn_networks.%d.' % i, cfg['stride'][i],
                               bool(cfg['residual'][i]), exchange)
    sums = reduce_sum(h.sum(dim=(2, 3)))
    t_final = total_frames
    for s in cfg['stride'][:cfg['layers']]:
        t_final = (t_final - 1) // s + 1
    pooled = (sums / float(t_final * h.shape[3])).view(h.shape[0], -1, 1, 1)
    return F.conv2d(pooled, sd['fcn_out.weight'], sd['fcn_out.bias']).squeeze(-1)
