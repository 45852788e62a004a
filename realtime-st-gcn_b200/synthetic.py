"""Seeded synthetic configs, weights and inputs shared by tests, bench and the
golden-vector generator (no datasets or checkpoints ship with the reference,
SURVEY.md §4/§8d).

Weights are drawn key-by-key (sorted key order) from one ``torch.Generator`` so
the same ``state_dict`` can be loaded into the reference modules, the oracle and
the B200 modules.  Defaults that would hide bugs are randomised: norm weights
~U(0.5,1.5), norm biases ~N(0,0.1), ``edge_importance`` ~U(0.5,1.5).
"""
import hashlib
import math

import torch

from .skeletons import load_graph

TRUNK_IN = [64, 64, 64, 64, 128, 128, 128, 256, 256]
TRUNK_OUT = [64, 64, 64, 128, 128, 128, 256, 256, 256]
TRUNK_STRIDE = [1, 1, 1, 2, 1, 1, 2, 1, 1]


def arch_config(model='st-gcn', graph='pku-mmd', normalization='LayerNorm',
                in_feat=3, num_classes=52, kernel=9, in_ch=None, out_ch=None,
                stride=None, residual=None, importance=True, strategy='spatial'):
    """Reference-style ``Model(**kwargs)`` dict (config/pku-mmd/*/*.json 'arch'
    group plus the ``num_classes``/``graph`` injected at processor.py:164-166)."""
    in_ch = list(TRUNK_IN if in_ch is None else in_ch)
    out_ch = list(TRUNK_OUT if out_ch is None else out_ch)
    stride = list(TRUNK_STRIDE if stride is None else stride)
    layers = len(in_ch)
    residual = [1] * layers if residual is None else list(residual)
    sub = {
        'latency': False, 'importance': importance, 'in_feat': in_feat,
        'stages': 1, 'layers': layers, 'kernel': kernel, 'in_ch': in_ch,
        'out_ch': out_ch, 'stride': stride, 'residual': residual,
        'dropout': [0.0] * layers,
    }
    if model == 'rt-st-gcn':
        sub['buffer'] = 1
    return {
        'strategy': strategy, 'in_feat': in_feat, 'stages': 1, 'kernel': kernel,
        'normalization': normalization, 'num_classes': num_classes,
        'graph': load_graph(graph),
        model: sub,
    }


def _uniform(gen, shape, lo, hi):
    return torch.rand(shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo


def synth_state_dict(template, seed):
    """Fill a ``state_dict`` (same keys/shapes as ``template``) with seeded values.

    ``template`` is any mapping key -> tensor (only shapes are used).  The
    adjacency buffer ``A`` is kept as is.
    """
    gen = torch.Generator().manual_seed(int(seed))
    out = {}
    for key in sorted(template.keys()):
        t = template[key]
        shape = tuple(t.shape)
        leaf = key.split('.')[-1]
        if key == 'A' or key.endswith('aggregate.A'):
            out[key] = t.detach().clone().float()
        elif 'edge_importance' in key:
            out[key] = _uniform(gen, shape, 0.5, 1.5)
        elif len(shape) == 4 and leaf == 'weight':           # conv kernels
            fan_in = shape[1] * shape[2] * shape[3]
            b = 1.0 / math.sqrt(fan_in)
            out[key] = _uniform(gen, shape, -b, b)
        elif leaf == 'weight':                               # norm scales
            out[key] = _uniform(gen, shape, 0.5, 1.5)
        elif leaf == 'bias' and len(shape) == 1 and _is_conv_bias(key):
            out[key] = _uniform(gen, shape, -0.1, 0.1)
        elif leaf == 'bias':                                 # norm shifts
            out[key] = torch.randn(shape, generator=gen, dtype=torch.float32) * 0.1
        else:
            out[key] = t.detach().clone()
    return out


def _is_conv_bias(key):
    parts = ['', '', ''] + key.split('.')
    if parts[3] in ('fcn_in', 'fcn_out'):
        return True
    # gcn_networks.i.gcn.conv.bias | tcn.2.bias | residual.0.bias | st_gcn.i.conv.bias
    return (parts[-2] == 'conv') or (parts[-3:-1] == ['tcn', '2']) or \
           (parts[-3:-1] == ['residual', '0'])


def synth_input(shape, seed):
    gen = torch.Generator().manual_seed(int(seed))
    return torch.randn(shape, generator=gen, dtype=torch.float32)


def state_digest(sd):
    """Order-independent sha256 over a state dict (detects RNG drift)."""
    h = hashlib.sha256()
    for key in sorted(sd.keys()):
        h.update(key.encode())
        h.update(sd[key].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]
