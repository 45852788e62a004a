"""Sliding-window inference (Stgcn.forward_windows) on the GPU vs the oracle on sampled windows; writes the full
outputs to argv[1] so that tests/test_gpu_parity.py can compare the shared-first-layer evaluation with the
per-window one (STGCN_WINDOWS_SHARE, read once per process) bit for bit.  GPU box only."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('realtime-st-gcn_b200')
from oracle import stgcn_oracle as O
syn = pkg.synthetic
dev = torch.device('cuda:0')


def rel(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max())


outs = {}
L, W = 300, 64
cases = [('res0', [0, 1, 1], 'bf16x3'), ('res1', [1, 1, 1], 'bf16x3'), ('res1', [1, 1, 1], 'bf16')]
for name, residual, math in cases:
    cfg = syn.arch_config('st-gcn', num_classes=12, in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1],
                          residual=residual)
    cfg['math'] = math
    m = pkg.Stgcn(**cfg)
    sd = syn.synth_state_dict(m.state_dict(), 91)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    cap = syn.synth_input((1, 3, L, 25), 92)
    out = m.forward_windows(cap.to(dev), W)
    torch.cuda.synchronize()
    padded = torch.nn.functional.pad(cap, (0, 0, W - 1, 0))
    pick = [0, 1, W - 2, W - 1, W, 127, 128, 129, L - 1]
    windows = torch.stack([padded[0, :, i:i + W] for i in pick])
    ref = O.stgcn_model(windows, sd, dict(layers=3, stride=[1, 2, 1], residual=residual, importance=True,
                                          normalization='LayerNorm'))
    print('%s %s err %.3e' % (math, name, rel(out[0, :, pick].t().unsqueeze(-1), ref)), flush=True)
    outs['%s_%s' % (math, name)] = out.cpu()
if len(sys.argv) > 1:
    torch.save(outs, sys.argv[1])
