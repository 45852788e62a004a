"""ST-GCN models of growing depth on the GPU path vs the oracle (GPU box only); run by
tests/test_gpu_parity.py in child processes with the graph-conv path switches set."""
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

cases = [([64], [64], [1]), ([64, 64], [64, 64], [1, 1]), ([64, 64], [64, 128], [1, 2]),
         ([64, 64, 128], [64, 128, 128], [1, 2, 1]), ([64, 128, 256], [128, 256, 256], [1, 2, 1])]
for math in ('bf16x3', 'bf16'):
    for in_ch, out_ch, stride in cases:
        for T in (20, 37):
            cfg = syn.arch_config('st-gcn', num_classes=12, in_ch=in_ch, out_ch=out_ch, stride=stride)
            cfg['math'] = math
            m = pkg.Stgcn(**cfg)
            sd = syn.synth_state_dict(m.state_dict(), 3)
            m.load_state_dict(sd)
            m = m.to(dev).eval()
            x = syn.synth_input((2, 3, T, 25), 4)
            logits, feats = m(x.to(dev), return_features=True)
            ocfg = dict(layers=len(in_ch), stride=stride, residual=[1] * len(in_ch), importance=True, normalization='LayerNorm')
            rl, rf = O.stgcn_model(x, sd, ocfg, return_features=True)
            d = (feats.cpu() - rf).abs()
            worst = d.flatten().argmax().item()
            idx = []
            for s in reversed(feats.shape):
                idx.append(worst % s); worst //= s
            print(math, in_ch, out_ch, stride, 'T', T, 'logits %.2e feats %.2e worst (n,c,t,v)=%s' % (
                rel(logits, rl), rel(feats, rf), tuple(reversed(idx))), flush=True)
            if rel(feats, rf) > 1e-3 and math == 'bf16x3':
                bad = (d > 1e-3 * rf.abs().max())
                print('   bad frac %.3f; bad per t:' % bad.float().mean().item(), bad.float().mean(dim=(0, 1, 3)).tolist()[:40])
                print('   bad per v:', bad.float().mean(dim=(0, 1, 2)).tolist())
                print('   bad per c (first 16 of %d):' % feats.shape[1], bad.float().mean(dim=(0, 2, 3)).tolist()[:16])
