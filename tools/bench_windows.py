"""Sliding-window inference timing (GPU box): one trial of L frames, windows of W frames, PKU trunk.
    python tools/bench_windows.py [--L 4000] [--W 50] [--math bf16x3]
Run with STGCN_WINDOWS_SHARE=0 for the per-window evaluation of the first layer."""
import argparse, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('realtime-st-gcn_b200')
syn = pkg.synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--L', type=int, default=4000)
ap.add_argument('--W', type=int, default=50)
ap.add_argument('--math', default='bf16x3')
a = ap.parse_args()
dev = torch.device('cuda:0')
cfg = syn.arch_config('st-gcn')
cfg['math'] = a.math
m = pkg.Stgcn(**cfg)
m.load_state_dict(syn.synth_state_dict(m.state_dict(), 1))
m = m.to(dev).eval()
cap = syn.synth_input((1, 3, a.L, 25), 2).to(dev)
for _ in range(3):
    out = m.forward_windows(cap, a.W)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for s, e in ev:
    s.record(); out = m.forward_windows(cap, a.W); e.record()
torch.cuda.synchronize()
ms = sorted(s.elapsed_time(e) for s, e in ev)
print('L=%d W=%d math=%s share=%s p50 %.3f ms  (%.0f windows/s)  checksum %.6f' % (
    a.L, a.W, a.math, os.environ.get('STGCN_WINDOWS_SHARE', '1'), ms[len(ms) // 2], a.L / (ms[len(ms) // 2] * 1e-3),
    float(out.double().abs().sum())))
