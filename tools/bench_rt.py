#!/usr/bin/env python
"""RT-ST-GCN continual step: p50 latency and achieved state traffic (development aid; bench.py
reports the same numbers in its `rt` key).

    python tools/bench_rt.py [--streams 1,4096] [--graph pku-mmd] [--math bf16x3] [--cuda-graph] [--steps 200]
"""
import argparse
import ctypes
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--streams', default='1,4096')
    ap.add_argument('--graph', default='pku-mmd')
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--cuda-graph', action='store_true')
    ap.add_argument('--profile', action='store_true')
    a = ap.parse_args()
    pkg = importlib.import_module('realtime-st-gcn_b200')
    syn, lib = pkg.synthetic, pkg._lib.load()
    dev = torch.device('cuda:0')
    kw = {} if a.graph == 'pku-mmd' else dict(graph=a.graph, in_feat=6, num_classes=8)
    for b in [int(s) for s in a.streams.split(',')]:
        cfg = syn.arch_config('rt-st-gcn', **kw)
        cfg['math'] = a.math
        m = pkg.RtStgcn(**cfg)
        m.load_state_dict(syn.synth_state_dict(m.state_dict(), 61))
        m = m.to(dev)
        m.prepare_benchmark({})
        m.enable_cuda_graph(a.cuda_graph)
        v, c = cfg['graph']['num_node'], cfg['in_feat']
        frames = torch.randn(8, b, c, 1, v, device=dev)
        for i in range(20):
            m.step(frames[i % 8])
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for i in range(a.steps):
            ev[i][0].record()
            m.step(frames[i % 8])
            ev[i][1].record()
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in ev)
        p50 = ms[len(ms) // 2]
        sum_cout = sum(cfg['rt-st-gcn']['out_ch'])
        state_bytes = 4 * sum_cout * v * 4 * b
        line = "streams=%d math=%s graph=%s p50=%.4f ms p90=%.4f ms  state %.3f GB/step -> %.1f GB/s" % (
            b, a.math, a.cuda_graph, p50, ms[int(len(ms) * 0.9)], state_bytes / 1e9, state_bytes / p50 / 1e6)
        g = getattr(m, '_graph', None)
        if a.cuda_graph and g is not None:                 # the captured step replayed back to back
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                g['graph'].replay()
            e1.record()
            torch.cuda.synchronize()
            line += "  back-to-back %.4f ms" % (e0.elapsed_time(e1) / 200)
        if a.profile and not a.cuda_graph:
            n = len(pkg._lib.KERNEL_CLASSES)
            cms, cn = (ctypes.c_float * n)(), (ctypes.c_longlong * n)()
            lib.stgcn_profile_begin()
            for i in range(10):
                m.step(frames[i % 8])
            pkg._lib.check(lib.stgcn_profile_end(cms, cn, n))
            line += "  per-step class ms: " + ", ".join(
                "%s %.3f (%d)" % (k, cms[i] / 10, cn[i] // 10) for i, k in enumerate(pkg._lib.KERNEL_CLASSES) if cn[i])
        print(line, flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
