#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c40_rt.log
: > $L
timeout 300 python tools/bench_rt.py --streams 2,4,8,9,10,12,14,16 --cuda-graph --steps 300 >> $L 2>&1
echo "== batched path (small_batch_kernel off)" >> $L
timeout 300 python - >> $L 2>&1 <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
pkg = importlib.import_module('realtime-st-gcn_b200')
syn = pkg.synthetic
dev = torch.device('cuda:0')
for b in (2, 4, 8, 9, 12, 16):
    cfg = syn.arch_config('rt-st-gcn'); cfg['math'] = 'bf16x3'; cfg['small_batch_kernel'] = False
    m = pkg.RtStgcn(**cfg); m.load_state_dict(syn.synth_state_dict(m.state_dict(), 61)); m = m.to(dev)
    m.prepare_benchmark({}); m.enable_cuda_graph(True)
    fr = torch.randn(8, b, 3, 1, 25, device=dev)
    for i in range(20): m.step(fr[i % 8])
    torch.cuda.synchronize()
    g = m._graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): g['graph'].replay()
    e1.record(); torch.cuda.synchronize()
    print('batched streams=%d back-to-back %.4f ms' % (b, e0.elapsed_time(e1) / 200))
PY
echo done
