#!/usr/bin/env python
"""BASELINE config 4: ST-GCN forward on ONE long synthetic trial, T-partitioned across the ranks of
a torchrun launch (NCCL send/recv halo exchange per layer + pooled all-reduce).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_tsplit.py [--frames 262144] [--steps 5] [--check]

Prints one JSON line on rank 0: frames/s (max over ranks), halo bytes per step, and with --check
the relative error against a single-GPU forward of the whole trial on rank 0."""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=262144)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--check', action='store_true')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pkg = importlib.import_module('realtime-st-gcn_b200')
    syn, ts, lib = pkg.synthetic, pkg.tsplit, pkg._lib.load()
    cfg = syn.arch_config('st-gcn')
    cfg['math'] = a.math
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), 1234))
    m = m.to(dev).eval()
    import ctypes
    desc, _ = m._descriptor()
    ex = ts.DistExchange(rank, world, lib.stgcn_model_halo_bytes(ctypes.byref(desc), 1), dev)
    x = syn.synth_input((1, 3, a.frames, 25), 4242)            # same trial on every rank (seeded)
    s, e = ts.chunk_bounds(a.frames, world, ts.total_stride(syn.TRUNK_STRIDE))[rank]
    xl = x[:, :, s:e].contiguous().to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = m.forward_tsplit(xl, a.frames, ex)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0 = ex.bytes_sent
    e0.record()
    for _ in range(a.steps):
        out = m.forward_tsplit(xl, a.frames, ex)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = {"workload": "ST-GCN fwd, single trial T=%d V=25, T-split over %d GPU(s), per-layer NCCL halo exchange"
                       % (a.frames, world), "n_gpus": world, "math": a.math, "ms_per_step": ms.item(),
           "frames_per_s": a.frames / (ms.item() * 1e-3),
           "halo_bytes_sent_per_rank_per_step": (ex.bytes_sent - b0) // a.steps, "exchanges_per_step": 9}
    if a.check and rank == 0:
        full = m(x.to(dev))
        res["rel_err_vs_single_gpu"] = ((out - full).abs().max() / full.abs().max()).item()
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
