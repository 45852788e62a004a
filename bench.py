#!/usr/bin/env python
"""Benchmark of the B200-native ST-GCN / RT-ST-GCN forward path.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Headline metric (BASELINE.json): ST-GCN forward skeleton-frames/s.  Workload = BASELINE config 3
(ST-GCN, 9 layers, PKU-MMD 25-joint graph, LayerNorm, N=256 synthetic trials x T=4000 per GPU,
trial-sharded, no collective => weak scaling).  One step = one forward over the whole batch.
`value` is timed with the input resident in HBM; `e2e` goes through the host-buffer C-ABI entry
(pinned host input -> H2D -> forward -> D2H logits) every step.  Extra keys: `roofline`
(dominant kernel class, timed live with CUDA events), `cpu_baseline` (oracle on host cores, N=1
only), `rt` (RT-ST-GCN continual p50 step latency at 1 and 4096 streams, N=1 only), `clocks`.
"""
import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "stgcn_fwd_skeleton_frames_per_s"
UNIT = "frames/s"
# algorithmic work per input frame for the PKU trunk (52 classes)
FLOP_PER_FRAME = 54.87e6


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=5)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    p.add_argument('--trials', type=int, default=256, help='trials per GPU')
    p.add_argument('--frames', type=int, default=4000, help='frames per trial')
    p.add_argument('--math', default=os.environ.get('STGCN_MATH', 'bf16x3'), choices=['fp32', 'bf16x3', 'bf16'])
    p.add_argument('--norm', default='LayerNorm', choices=['LayerNorm', 'BatchNorm'])
    p.add_argument('--rt-streams', type=int, default=4096)
    p.add_argument('--rt-steps', type=int, default=200)
    p.add_argument('--no-rt', action='store_true')
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--no-bf16-leg', action='store_true', help='skip the extra single-product bf16 measurement')
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'],
                    bf16_tflops_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def trunk_flops(T, V=25, K=3, gamma=9, in_feat=3, classes=52):
    """Per-trial algorithmic FLOPs by kernel class for the PKU trunk (SURVEY.md §8d formula)."""
    syn = importlib.import_module('realtime-st-gcn_b200.synthetic')
    f = {'gemm_1x1': 0.0, 'gemm_tcn': 0.0, 'frame': 0.0, 'embed': 2.0 * T * V * in_feat * 64,
         'pool_fc': 2.0 * 256 * classes}
    t = T
    for ci, co, s in zip(syn.TRUNK_IN, syn.TRUNK_OUT, syn.TRUNK_STRIDE):
        to = (t - 1) // s + 1
        f['gemm_1x1'] += 2.0 * t * V * ci * K * co
        f['frame'] += 2.0 * t * V * V * K * co              # dense-equivalent adjacency MACs
        f['gemm_tcn'] += 2.0 * to * V * gamma * co * co
        if ci != co or s != 1:
            f['gemm_1x1'] += 2.0 * to * V * ci * co
        t = to
    return f


def ncu_traffic(kernel_class):
    """dram bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full`
    capture (profiles/r01_ncu_top_kernel.json, written by tools/ncu_summary.py), or None."""
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_top_kernel.json')
    try:
        d = json.load(open(path))
        return d.get(kernel_class, {}).get('dram_bytes_per_launch')
    except (OSError, ValueError):
        return None


def oracle_cfg(syn, norm):
    return dict(layers=9, stride=syn.TRUNK_STRIDE, residual=[1] * 9, importance=True, normalization=norm)


def cpu_reference_step(x, sd, cfg):
    from oracle import stgcn_oracle as O
    with torch.no_grad():
        return O.stgcn_model(x, sd, cfg)


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port: same ATen CPU ops) on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pkg = importlib.import_module('realtime-st-gcn_b200')
    syn = pkg.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = syn.arch_config('st-gcn', normalization=args.norm)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 1234)
    x = syn.synth_input((1, 3, args.frames, 25), 99)           # one trial of the workload per step
    ocfg = oracle_cfg(syn, args.norm)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(x, sd, ocfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(x, sd, ocfg)
    dt = (time.perf_counter() - t0) / args.steps
    value = args.frames / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ST-GCN fwd, PKU-MMD graph, %s, T=%d V=25 (BASELINE config 3)" % (args.norm, args.frames),
                   "trials_per_gpu": args.trials, "frames_per_trial": args.frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 trial x T=%d per step (of %d trials), torch CPU fp32" % (args.frames, args.trials)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def rt_latency(pkg, dev, streams, steps, graph_kw, math, hbm_gbs=None, cuda_graph=True):
    """RT-ST-GCN continual step latency (one frame for every stream), CUDA-event timed per step."""
    syn = pkg.synthetic
    cfg = syn.arch_config('rt-st-gcn', **graph_kw)
    cfg['math'] = math
    m = pkg.RtStgcn(**cfg)
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), 61))
    m = m.to(dev)
    m.prepare_benchmark({})
    m.enable_cuda_graph(cuda_graph)
    v, c = cfg['graph']['num_node'], cfg['in_feat']
    frames = torch.randn(8, streams, c, 1, v, device=dev)
    for i in range(20):                                       # FIFO fill is 17 frames
        m.step(frames[i % 8])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        ev[i][0].record()
        m.step(frames[i % 8])
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    p50 = ms[len(ms) // 2]
    # state traffic per stream-frame = 4 * sum(C_out) * V * 4 B (read slot, write slot, read+write acc)
    state_bytes = 4 * sum(cfg['rt-st-gcn']['out_ch']) * v * 4 * streams
    out = {"streams": streams, "p50_ms": p50, "p90_ms": ms[int(len(ms) * 0.9)],
           "stream_frames_per_s": streams / (p50 * 1e-3), "cuda_graph": bool(cuda_graph),
           "state_gb_per_step": state_bytes / 1e9, "achieved_gbs": state_bytes / (p50 * 1e-3) / 1e9}
    if hbm_gbs:
        out["roofline"] = {"bound": "hbm", "achieved": out["achieved_gbs"], "peak": hbm_gbs, "unit": "GB/s",
                           "frac": out["achieved_gbs"] / hbm_gbs}
    return out


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    pkg = importlib.import_module('realtime-st-gcn_b200')
    syn, lib = pkg.synthetic, pkg._lib.load()
    cfg = syn.arch_config('st-gcn', normalization=args.norm)
    cfg['math'] = args.math
    model = pkg.Stgcn(**cfg)
    sd = syn.synth_state_dict(model.state_dict(), 1234)
    model.load_state_dict(sd)
    model = model.to(dev).eval()

    N, T, V = args.trials, args.frames, 25
    x_host = syn.synth_input((N, 3, T, V), 1000 + rank).pin_memory()
    x = x_host.to(dev)
    frames_per_step = N * T

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    out = None

    def step():
        nonlocal out
        out = model(x)

    for _ in range(args.warmup):
        step()
    l0 = lib.stgcn_launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(step, args.steps)
    launches = lib.stgcn_launch_count() - l0
    value = world * frames_per_step / (ms * 1e-3)

    # ---- e2e: host-buffer C-ABI entry, H2D + forward + D2H every step ----
    e2e = None
    if not args.no_e2e:
        desc, _ = model._descriptor()
        ws = torch.empty(lib.stgcn_model_workspace_bytes(ctypes.byref(desc), N, T), dtype=torch.uint8, device=dev)
        io = torch.empty(x_host.numel() + N * 52, device=dev)
        logits_host = torch.empty(N, 52).pin_memory()
        stream = torch.cuda.current_stream(dev).cuda_stream

        def step_host():
            pkg._lib.check(lib.stgcn_model_forward_host(
                ctypes.byref(desc), x_host.data_ptr(), logits_host.data_ptr(), N, T, io.data_ptr(),
                ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)))

        step_host()
        ms_e2e = timed(step_host, args.steps)
        assert torch.equal(logits_host, out.squeeze(-1).cpu()), "e2e path disagrees with device path"
        e2e = {"value": world * frames_per_step / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": logits_host.numel() * 4,
               "ms_per_step": ms_e2e}
        del ws, io

    # ---- roofline of the dominant kernel class: one event-bracketed step (outside the timed region) ----
    pk = peaks()
    roofline = None
    if rank == 0:
        ncls = len(pkg._lib.KERNEL_CLASSES)
        cls_ms = (ctypes.c_float * ncls)()
        cls_n = (ctypes.c_longlong * ncls)()
        lib.stgcn_profile_begin()
        step()
        pkg._lib.check(lib.stgcn_profile_end(cls_ms, cls_n, ncls))
        shares = {n: cls_ms[i] for i, n in enumerate(pkg._lib.KERNEL_CLASSES) if cls_n[i]}
        total = sum(shares.values())
        dom = max(shares, key=shares.get)
        flops = trunk_flops(T)
        dom_flops = flops.get(dom, 0.0) * N                    # algorithmic FLOPs of that class per step
        i = pkg._lib.KERNEL_CLASSES.index(dom)
        ach = dom_flops / (cls_ms[i] * 1e-3) / 1e12
        mma_per_product = {'bf16x3': 3, 'bf16': 1, 'fp32': 0}[args.math]
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk['bf16_tflops_sustained'],
                    "unit": "TFLOP/s", "frac": ach / pk['bf16_tflops_sustained'], "traffic": ncu_traffic(dom),
                    "mma_per_product": mma_per_product,
                    "tensor_pipe_frac": (ach * mma_per_product / pk['bf16_tflops_sustained']) if mma_per_product else None,
                    "note": "achieved = algorithmic FLOPs (SURVEY 8d) of this kernel class / its CUDA-event time; "
                            "fp32-parity mode issues 3 bf16 MMAs per product, so the tensor pipe is busy "
                            "tensor_pipe_frac of the measured peak",
                    "peak_source": pk['source'] + " bf16 sustained (kernel timed inside a long step)",
                    "launches": int(cls_n[i]), "avg_launch_ms": cls_ms[i] / max(cls_n[i], 1),
                    "share_of_step": cls_ms[i] / total if total else None,
                    "class_ms": {k: round(v, 3) for k, v in shares.items()},
                    "whole_step_tflops": FLOP_PER_FRAME * frames_per_step / (ms * 1e-3) / 1e12}

    # ---- the same workload in single-product bf16 mode (north_star: "a stated bf16 tolerance when that mode is
    # enabled"): reported next to the fp32-parity headline, never instead of it ----
    bf16_leg = None
    if rank == 0 and world == 1 and args.math == 'bf16x3' and not args.no_bf16_leg:
        cfg16 = dict(cfg)
        cfg16['math'] = 'bf16'
        m16 = pkg.Stgcn(**cfg16)
        m16.load_state_dict(sd)
        m16 = m16.to(dev).eval()
        o16 = None

        def step16():
            nonlocal o16
            o16 = m16(x)

        for _ in range(2):
            step16()
        ms16 = timed(step16, 3)
        ref = out.float()
        bf16_leg = {"value": frames_per_step / (ms16 * 1e-3), "unit": UNIT, "ms_per_step": ms16, "steps": 3,
                    "max_rel_diff_vs_parity_mode": float((o16.float() - ref).abs().max() / ref.abs().max()),
                    "stated_tolerance": "3e-2 relative on logits vs the fp32 reference (tests/test_gpu_parity.py BF16_TOL)"}
        del m16, o16

    rt = None
    if rank == 0 and world == 1 and not args.no_rt:
        del x
        torch.cuda.empty_cache()
        imu = dict(graph='imu_fogit_ABCD', in_feat=6, num_classes=8)
        rt = {"pku": [rt_latency(pkg, dev, b, args.rt_steps, {}, args.math, pk['hbm_gbs'])
                      for b in (1, args.rt_streams)],
              "imu_bf16": [rt_latency(pkg, dev, b, args.rt_steps, imu, 'bf16', pk['hbm_gbs'])
                           for b in (1, args.rt_streams)]}
        rt["note"] = ("BASELINE configs 2 and 5: one continual step for all streams (LayerNorm, fp32 FIFO/accumulator "
                      "state, CUDA-graph replay), p50 over %d steps after a 20-step FIFO fill; pku math=%s, imu "
                      "math=bf16 (stated tolerance 3e-2 rel., tests/test_gpu_parity.py)" % (args.rt_steps, args.math))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        xs = x_host[:1].clone()
        ocfg = oracle_cfg(syn, args.norm)
        cpu_reference_step(xs, sd, ocfg)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            cpu_reference_step(xs, sd, ocfg)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": T / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d x (1 trial, T=%d) of the %d-trial workload; oracle = same ATen CPU ops as the reference"
                         % (reps, T, N)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3(f32-parity)", "bf16": "bf16"}[args.math],
            "data": "synthetic",
            "config": {"workload": "ST-GCN fwd, 9 layers, PKU-MMD 25-joint graph, %s, N=%d trials x T=%d per GPU "
                                   "(BASELINE config 3), trial-sharded, no collective" % (args.norm, N, T),
                       "trials_per_gpu": N, "frames_per_trial": T, "math": args.math,
                       "l2": "inputs and activations (>= 0.3 GB per tensor) exceed the 126 MB L2"},
            "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "rt": rt,
            "bf16_mode": bf16_leg,
            "clocks": clocks.summary(),
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
