"""Host harness for the continual path: the B200 counterpart of the reference's
``Processor._forward_rt`` (processor.py:395-427) and the FP32 leg of ``Processor.benchmark``
(processor.py:870-902, 977).

Differences that do not change results: the per-frame timer is a pair of CUDA events on the
launching stream (the reference wraps ``time.time()`` around an unsynchronised call, which measures
nothing on a GPU); ``torch.jit.script`` is skipped (a ctypes-backed forward needs no scripting);
the INT8 FX-quantisation leg is out of scope, so ``latency_int8`` is written as NaN to keep the
reference's ``latency.csv`` schema.  Top-1/top-5 follow ``utils/statistics.py:4-16``.
"""
import os

import torch


def statistics(predictions, labels):
    """``Statistics.__call__`` (utils/statistics.py:4-16): predictions (N, classes, L), labels (N, L)."""
    k = min(5, predictions.shape[1])
    _, top5 = torch.topk(predictions, k=k, dim=1)
    top1 = top5[:, 0, :]
    top1_cor = int((top1 == labels).sum().item())
    top5_cor = int((top5 == labels[:, None]).sum().item())
    return top1, top5, top1_cor, top5_cor, labels.numel()


@torch.no_grad()
def forward_rt(model, captures, labels=None):
    """Frame-by-frame continual forward over ``captures (N, C, L, V)`` (``get_segment_rt``,
    utils/segment_generator.py:79-81, yields ``captures[:, :, i:i+1]``).  Returns a dict with the
    predictions ``(N, classes, L)``, the reference's metric ``latency`` = seconds per frame (mean),
    the per-frame device times, and top-1/top-5 counts when ``labels`` are given."""
    n, _, length, _ = captures.shape
    dev = captures.device
    predictions = None
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(length)]
    # with labels, rank the classes where the logits are produced (rtstgcn_step_top5) instead of a separate
    # torch.topk pass over the predictions
    fused = labels is not None and hasattr(model, 'step_top5') and getattr(model, 'is_online', False) \
        and model.num_classes >= 5
    ranks = torch.empty((n, 5, length), device=dev, dtype=torch.int64) if fused else None
    for i in range(length):
        frame = captures[:, :, i:i + 1]
        events[i][0].record()
        if fused:
            out, top = model.step_top5(frame)
            ranks[:, :, i] = top
        else:
            out = model(frame)
        events[i][1].record()
        if predictions is None:
            predictions = torch.empty((n, out.shape[1], length), device=dev, dtype=out.dtype)
        predictions[:, :, i:i + 1] = out
    torch.cuda.synchronize(dev)
    frame_ms = [a.elapsed_time(b) for a, b in events]
    res = {"predictions": predictions, "latency": sum(frame_ms) / length * 1e-3, "frame_ms": frame_ms,
           "p50_ms": sorted(frame_ms)[length // 2]}
    if labels is not None:
        lab = labels.to(dev)
        if fused:
            top1 = ranks[:, 0, :]
            res.update(top1_predicted=top1, top5_predicted=ranks, top1_cor=int((top1 == lab).sum().item()),
                       top5_cor=int((ranks == lab[:, None]).sum().item()), tot=lab.numel())
        else:
            top1, top5, c1, c5, tot = statistics(predictions, lab)
            res.update(top1_predicted=top1, top5_predicted=top5, top1_cor=c1, top5_cor=c5, tot=tot)
    return res


def benchmark(model, captures, labels=None, arch_conf=None, save_dir=None, log=None, cuda_graph=True):
    """FP32 leg of ``Processor.benchmark``: swap to the inference-only layers, run the continual
    loop on one trial batch, print and save seconds-per-frame in the reference's format."""
    model.prepare_benchmark(arch_conf if arch_conf is not None else {})
    model.eval()
    if hasattr(model, 'enable_cuda_graph'):
        model.enable_cuda_graph(cuda_graph)
    if hasattr(model, 'reset_streams'):
        model.reset_streams()
    res = forward_rt(model, captures, labels)
    print("[benchmark FP32]: {0} spf".format(res["latency"]), flush=True, file=log)
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        with open(os.path.join(save_dir, 'latency.csv'), 'w') as f:       # processor.py:977
            f.write(",latency_fp32,latency_int8\n0,%r,nan\n" % res["latency"])
        if labels is not None:
            with open(os.path.join(save_dir, 'accuracy.csv'), 'w') as f:  # processor.py:958-965 (fp32 columns)
                f.write(",top1_fp32,top1_int8,top5_fp32,top5_int8\n0,%r,nan,%r,nan\n"
                        % (res["top1_cor"] / res["tot"], res["top5_cor"] / res["tot"]))
    return res
