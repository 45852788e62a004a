"""T-split host logic on CPU: chunking arithmetic, and a world_size-2 (and 3) `gloo` run of the
per-layer ring-neighbour halo exchange + pooled all-reduce -- the product's `tsplit.DistExchange`
moving the oracle's boundary frames -- against the single-process oracle.  Trial sharding (the
no-collective case) is covered by `test_trial_sharding_partition`."""
import importlib
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

SMALL = dict(in_ch=[16, 16, 32, 32], out_ch=[16, 32, 32, 32], stride=[1, 2, 1, 2])


def test_chunk_bounds(pkg):
    ts = pkg.tsplit
    assert ts.chunk_bounds(262144, 8, 4) == [(i * 32768, (i + 1) * 32768) for i in range(8)]
    b = ts.chunk_bounds(301, 3, 4)                    # ragged tail goes to the last chunk
    assert b[0][0] == 0 and b[-1][1] == 301
    assert all(b[i][1] == b[i + 1][0] for i in range(2))
    assert all((e - s) % 4 == 0 for s, e in b[:-1])
    assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 4 + 3
    with pytest.raises(ValueError):
        ts.chunk_bounds(10, 4, 4)
    assert ts.total_stride([1, 1, 1, 2, 1, 1, 2, 1, 1]) == 4
    assert ts.frames_after(300, [1, 1, 1, 2, 1, 1, 2, 1, 1]) == 75
    assert ts.frames_after(301, [2, 2]) == 76


def test_validate_chunks_is_collective(pkg):
    """a ragged tail of fewer than HALO_FRAMES frames (at any layer's resolution) must be
    rejected from the globally known sizes -- identically on every rank, before any launch."""
    ts = pkg.tsplit
    strides = [1, 1, 1, 2, 1, 1, 2, 1, 1]
    assert ts.validate_chunks(262144, 8, strides) == ts.chunk_bounds(262144, 8, 4)
    assert ts.validate_chunks(17, 1, strides) == [(0, 17)]          # a single rank needs no halo
    assert ts.validate_chunks(64, 4, strides)[-1] == (48, 64)       # 16 frames per rank -> 4 at the last layers
    with pytest.raises(ValueError, match='halo'):
        ts.validate_chunks(5, 2, strides)                            # the ragged tail: rank 1 would hold 1 frame
    with pytest.raises(ValueError, match='halo'):
        ts.validate_chunks(32, 4, strides)                           # 8 frames per rank -> 2 at the last layers


def test_trial_sharding_partition():
    """Trial/stream sharding needs no collective: the per-rank slices tile the batch exactly."""
    def shard(n, world, rank):
        per = -(-n // world)
        return range(min(rank * per, n), min((rank + 1) * per, n))
    for n, world in [(256, 8), (256, 3), (5, 8), (4096, 4)]:
        seen = [i for r in range(world) for i in shard(n, world, r)]
        assert seen == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total_frames, out_path):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        pkg = importlib.import_module('realtime-st-gcn_b200')
        from oracle import tsplit_oracle as TO
        syn, ts = pkg.synthetic, pkg.tsplit
        cfg = syn.arch_config('st-gcn', num_classes=12, **SMALL)
        sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 21)
        x = syn.synth_input((2, 3, total_frames, 25), 22)
        a, b = ts.chunk_bounds(total_frames, world, ts.total_stride(SMALL['stride']))[rank]
        ex = ts.DistExchange(rank, world, capacity=2 * 4 * 25 * 32 * 4, device='cpu')

        def exchange(send_l, send_r):
            # the product exchange moves raw bytes between its staging buffers: stage, swap, unstage
            n = send_l.numel() * 4
            ex.send_left[:n] = send_l.view(torch.uint8).flatten()
            ex.send_right[:n] = send_r.view(torch.uint8).flatten()
            assert ex(0, n) == 0
            left = ex.recv_left[:n].clone().view(torch.float32).view_as(send_l) if ex.has_left else None
            right = ex.recv_right[:n].clone().view(torch.float32).view_as(send_r) if ex.has_right else None
            return left, right

        ocfg = dict(layers=4, stride=SMALL['stride'], residual=[1] * 4, normalization='LayerNorm')
        logits = TO.stgcn_model_tsplit(x[:, :, a:b].contiguous(), sd, ocfg, exchange, ex.all_reduce_sum,
                                       total_frames)
        assert ex.calls == 4                                  # one exchange per layer
        if rank == 0:
            torch.save(logits, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,total_frames', [(2, 64), (2, 50), (3, 61)])
def test_tsplit_gloo_matches_single_process_oracle(pkg, syn, tmp_path, world, total_frames):
    from oracle import stgcn_oracle as O
    out = str(tmp_path / 'logits.pt')
    mp.spawn(_worker, args=(world, _free_port(), total_frames, out), nprocs=world, join=True)
    got = torch.load(out)
    cfg = syn.arch_config('st-gcn', num_classes=12, **SMALL)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 21)
    x = syn.synth_input((2, 3, total_frames, 25), 22)
    ref = O.stgcn_model(x, sd, dict(layers=4, stride=SMALL['stride'], residual=[1] * 4,
                                    normalization='LayerNorm'))
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_tsplit_oracle_single_rank_is_exact(pkg, syn):
    """No neighbours: zero halos are the conv's zero padding -> bit-identical to stgcn_model."""
    from oracle import stgcn_oracle as O, tsplit_oracle as TO
    cfg = syn.arch_config('st-gcn', num_classes=12, **SMALL)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 21)
    x = syn.synth_input((1, 3, 37, 25), 23)
    ocfg = dict(layers=4, stride=SMALL['stride'], residual=[1] * 4, normalization='LayerNorm')
    got = TO.stgcn_model_tsplit(x, sd, ocfg, lambda a, b: (None, None), lambda t: t, 37)
    ref = O.stgcn_model(x, sd, ocfg)
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 1e-6
