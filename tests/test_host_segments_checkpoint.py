"""Host-side evaluation plumbing on CPU: trial segmentation / stitching against vectors
produced by the reference's ``BufferSegment`` / ``WindowSegment`` (tools/make_golden.py gen_segments), and
checkpoint ingest of a file written in the reference's format by the reference model (gen_checkpoint)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, load_golden

BASE = dict(stages=1, num_classes=3, graph=dict(num_node=5), in_feat=2)


def _ramp(L, ps, pe):
    cap = torch.arange(L, dtype=torch.float32).view(1, 1, L, 1).expand(1, 2, L, 5).contiguous()
    return F.pad(cap, (0, 0, ps, pe), value=-1.0)


def test_buffer_segment_matches_reference(pkg):
    z = np.load(os.path.join(GOLDEN, 'segments.npz'))
    for key in [k for k in z.files if k.startswith('buf|') and k.endswith('pad')]:
        _, L, S, G, W, _ = key.split('|')
        L, S, G, W = int(L), int(S), int(G), int(W)
        tag = 'buf|%d|%d|%d|%d|' % (L, S, G, W)
        seg = pkg.segment.BufferSegment(rank='cpu', world_size=W, kernel=G, segment=S, **BASE)
        ps, pe = seg.pad_sequence(L)
        assert [ps, pe, seg.num_segments()] == z[tag + 'pad'].tolist()
        cap, labels = _ramp(L, ps, pe), torch.arange(L).view(1, L)
        starts, lo, hi, kept = [], [], [], []
        for i, (x, y, n) in enumerate(seg.get_segment(cap, labels)):
            starts.append(int(x[0, 0, 0, 0]))
            lo.append(int(y[0, 0]) if y.numel() else -1)
            hi.append(int(y[0, -1]) if y.numel() else -1)
            kept.append(seg.mask_segment(i, n, L, ps, pe, x[:, :1, :, 0].expand(1, 3, S))[0, 0])
        assert starts == z[tag + 'starts'].tolist()
        assert [lo, hi] == z[tag + 'labels'].tolist()
        ref_kept = z[tag + 'kept']
        got = torch.cat(kept).numpy()
        if pe > 0:
            assert np.array_equal(got, ref_kept)
        else:
            # the reference drops the whole last chunk when P_end == 0 ([G-1:-0] is empty); the tail is kept here
            assert np.array_equal(got[:len(ref_kept)], ref_kept) and got[-1] == L - 1
        # every frame of the trial is covered, in order (the reference keeps one duplicate frame per boundary)
        assert np.all(np.diff(got[got >= 0]) >= 0) and set(got[got >= 0].tolist()) == set(range(L))


def test_buffer_segment_fold_mode_matches_reference(pkg):
    z = np.load(os.path.join(GOLDEN, 'segments.npz'))
    for key in [k for k in z.files if k.startswith('fold|') and k.endswith('pad')]:
        _, L, G, W, _ = key.split('|')
        L, G, W = int(L), int(G), int(W)
        tag = 'fold|%d|%d|%d|' % (L, G, W)
        seg = pkg.segment.BufferSegment(rank='cpu', world_size=W, kernel=G, **BASE)
        ps, pe = seg.pad_sequence(L)
        assert [ps, pe, seg.S] == z[tag + 'pad'].tolist()
        batches = list(seg.get_segment(_ramp(L, ps, pe), torch.arange(L).view(1, L)))
        assert len(batches) == 1                       # (the reference's generator returns instead of yielding)
        x, _, n = batches[0]
        assert x[:, 0, 0, 0].tolist() == z[tag + 'starts'].tolist()
        pred = (x[:, :1, :, 0] + 1.0).expand(W, 3, seg.S).contiguous()
        assert np.array_equal(seg.mask_segment(0, n, L, ps, pe, pred).numpy(), z[tag + 'out'])


def test_window_segment_matches_reference(pkg):
    z = np.load(os.path.join(GOLDEN, 'segments.npz'))
    for key in [k for k in z.files if k.startswith('win|') and k.endswith('pad')]:
        _, L, RF, S, _ = key.split('|')
        L, RF, S = int(L), int(RF), int(S)
        tag = 'win|%d|%d|%d|' % (L, RF, S)
        seg = pkg.segment.WindowSegment(rank='cpu', world_size=1, receptive_field=RF, segment=S, **BASE)
        ps, pe = seg.pad_sequence(L)
        first, last, nwin, lab = [], [], [], []
        for x, y, n in seg.get_segment(_ramp(L, ps, pe), torch.arange(L).view(1, L)):
            first.append(int(x[0, 0, -1, 0])); last.append(int(x[-1, 0, -1, 0])); nwin.append(x.shape[0])
            lab.append([int(y[0, 0]), int(y[0, -1])])
        assert [ps, pe, n] == z[tag + 'pad'].tolist()
        assert [first, last, nwin] == z[tag + 'ends'].tolist()
        assert lab == z[tag + 'labels'].tolist()


def test_forward_buffered_stitches_a_causal_model(pkg):
    """A causal per-frame 'model' whose output at t depends on the last G frames: stitched chunked evaluation
    equals the whole-trial evaluation (except the first G-1 frames of nothing -- chunk 0 starts at the trial start)."""
    G, L = 9, 173
    x = torch.randn(1, 2, L, 5)

    def model(c):                                       # (n, C, S, V) -> (n, classes, S): causal window sum
        s = c.mean(dim=(1, 3))
        out = torch.stack([F.pad(s, (G - 1, 0)).unfold(1, G, 1).sum(-1) * k for k in (1.0, 2.0, 3.0)], dim=1)
        return out
    seg = pkg.segment.BufferSegment(rank='cpu', world_size=1, kernel=G, segment=40, **BASE)
    got = pkg.segment.forward_buffered(model, x, seg)
    ref = model(x)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-5)


def test_checkpoint_ingest_reference_format(pkg, syn):
    """The file the reference's Processor._save_model writes (bare and DataParallel-prefixed) loads into the drop-in
    model key by key; wrong shapes are rejected; save_checkpoint round-trips."""
    _, w = load_golden('stgcn_model_small_ln')
    cfg = syn.arch_config('st-gcn', normalization='LayerNorm', num_classes=12, in_ch=[16, 16, 32], out_ch=[16, 32, 32],
                          stride=[1, 2, 1])
    for name in ('ckpt_stgcn_small.pt', 'ckpt_stgcn_small_dp.pt'):
        m = pkg.Stgcn(**cfg)
        meta = pkg.checkpoint.load_checkpoint(m, os.path.join(GOLDEN, name))
        assert meta['epoch'] == 7 and abs(meta['loss'] - 0.25) < 1e-12
        sd = m.state_dict()
        assert set(sd) == set(w)
        assert all(torch.equal(sd[k], w[k]) for k in w)
    wrapped = torch.nn.DataParallel(pkg.Stgcn(**cfg))
    pkg.checkpoint.load_checkpoint(wrapped, os.path.join(GOLDEN, 'ckpt_stgcn_small.pt'))
    assert torch.equal(wrapped.module.fcn_out.weight, w['fcn_out.weight'])
    other = pkg.Stgcn(**syn.arch_config('st-gcn', normalization='LayerNorm', num_classes=13, in_ch=[16, 16, 32],
                                        out_ch=[16, 32, 32], stride=[1, 2, 1]))
    with pytest.raises(ValueError, match='shape mismatch'):
        pkg.checkpoint.load_checkpoint(other, os.path.join(GOLDEN, 'ckpt_stgcn_small.pt'))


def test_checkpoint_roundtrip(pkg, syn, tmp_path):
    cfg = syn.arch_config('rt-st-gcn', num_classes=12, in_ch=[16, 16], out_ch=[16, 32], stride=[1, 2])
    a, b = pkg.RtStgcn(**cfg), pkg.RtStgcn(**cfg)
    a.load_state_dict(syn.synth_state_dict(a.state_dict(), 5))
    path = str(tmp_path / 'final.pt')
    pkg.checkpoint.save_checkpoint(a, path, epoch=3, loss=1.5)
    meta = pkg.checkpoint.load_checkpoint(b, path)
    assert meta == {'epoch': 3, 'loss': 1.5}
    assert all(torch.equal(a.state_dict()[k], b.state_dict()[k]) for k in a.state_dict())
