"""Trial segmentation for evaluation: the B200 counterparts of the reference's ``BufferSegment`` and
``WindowSegment`` (utils/segment_generator.py:6-154) plus the stitched evaluation loop that
``Processor._forward`` builds from them (processor.py:346-392).

``BufferSegment`` cuts a long trial into chunks that overlap by the temporal kernel ``G`` (the overlap
re-creates the FIFO contents a continual model would hold at the chunk boundary), runs the chunks as a
batch and stitches the per-frame predictions back together; ``WindowSegment`` turns every frame into a
window of ``receptive_field`` frames ending at it.  The arithmetic (padding, chunk count, which
predictions are dropped) follows the reference line by line because it defines which numbers the
accuracy tables were computed from; it is pinned to the reference's classes by ``tests/golden/segments.npz``.

Two deliberate differences, both where the reference is broken at HEAD:
  * ``BufferSegment.get_segment`` without a ``segment`` size is a generator that ``return``s its tuple
    (so iterating it yields nothing, segment_generator.py:73-77); here it yields the one batch.
  * ``mask_segment`` of the last chunk slices ``[G-1:-P_end]``, which is empty when ``P_end == 0``
    (segment_generator.py:91); here ``P_end == 0`` keeps the tail.
"""
import torch
import torch.nn.functional as F


class Segment:
    def __init__(self, rank=None, world_size=1, **kwargs):
        self.num_stages = kwargs.get('stages', 1)
        self.num_classes = kwargs['num_classes']
        self.V = kwargs['graph']['num_node']
        self.C = kwargs['in_feat']
        self.rank, self.world_size = rank, world_size

    def alloc_output(self, L, dtype):
        return torch.zeros(self.num_stages, self.num_classes, L, dtype=dtype, device=self.rank)


class BufferSegment(Segment):
    """Overlapped chunks of ``segment`` frames (stride ``segment - kernel``), ``world_size`` chunks per
    batch; without ``segment``: one chunk per executor."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.G = kwargs['kernel']
        self.subsegment_size = kwargs.get('segment')

    # -- padding (segment_generator.py:25-56) -------------------------------------------------
    def pad_sequence(self, L):
        self.L = L
        W, G, S = self.world_size, self.G, self.subsegment_size
        if S:
            hop = S - G
            tail = (L - S) % hop                               # frames past the last full chunk
            spare = (((L - G - tail) // hop) + 1) % W          # chunks past the last full batch
            self.P_end = (hop - tail if tail else 0) + (hop * (W - spare) if spare else 0)
        else:
            rem = (L - (W - 1) * (G - 1)) % W
            self.P_end = (W - rem) if rem else 0
            self.S = ((L + self.P_end - (W - 1) * (G - 1)) // W) + (0 if W == 1 else G - 1)
        return 0, self.P_end

    def pad_sequence_rt(self, L):
        self.L = L
        return 0, 0

    # -- chunking (segment_generator.py:62-81) ------------------------------------------------
    def num_segments(self):
        S, G = self.subsegment_size, self.G
        return ((self.L + self.P_end - S) // (S - G)) + 1 if S else 1

    def _chunks(self, captures, size, hop):
        # (1, C, L', V) -> (n, C, size, V): chunk i = frames [i*hop, i*hop + size)
        n = (captures.size(2) - size) // hop + 1
        return captures.unfold(2, size, hop).permute(0, 2, 1, 4, 3).contiguous().view(n, self.C, size, self.V)

    def get_segment(self, captures, labels):
        S, G, W = self.subsegment_size, self.G, self.world_size
        if not S:
            yield self._chunks(captures, self.S, self.S - G)[:W], labels, 1
            return
        n, hop = self.num_segments(), S - G
        data = self._chunks(captures, S, hop)[:n]
        for i in range(0, n, W):
            lo = 0 if i == 0 else S + hop * (i - 1)
            hi = S + hop * (i + W) if i + W < n - 1 else self.L
            yield data[i:i + W], labels[:, lo:hi], n

    def get_segment_rt(self, captures):
        for i in range(self.L):
            yield captures[:, :, i:i + 1]

    # -- stitching (segment_generator.py:83-106) ----------------------------------------------
    def mask_segment(self, i, num_segments, L, P_start, P_end, predictions):
        G = self.G
        if self.subsegment_size:
            if i == 0:
                return predictions
            if i < num_segments - 1:
                return predictions[:, :, G - 1:]
            return predictions[:, :, G - 1:predictions.size(2) - P_end]
        W, S = self.world_size, self.S
        predictions = predictions.clone()
        predictions[1:, :, :G] = 0                              # the overlap belongs to the previous chunk
        cols = predictions[None].permute(0, 2, 3, 1).contiguous().view(1, self.num_classes * S, W)
        full = F.fold(cols, output_size=(1, L + P_end), kernel_size=(1, S), stride=(1, S - G))[:, :, 0]
        return full[:, :, :L]


class WindowSegment(Segment):
    """Every frame becomes a window of ``receptive_field`` frames ending at it (zeros before the start)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.W = kwargs['receptive_field']
        self.subsegment_size = kwargs.get('segment')

    def pad_sequence(self, L):
        self.S = L
        return self.W - 1, 0

    def pad_sequence_rt(self, L):
        self.L = L
        return self.W - 1, 0

    def get_segment(self, captures, labels):
        S, W = self.subsegment_size, self.W
        n = (self.S + self.S % S) // S
        for i in range(n):
            x0 = S * i - (1 if i > 0 else 0)
            x1 = S * (i + 1) + (W - 1) if i < n - 1 else captures.size(2)
            y0, y1 = S * i, (S * (i + 1) if i < n - 1 else labels.size(1))
            win = captures[:, :, x0:x1].unfold(2, W, 1).permute(0, 2, 1, 4, 3).contiguous()
            yield win.view(x1 - x0 - (W - 1), self.C, W, self.V), labels[:, y0:y1], n

    def get_segment_rt(self, captures):
        for i in range(self.L):
            yield captures[:, :, i:i + self.W]

    def mask_segment(self, L, P_start, P_end, predictions):
        return predictions.permute(2, 1, 0)                      # (N', C', 1) -> (1, C', N')


@torch.no_grad()
def forward_buffered(model, captures, segment):
    """Overlapped-chunk evaluation of one trial, the ``BufferSegment`` branch of ``Processor._forward``
    (processor.py:346-392): pad, cut into chunks, run each batch of chunks through ``model``
    (``(n, C, S, V) -> (n, classes, S)``), drop the re-computed overlap and stitch.
    ``captures (1, C, L, V)`` -> predictions ``(1, classes, L)``.

    Chunks start every ``S - G`` frames, i.e. overlap by ``G`` frames; chunk ``i > 0`` contributes its frames
    from ``G`` on, which is what the reference's LABEL slices assume (segment_generator.py:70-72).  Its
    ``mask_segment`` keeps one frame more (``[G-1:]``), so its predictions and labels disagree by one frame per
    chunk boundary; ``BufferSegment.mask_segment`` above stays pinned to the reference, the stitching here
    uses the consistent cut."""
    L = captures.size(2)
    p_start, p_end = segment.pad_sequence(L)
    padded = F.pad(captures, (0, 0, p_start, p_end))
    labels = torch.zeros(1, L, dtype=torch.long, device=captures.device)
    if not segment.subsegment_size:
        chunk, _, n = next(iter(segment.get_segment(padded, labels)))
        return segment.mask_segment(0, n, L, p_start, p_end, model(chunk.contiguous()))
    pieces = []
    for i, (chunk, _, n) in enumerate(segment.get_segment(padded, labels)):
        out = model(chunk.contiguous())
        for j in range(out.size(0)):
            first = i * segment.world_size + j == 0
            pieces.append(out[j:j + 1] if first else out[j:j + 1, :, segment.G:])
    return torch.cat(pieces, dim=2)[:, :, :L]
