"""T-split host logic: one long trial partitioned along time across ranks (BASELINE config 4).

Every rank runs the whole trunk on a contiguous chunk of frames.  All stages of an st_gcn layer are
frame-local except the Gamma x 1 temporal convolution, whose input needs (Gamma-1)/2 frames of the
neighbouring chunks; those boundary frames are exchanged once per layer with ``send``/``recv``
between ring neighbours (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests).
The global average pool becomes an all-reduce of per-rank channel sums.  The reference has no
counterpart: its long-sequence mechanism recomputes an overlapping halo on every replica
(utils/segment_generator.py:49-54, 91-106) instead of exchanging one.
"""
import ctypes

import torch

HALO_FRAMES = 4                      # kHalo in csrc/stgcn_api.cu
EXCHANGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t)


class HaloDesc(ctypes.Structure):
    """``stgcn_halo_desc`` (include/stgcn_b200.h)."""
    _fields_ = [('has_left', ctypes.c_int32), ('has_right', ctypes.c_int32),
                ('send_left', ctypes.c_void_p), ('send_right', ctypes.c_void_p),
                ('recv_left', ctypes.c_void_p), ('recv_right', ctypes.c_void_p),
                ('capacity', ctypes.c_size_t), ('exchange', EXCHANGE_FN), ('ctx', ctypes.c_void_p)]


def total_stride(strides):
    s = 1
    for x in strides:
        s *= int(x)
    return s


def chunk_bounds(num_frames, world, align):
    """Contiguous [start, stop) frame ranges, one per rank; every chunk but the last is a multiple
    of ``align`` frames (the trunk's total temporal stride), sizes as equal as that allows."""
    if world < 1 or num_frames < 1:
        raise ValueError("need at least one rank and one frame")
    units = -(-num_frames // align)                   # ceil: the ragged tail belongs to the last chunk
    if units < world:
        raise ValueError("%d frames cannot be split into %d chunks of multiples of %d frames"
                         % (num_frames, world, align))
    base, extra = divmod(units, world)
    bounds, start = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        stop = min(start + n, num_frames) if r < world - 1 else num_frames
        bounds.append((start, stop))
        start = stop
    return bounds


def validate_chunks(num_frames, world, strides):
    """Host-side precondition of the T-split forward, evaluated identically on EVERY rank from the
    globally known sizes (so a bad split raises everywhere instead of failing on one rank while the
    others block in the halo exchange): every chunk must keep at least ``HALO_FRAMES`` frames at every
    layer's resolution, because a rank can only send halo frames it owns."""
    bounds = chunk_bounds(num_frames, world, total_stride(strides))
    if world == 1:
        return bounds
    for r, (a, b) in enumerate(bounds):
        t = b - a
        for i, s in enumerate(strides):
            if t < HALO_FRAMES:
                raise ValueError("T-split: rank %d holds %d frame(s) at layer %d (< %d halo frames); use fewer "
                                 "ranks or a longer trial (%d frames over %d ranks)"
                                 % (r, t, i, HALO_FRAMES, num_frames, world))
            t = (t - 1) // int(s) + 1
    return bounds


def frames_after(t, strides):
    """Frames left after the trunk's temporal down-sampling (T_out = (T-1)//s + 1 per layer)."""
    for s in strides:
        t = (t - 1) // int(s) + 1
    return t


class DistExchange:
    """Ring-neighbour halo exchange over ``torch.distributed`` point-to-point ops.

    Owns the four staging buffers (uint8, ``capacity`` bytes each) on ``device``.  ``__call__`` is
    the host callback of ``stgcn_halo_desc``: it posts, in one batch, send_left -> rank-1,
    send_right -> rank+1 and the matching receives, and waits for them (on CUDA the wait is a
    stream dependency, not a host block)."""

    def __init__(self, rank, world, capacity, device, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.has_left, self.has_right = rank > 0, rank < world - 1
        self.capacity = int(capacity)
        mk = lambda: torch.zeros(self.capacity, dtype=torch.uint8, device=device)   # noqa: E731
        self.send_left, self.send_right, self.recv_left, self.recv_right = mk(), mk(), mk(), mk()
        self.calls, self.bytes_sent = 0, 0

    def peer(self, offset):
        import torch.distributed as dist
        r = self.rank + offset
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def __call__(self, layer, nbytes):
        import torch.distributed as dist
        ops = []
        if self.has_left:
            ops.append(dist.P2POp(dist.isend, self.send_left[:nbytes], self.peer(-1), self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_left[:nbytes], self.peer(-1), self.group))
        if self.has_right:
            ops.append(dist.P2POp(dist.isend, self.send_right[:nbytes], self.peer(+1), self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_right[:nbytes], self.peer(+1), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.calls += 1
        self.bytes_sent += nbytes * (int(self.has_left) + int(self.has_right))
        return 0

    def all_reduce_sum(self, t):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def descriptor(self):
        """(HaloDesc, keep-alive) for the C ABI."""
        def cb(_ctx, layer, nbytes):
            try:
                return int(self(layer, nbytes))
            except Exception as e:                       # never let an exception cross the C frame
                self.error = e
                return 1
        fn = EXCHANGE_FN(cb)
        d = HaloDesc()
        d.has_left, d.has_right = int(self.has_left), int(self.has_right)
        d.send_left, d.send_right = self.send_left.data_ptr(), self.send_right.data_ptr()
        d.recv_left, d.recv_right = self.recv_left.data_ptr(), self.recv_right.data_ptr()
        d.capacity = self.capacity
        d.exchange = fn
        d.ctx = None
        return d, fn
