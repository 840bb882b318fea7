"""Frame-range sharding of one clip across ranks (one process per GPU, torch.distributed for the plumbing).

The difference path shards naturally by frames: rank r of R owns the contiguous range
[r*N/R, (r+1)*N/R).  The only data exchanged is
  * overall mode:   raw frame 0 (the reference), broadcast from rank 0, each rank builds its reference plane itself;
  * per-frame mode: the one-frame halo -- rank r's last raw frame goes to rank r+1, which primes its state plane;
  * at the end:     one all-reduce (sum) of the u32 accumulator planes, viewed as int32 (two's-complement addition
                    is the same bits as unsigned addition, so the result is exact and order independent).
Per-frame scalars are disjoint by shard and stay on their rank.

`run_sharded` is engine-agnostic: the GPU engine (bench.py / tests -m gpu) wraps a dips_b200.Context, the CPU tests
(gloo, world_size 2) inject an engine built on numpy arrays.  The engine protocol:
    first_frame() -> tensor      raw bytes of the shard's first frame
    last_frame()  -> tensor      raw bytes of the shard's last frame
    frame_buffer() -> tensor     scratch tensor of one frame (receive buffer)
    prime(tensor) -> None        state plane := I2(frame)
    run(first_frame_index) -> None   process the local shard
    acc_tensor() -> tensor       int32 tensor whose element-wise sum over the ranks is the combined accumulators
optional:
    after_reduce() -> None       called after the all-reduce (unpacks an exchange format into the accumulators)
    state_tensor() -> tensor     the reference plane itself (as bytes); when present it is broadcast instead of raw frame 0
    mark_primed() -> None        the state plane was filled by the broadcast
    replicated (attribute) + local_reference(mode) -> tensor | None
                                 `replicated` is True when the shards carry their own copy of what they need before their
                                 first frame -- the clip's frame 0 (overall) or frame t0-1 (per-frame halo), loaded with the
                                 shard.  Then there is no exchange before the pass at all: every rank > 0 primes from
                                 local_reference(mode) and starts at once (rank 0 needs nothing).  The attribute must have
                                 the same value on every rank -- it decides whether the ranks meet in a collective.
"""
from __future__ import annotations

from typing import Protocol, Tuple

MODE_OVERALL, MODE_PERFRAME = 0, 1


def shard_range(rank: int, world: int, n_frames: int) -> Tuple[int, int]:
    """[t0, t1) owned by `rank`; ranges are contiguous, disjoint and cover [0, n_frames)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


class ShardEngine(Protocol):
    def first_frame(self): ...
    def last_frame(self): ...
    def frame_buffer(self): ...
    def prime(self, frame) -> None: ...
    def run(self, first_frame_index: int) -> None: ...
    def acc_tensor(self): ...


def exchange_reference(engine: ShardEngine, mode: int, rank: int, world: int, dist=None, group=None) -> None:
    """Give every rank what it needs before its first frame (no-op for a single rank)."""
    if world == 1:
        return
    if getattr(engine, "replicated", False):      # same on every rank by contract: nobody enters a collective here
        ref = engine.local_reference(mode)
        if ref is not None:
            engine.prime(ref)
        elif rank != 0:                            # rank 0 starts the clip: its pass primes itself from its first frame
            raise ValueError("replicated shard of rank %d has no reference / halo frame for mode %d" % (rank, mode))
        return
    if mode == MODE_OVERALL and hasattr(engine, "state_tensor"):
        # rank 0 builds the reference plane and broadcasts it (2 B/px instead of the raw frame's 3-4 B/px)
        if rank == 0:
            engine.prime(engine.first_frame())
        dist.broadcast(engine.state_tensor(), src=0, group=group)
        if rank > 0:
            engine.mark_primed()
    elif mode == MODE_OVERALL:
        buf = engine.frame_buffer()
        if rank == 0:
            buf.copy_(engine.first_frame())
        dist.broadcast(buf, src=0, group=group)
        engine.prime(buf)
    else:
        ops = []
        buf = engine.frame_buffer() if rank > 0 else None
        if rank + 1 < world:
            ops.append(dist.P2POp(dist.isend, engine.last_frame().contiguous(), rank + 1, group=group))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, buf, rank - 1, group=group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if rank > 0:
            engine.prime(buf)


def run_sharded(engine: ShardEngine, mode: int, first_frame_index: int, rank: int = 0, world: int = 1, dist=None,
                group=None) -> None:
    """One clip pass on this rank's shard: reference/halo exchange, local pass, accumulator all-reduce."""
    exchange_reference(engine, mode, rank, world, dist, group)
    engine.run(first_frame_index)
    if world > 1:
        dist.all_reduce(engine.acc_tensor(), op=dist.ReduceOp.SUM, group=group)
        if hasattr(engine, "after_reduce"):
            engine.after_reduce()


class _DeviceBuffer:
    """Zero-copy view of library-owned device memory for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n_items: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n_items,), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuShardEngine:
    """ShardEngine over a dips_b200.Context and a device-resident shard (a torch uint8 tensor [n, frame_bytes])."""

    def __init__(self, ctx, frames, torch, total_frames=None, replicated=False, reference=None):
        """total_frames: frames of the whole clip over all ranks -- enables the packed accumulator exchange
        (dipsb_pack_accumulators_device); None exchanges the two u32 planes as they are.
        replicated: the shards carry their own reference / halo frame (pass the SAME value on every rank); then
        reference = {mode: device tensor} holds this shard's copy of frame 0 (MODE_OVERALL) / frame t0-1 (MODE_PERFRAME)
        on every rank but 0.  replicated=False: broadcast / halo exchange per clip."""
        self.ctx, self.frames, self.torch = ctx, frames, torch
        # torch.distributed issues its collectives on torch's current stream; a Context launches on a private stream by
        # default.  Put the context on the stream that is current now, so that prime / clip kernel / pack precede the
        # collective and unpack / the next reset follow it.  Callers that switch streams later must call ctx.set_stream too.
        if getattr(frames, "is_cuda", False):
            ctx.set_stream(torch.cuda.current_stream(frames.device).cuda_stream)
        self.total_frames = total_frames
        self.replicated = bool(replicated)
        self.reference = reference
        self._buf = None
        self._views = {}

    def first_frame(self):
        return self.frames[0]

    def last_frame(self):
        return self.frames[-1]

    def frame_buffer(self):
        if self._buf is None:
            self._buf = self.torch.empty(self.frames.shape[1], dtype=self.torch.uint8, device=self.frames.device)
        return self._buf

    def prime(self, frame) -> None:
        self.ctx.prime_device(frame.data_ptr())

    def local_reference(self, mode):
        return None if not self.reference else self.reference.get(mode)

    def run(self, first_frame_index: int) -> None:
        self.ctx.run_clip_device(self.frames.data_ptr(), self.frames.shape[0], self.frames.stride(0), first_frame_index)

    def _view(self, ptr, n_items, typestr):
        key = (ptr, n_items, typestr)
        if key not in self._views:
            self._views[key] = self.torch.as_tensor(_DeviceBuffer(ptr, n_items, typestr), device=self.frames.device)
        return self._views[key]

    def acc_tensor(self):
        if self.total_frames is None:
            ptr, n = self.ctx.accumulators_device()
            return self._view(ptr, 2 * n, "<i4")
        ptr, n_words = self.ctx.pack_accumulators_device(self.total_frames)
        return self._view(ptr, n_words, "<i4")

    def after_reduce(self) -> None:
        if self.total_frames is not None:
            self.ctx.unpack_accumulators_device()

    def state_tensor(self):
        return self._view(self.ctx.state_plane_device(), 2 * self.ctx.npx, "|u1")   # u16 plane as bytes (any backend moves bytes)

    def mark_primed(self) -> None:
        self.ctx.mark_state_valid(True)
