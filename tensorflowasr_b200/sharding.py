"""Utterance sharding across the GPUs of one box: one process per GPU, full weight replica, contiguous
split of the batch, and ONE exchange per batch -- a single all_gather of the decoded token ids + lengths (NCCL on GPUs, gloo
in the CPU tests).  No tensor / sequence parallelism: d_model is 144."""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_range(num_utts: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the batch owned by `rank`; the first (num_utts % world) ranks get one extra."""
    base, extra = divmod(num_utts, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class IdsExchange:
    """One collective per batch.  ids [B, T] and lens [B] of a rank live back to back in ONE flat int32 buffer, so the decoder
    writes them where the collective reads them (no packing kernels) and a single `all_gather_into_tensor` moves both.  The
    collective is issued asynchronously: NCCL's stream waits for the producing stream, the producing stream does NOT wait for NCCL,
    so the next batch's kernels launch immediately and the exchange overlaps them.  `slots` buffers rotate; a slot is waited for
    before it is handed out again.  Every rank must use the same (B, T)."""

    def __init__(self, B: int, T: int, device, group=None, slots: int = 2):
        import torch
        import torch.distributed as dist
        self.B, self.T, self.group = B, T, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = B * T + B
        self.flat = [torch.full((n,), -1, device=device, dtype=torch.int32) for _ in range(slots)]
        # (world == 1: `gathered` aliases the slot)
        self.gathered = [torch.empty((self.world * n,), device=device, dtype=torch.int32) if self.world > 1 else f for f in self.flat]
        self.gathered = list(self.gathered)
        self.work: List[Optional[object]] = [None] * slots
        self._next = 0

    def acquire(self) -> int:
        """Next slot; blocks the CURRENT STREAM (not the host) until that slot's previous collective has finished."""
        s = self._next
        self._next = (self._next + 1) % len(self.flat)
        if self.work[s] is not None:
            self.work[s].wait()
            self.work[s] = None
        return s

    def buffers(self, slot: int):
        """(ids [B, T], lens [B]) views of the slot's flat buffer: hand these to Engine.recognize."""
        n = self.B * self.T
        return self.flat[slot][:n].view(self.B, self.T), self.flat[slot][n:]

    def gather(self, slot: int):
        """Issue the slot's all_gather (async).  Call on the stream that produced the slot's contents."""
        import torch.distributed as dist
        if self.world == 1:
            return                                   # a single rank already holds everything: result() hands back the slot itself
        self.work[slot] = dist.all_gather_into_tensor(self.gathered[slot], self.flat[slot], group=self.group, async_op=True)

    def result(self, slot: int):
        """(ids [world * B, T], lens [world * B]) of the slot, after making the current stream wait for its collective."""
        import torch
        if self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None
        n = self.B * self.T
        if self.world == 1:
            return self.buffers(slot)
        g = self.gathered[slot].view(self.world, n + self.B)
        return g[:, :n].reshape(self.world * self.B, self.T), g[:, n:].reshape(self.world * self.B)


def gather_ids(ids, lens, group=None):
    """ids [B_local, T] int32 (-1 padded), lens [B_local] -> (ids [B_total, T_max], lens [B_total]) on every rank.
    Ranks may hold different B_local and T: one small all_gather of the shapes, then ONE all_gather of a packed buffer
    (length in column 0, ids behind it) padded to the group maximum."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    shape = torch.tensor([ids.shape[0], ids.shape[1]], device=ids.device, dtype=torch.int64)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    bmax = max(int(s[0]) for s in shapes)
    tmax = max(int(s[1]) for s in shapes)
    packed = torch.full((bmax, tmax + 1), -1, device=ids.device, dtype=torch.int32)
    packed[:, 0] = 0
    packed[:lens.shape[0], 0] = lens
    packed[:ids.shape[0], 1:ids.shape[1] + 1] = ids
    out = torch.empty((world * bmax, tmax + 1), device=ids.device, dtype=torch.int32)
    if hasattr(dist, "all_gather_into_tensor") and ids.device.type == "cuda":
        dist.all_gather_into_tensor(out, packed, group=group)
    else:                                            # gloo (CPU tests) has no all_gather_into_tensor
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed, group=group)
        out = torch.cat(parts, dim=0)
    out = out.view(world, bmax, tmax + 1)
    rows = [out[r, :int(s[0])] for r, s in enumerate(shapes)]
    allp = torch.cat(rows, dim=0)
    return allp[:, 1:].contiguous(), allp[:, 0].contiguous()


def recognize_sharded(engine, wavs, group=None):
    """Every rank passes the SAME global batch `wavs` [B, L]; each encodes + decodes its slice on its own GPU and the ids
    are all-gathered.  Returns (ids [B, T'], lens [B]) on every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b0, b1 = shard_range(len(wavs), rank, world)
    ids, lens = engine.recognize(wavs[b0:b1])
    return gather_ids(ids, lens, group)


class InFlight:
    """Several batches in flight on ONE GPU: `n` engine handles (same weights; own workspace and CUDA-graph cache each) on `n` streams,
    batch i goes to handle i % n.  Every kernel of the path is a single under-filled wave (126 CTAs of one per SM on 148 SMs at
    32 x 10 s) with a serial load -> MMA -> epilogue chain inside each CTA, so the kernels of batch i+1 fill the SMs and the launch
    gaps batch i leaves idle: measured +12.7 % frames/s at two in flight on configs[1] (profiles/r02_bench_c2_two_in_flight.json
    against ..._single_in_flight.json).  The latency of ONE batch does not improve (it grows slightly); this is a throughput mode.

    `make_engine()` builds one handle.  `fork()` makes the side streams wait for the caller's current stream (inputs written
    there), `join()` makes the caller's stream wait for all of them; between the two, `stream_of(i)` is the stream batch i runs on."""

    def __init__(self, make_engine, n: int = 2, device=None):
        import torch
        if n < 1:
            raise ValueError("InFlight: n must be >= 1")
        self.engines = [make_engine() for _ in range(n)]
        self.device = torch.device("cuda", self.engines[0].device) if device is None else device
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n)]

    def __len__(self):
        return len(self.engines)

    def stream_of(self, i: int):
        return self.streams[i % len(self.streams)]

    def engine_of(self, i: int):
        return self.engines[i % len(self.engines)]

    def fork(self):
        import torch
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def join(self):
        import torch
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def recognize(self, i: int, wav, ids=None, lens=None, frame_lengths=None):
        """Batch i: Engine.recognize on handle / stream i % n (no host sync).  Returns (ids, lens) on the GPU; they are complete once
        `stream_of(i)` has reached this point (join(), or an event recorded on that stream)."""
        import torch
        with torch.cuda.stream(self.stream_of(i)):
            return self.engine_of(i).recognize(wav, ids, lens, frame_lengths)

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.engines)
