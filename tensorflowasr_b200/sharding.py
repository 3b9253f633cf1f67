"""Utterance sharding across the GPUs of one box: one process per GPU, full weight replica, contiguous
split of the batch, and ONE exchange per batch -- an all_gather of the decoded token ids + lengths (NCCL on GPUs, gloo
in the CPU tests).  No tensor / sequence parallelism: d_model is 144."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(num_utts: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the batch owned by `rank`; the first (num_utts % world) ranks get one extra."""
    base, extra = divmod(num_utts, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_ids(ids, lens, group=None):
    """ids [B_local, T] int32 (-1 padded), lens [B_local] -> (ids [B_total, T_max], lens [B_total]) on every rank.
    Ranks may hold different B_local and T (ragged shards are padded to the group maximum before the collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    shape = torch.tensor([ids.shape[0], ids.shape[1]], device=ids.device, dtype=torch.int64)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    bmax = max(int(s[0]) for s in shapes)
    tmax = max(int(s[1]) for s in shapes)
    pad = torch.full((bmax, tmax), -1, device=ids.device, dtype=torch.int32)
    pad[:ids.shape[0], :ids.shape[1]] = ids
    lpad = torch.zeros((bmax,), device=ids.device, dtype=torch.int32)
    lpad[:lens.shape[0]] = lens
    all_ids = [torch.empty_like(pad) for _ in range(world)]
    all_lens = [torch.empty_like(lpad) for _ in range(world)]
    dist.all_gather(all_ids, pad, group=group)
    dist.all_gather(all_lens, lpad, group=group)
    out_ids = torch.cat([a[:int(s[0])] for a, s in zip(all_ids, shapes)], dim=0)
    out_lens = torch.cat([l[:int(s[0])] for l, s in zip(all_lens, shapes)], dim=0)
    return out_ids, out_lens


def recognize_sharded(engine, wavs, group=None):
    """Every rank passes the SAME global batch `wavs` [B, L]; each encodes + decodes its slice on its own GPU and the ids
    are all-gathered.  Returns (ids [B, T'], lens [B]) on every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b0, b1 = shard_range(len(wavs), rank, world)
    ids, lens = engine.recognize(wavs[b0:b1])
    return gather_ids(ids, lens, group)
