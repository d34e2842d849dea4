"""Multi-GPU partitioning of a batch of independent slices.

Every slice is encoded/decoded with no reference to any other (reference core.py:172-186, 369-378), so a
batch is split into contiguous shards, one per rank, with NO collective on the data path.  The only exchange
is the optional gather of the per-file lengths into a global offsets index (a few KB).
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) of `total` units for `rank`; the first total % world ranks get one extra unit."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    q, r = divmod(total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_index(local_offsets, total: int, group=None):
    """All-gather the per-rank file lengths into the global offsets array int64[total + 1].

    local_offsets: int64[n_local + 1] offsets of this rank's arena (numpy or torch, any device).  Ranks may own
    different counts (shard_range).  Returns a numpy array on every rank; arena i of rank r starts at
    global_offsets[lo_r]."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lens = torch.as_tensor(np.diff(np.asarray(local_offsets.cpu() if hasattr(local_offsets, "cpu") else local_offsets)),
                           dtype=torch.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    q, r = divmod(total, world)
    pad = q + (1 if r else 0)
    buf = torch.zeros(pad, dtype=torch.int64, device=dev)
    buf[:lens.numel()] = lens.to(dev)
    out = [torch.zeros(pad, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for rk in range(world):
        lo, hi = shard_range(total, rk, world)
        parts.append(out[rk][:hi - lo].cpu().numpy())
    offsets = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(np.concatenate(parts), out=offsets[1:])
    return offsets
