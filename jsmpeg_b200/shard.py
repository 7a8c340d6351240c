"""Multi-GPU sharding of independent streams.

The path has no exchange step: stream i lives on rank i % world, every rank decodes its own shard
with its own BatchDecoder, and the only cross-rank operation is the bookkeeping reduction of the
counters (frames decoded, elapsed time) -- a tiny all_reduce, not a data-path collective.
"""
from __future__ import annotations


def assign_streams(n_streams, rank, world):
    """Indices of the streams rank `rank` owns (round-robin)."""
    return list(range(rank, n_streams, world))


def aggregate(frames, seconds, device=None):
    """Whole-job (sum of frames, max of seconds) over all ranks of the default process group.
    Works with gloo (CPU tensors) and nccl (pass device='cuda')."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames, seconds
    f = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(f.item()), float(t.item())
