"""Multi-GPU plumbing around the search path: one process per GPU (`torch.distributed`), games sharded by uid, no collective
on the search path.  NCCL (or gloo in CPU tests) is used only where the path has a real exchange:
gathering the variable-length sample blocks of all ranks after self-play (the reference has one process and one buffer,
mcts_gpu.jl:515 -> main4IARow.jl:49-63), and reducing timings/counters for the benchmark.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def shard_games(total_games: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of game uids: returns (uid_base, count) of this rank.  Results per game do not depend on the
    partition because the RNG is keyed by uid (tests: shard invariance)."""
    base, rem = divmod(total_games, world)
    count = base + (1 if rank < rem else 0)
    uid_base = rank * base + min(rank, rem)
    return uid_base, count


def reduce_max_sum(max_vals, sum_vals, device=None):
    """MAX-reduce timings and SUM-reduce counters over all ranks (bench contract: time = max over ranks, work = sum)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(x) for x in max_vals], [float(x) for x in sum_vals]
    t = torch.tensor(list(max_vals), dtype=torch.float64, device=device)
    w = torch.tensor(list(sum_vals), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in w.tolist()]


_PINNED_CACHE: Dict[tuple, "object"] = {}


def _pinned_rows(key, rows, tail_shape, dtype):
    """A page-locked host tensor of at least `rows` rows, cached per (field, row shape, dtype)."""
    import torch
    ck = (key, tuple(tail_shape), dtype)
    t = _PINNED_CACHE.get(ck)
    if t is None or t.shape[0] < rows:
        t = torch.empty((max(rows, 1),) + tuple(tail_shape), dtype=dtype, pin_memory=True)
        _PINNED_CACHE[ck] = t
    return t


def gather_samples(samples: Dict[str, np.ndarray], device=None, dst: Optional[int] = None, reuse_buffers: bool = False) -> Optional[Dict[str, np.ndarray]]:
    """All ranks contribute their sample block (dict of arrays with equal leading length); every rank (dst=None) or only `dst`
    receives the concatenation in rank order — which, with block-partitioned uids, is ascending game uid within each rank block.
    Variable lengths: counts are all-gathered first, blocks are padded to the maximum for one all_gather per field.

    device (NCCL): one H2D copy, one all_gather_into_tensor, one on-device compaction and one D2H copy into a page-locked buffer per field.
    reuse_buffers=True returns views of those cached buffers (valid until the next call) instead of copies."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return samples
    world, rank = dist.get_world_size(), dist.get_rank()
    n = int(next(iter(samples.values())).shape[0])
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = n
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts_l = [int(c) for c in counts.tolist()]
    nmax, total = max(counts_l), sum(counts_l)
    out = {}
    for key, arr in samples.items():
        t = torch.from_numpy(np.ascontiguousarray(arr))
        tail = tuple(t.shape[1:])
        if device is not None:
            dev_in = torch.empty((nmax,) + tail, dtype=t.dtype, device=device)      # rows beyond n are never read back
            dev_in[:n].copy_(t, non_blocking=True)
            dev_out = torch.empty((world * nmax,) + tail, dtype=t.dtype, device=device)
            dist.all_gather_into_tensor(dev_out, dev_in)
            if dst is None or dst == rank:
                comp = torch.cat([dev_out[r * nmax:r * nmax + c] for r, c in enumerate(counts_l)], dim=0) if total else dev_out[:0]
                host = _pinned_rows(key, total, tail, t.dtype)[:total]
                host.copy_(comp)
                out[key] = host.numpy() if reuse_buffers else host.numpy().copy()
            continue
        pad = torch.zeros((nmax,) + tail, dtype=t.dtype)
        pad[:n] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        if dst is None or dst == rank:
            out[key] = np.concatenate([b[:c].numpy() for b, c in zip(bufs, counts_l)], axis=0)
    return out if (dst is None or dst == rank) else None
