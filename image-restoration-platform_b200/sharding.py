"""Batch sharding across GPUs: images are independent, so ranks split the batch
and nothing is reduced across them — no NCCL on the data path.  One process per GPU."""
from __future__ import annotations

import heapq
from typing import List, Sequence


def lpt_assign(costs: Sequence[float], n_ranks: int) -> List[List[int]]:
    """Longest-processing-time-first: sort by cost (pixels) descending, give each item to the least
    loaded rank.  Deterministic, so every rank derives the same partition without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + float(costs[i]), r))
    for s in shards:
        s.sort()
    return shards


def shard_groups(n_groups: int, n_ranks: int) -> List[List[int]]:
    """Fusion triplets stay on one GPU: contiguous round-robin of whole groups."""
    return [list(range(r, n_groups, n_ranks)) for r in range(n_ranks)]


def gather_results(local: list, indices: Sequence[int], total: int, group=None) -> list:
    """Host-side gather of per-image results (~1.2 KB each) to every rank, in batch order."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * total
        for i, r in zip(indices, local):
            out[i] = r
        return out
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (list(indices), local), group=group)
    out = [None] * total
    for idx, res in parts:
        for i, r in zip(idx, res):
            out[i] = r
    return out
