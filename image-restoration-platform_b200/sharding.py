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


class PullQueue:
    """The mixed-resolution job queue of BASELINE.json configs[4] ("dynamic pull from a shared
    host queue"): every rank draws the next chunk index from ONE counter, so a rank that drew small images simply comes
    back sooner.  The counter is an atomic add in the rendezvous store torch.distributed already runs (host side, a TCP
    round trip per draw — no collective, nothing on the GPUs' data path); without a process group it is a local counter.
    One key per pass over the queue."""

    def __init__(self, n_items: int, key: str, store=None):
        self.n, self.key, self.store, self._local = n_items, key, store, 0

    @staticmethod
    def default_store():
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.distributed_c10d._get_default_store()
        return None

    def pull(self):
        """Next item index, or None when the queue is drained."""
        if self.store is None:
            i, self._local = self._local, self._local + 1
        else:
            i = self.store.add(self.key, 1) - 1
        return i if i < self.n else None
