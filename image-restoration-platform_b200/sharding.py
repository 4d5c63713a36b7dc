"""Batch sharding across GPUs: images are independent, so ranks split the batch
and nothing is reduced across them — no NCCL on the data path.  One process per GPU."""
from __future__ import annotations

import fcntl
import heapq
import os
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
    back sooner.  On one box the counter is eight bytes of a file in /dev/shm, incremented under flock (a few
    microseconds; host side, nothing on the GPUs' data path).  The rendezvous store's atomic add is the fallback for
    ranks that do not share a /dev/shm — measured at 8 ranks it costs milliseconds per draw (small TCP writes), which
    starved ranks outright.  Without a process group it is a local counter.  One key per pass over the queue."""

    def __init__(self, n_items: int, key: str, store=None, shm_dir: str = "/dev/shm"):
        self.n, self.key, self.store, self._local, self._fd = n_items, key, store, 0, None
        if store is not None and shm_dir and os.path.isdir(shm_dir) and int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0) == _world():
            run = os.environ.get("TORCHELASTIC_RUN_ID", "run") + "_" + os.environ.get("MASTER_PORT", "0")
            self._path = os.path.join(shm_dir, f"irp_queue_{run}_{key}")
            self._fd = os.open(self._path, os.O_CREAT | os.O_RDWR, 0o600)

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
        elif self._fd is not None:
            fcntl.flock(self._fd, fcntl.LOCK_EX)
            try:
                raw = os.pread(self._fd, 8, 0)
                i = int.from_bytes(raw, "little") if len(raw) == 8 else 0
                os.pwrite(self._fd, (i + 1).to_bytes(8, "little"), 0)
            finally:
                fcntl.flock(self._fd, fcntl.LOCK_UN)
        else:
            i = self.store.add(self.key, 1) - 1
        return i if i < self.n else None

    def close(self, unlink: bool = False):
        """Every rank closes; one of them (after a barrier) unlinks."""
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None
            if unlink:
                try:
                    os.unlink(self._path)
                except FileNotFoundError:
                    pass


def cleanup_queue_files(shm_dir: str = "/dev/shm") -> None:
    """Remove this run's counter files (call on one rank, after a barrier)."""
    import glob

    run = os.environ.get("TORCHELASTIC_RUN_ID", "run") + "_" + os.environ.get("MASTER_PORT", "0")
    for f in glob.glob(os.path.join(shm_dir, f"irp_queue_{run}_*")):
        try:
            os.unlink(f)
        except OSError:
            pass


def _world() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
