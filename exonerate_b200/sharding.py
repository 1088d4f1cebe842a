"""Multi-GPU sharding of a pair list.

Pairs are independent lattices (src/hub/gam.c:1140-1180), so the path shards
with NO data-path collective: every rank fills its own pairs.  The only exchange
is the gather of the fixed-size per-pair result records, after which results
are put back into the reference's output order (query-major, target-minor:
src/database/fastapipe.c:106-137) -- here simply the caller's pair order.
"""
import heapq

import numpy as np


def lpt_shards(costs, world):
    """Longest-processing-time-first deal of pairs to `world` ranks by lattice
    cost (query_length * target_length).  Returns a list of index arrays, each
    sorted ascending (so a rank keeps the caller's relative order)."""
    order = sorted(range(len(costs)), key=lambda k: (-int(costs[k]), k))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for k in order:
        load, r = heapq.heappop(heap)
        shards[r].append(k)
        heapq.heappush(heap, (load + int(costs[k]), r))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_records(local_records, shards, rank, world, group=None):
    """All-gather the per-pair records (int32 [n_local, width]) of every rank and
    scatter them back into global pair order.  `local_records` may be a CPU or a
    CUDA tensor (gloo / NCCL); returns a tensor [n_total, width] on the same device."""
    import torch
    import torch.distributed as dist

    width = local_records.shape[1]
    n_max = max(len(s) for s in shards)
    padded = torch.zeros((n_max, width), dtype=local_records.dtype, device=local_records.device)
    padded[: local_records.shape[0]] = local_records
    if world > 1:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    else:
        parts = [padded]
    total = sum(len(s) for s in shards)
    out = torch.zeros((total, width), dtype=local_records.dtype, device=local_records.device)
    for r in range(world):
        idx = torch.as_tensor(shards[r], device=local_records.device)
        out[idx] = parts[r][: len(shards[r])]
    return out
