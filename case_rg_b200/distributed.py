"""Multi-GPU story of the decode path: queries are independent, so each rank decodes its own slice
and there is exactly one exchange at the end.

Partitioning is the reference's own: ``DistributedSampler(dataset, shuffle=False)``
(common/CumulativeTrainer.py:139) - rank r of G takes indices r, r+G, ... and the index list is
padded by wrapping so that every rank gets the same count.  The reference "gathers" by having every
rank write its own result files which are concatenated later (Utils.py:38-49,
Run_Evaluation.py:45-71; duplicates from the padding collapse on the sample id); here one
``all_gather`` of int64 [n_rank, 1 + T] rows (id, tokens) does the same and rank order is undone.
Works on NCCL (GPU) and gloo (CPU tests).
"""
import math
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Indices DistributedSampler(shuffle=False, drop_last=False) gives to ``rank``."""
    if n == 0:
        return []
    per = math.ceil(n / world)
    total = per * world
    idx = list(range(n))
    pad = total - n
    if pad:
        idx += (idx * math.ceil(pad / n))[:pad]
    return idx[rank:total:world]


def gather_answers(ids: torch.Tensor, tokens: torch.Tensor, T: int, n_total: int,
                   group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks call this with their sample ids [n] and tokens [n, <=T]; every rank gets back
    (ids [n_total], tokens [n_total, T]) ordered by sample id with the wrap-around duplicates dropped."""
    dev = tokens.device
    n = tokens.size(0)
    row = torch.zeros(n, 1 + T, dtype=torch.int64, device=dev)
    row[:, 0] = ids.to(dev)
    row[:, 1:1 + tokens.size(1)] = tokens
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        parts = [torch.empty_like(row) for _ in range(world)]
        dist.all_gather(parts, row, group=group)
        row = torch.cat(parts, dim=0)
    out = torch.zeros(n_total, T, dtype=torch.int64, device=dev)
    seen = torch.zeros(n_total, dtype=torch.bool, device=dev)
    ids_all = row[:, 0]
    # later duplicates overwrite earlier ones with identical content (same query -> same answer)
    out[ids_all] = row[:, 1:]
    seen[ids_all] = True
    if not bool(seen.all()):
        raise RuntimeError('gather_answers: some sample ids were never produced')
    return torch.arange(n_total, device=dev), out
