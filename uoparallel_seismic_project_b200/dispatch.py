"""Multi-start dispatcher for one-process-per-GPU launches (torchrun).

Sources are independent shortest-path problems on a read-only slowness box
(mpi/backup.c:351-363 gives start point r to rank r and never communicates during the sweeps),
so the data path needs NO collective: every rank solves its own sources on its own replica of
the box.  torch.distributed is only used for the barrier around the timed region and to
combine the per-rank counters (sum of relaxations / sources, max of elapsed time).
"""
from __future__ import annotations

import numpy as np

from . import workloads as W

SOURCES_PER_RANK = 4


def sources_for_rank(rank: int, world: int, per_rank: int = SOURCES_PER_RANK) -> np.ndarray:
    """Weak scaling over BASELINE configs 2/3: rank 0 takes docs/start-4-241-241-51.txt (config 2
    exactly); rank r >= 1 takes the next `per_rank` rows of docs/start-111-241-241-51.txt."""
    if rank == 0 and per_rank == 4:
        return W.starts(4).copy()
    pool = W.starts(111)
    lo = ((rank - 1) * per_rank) % len(pool) if per_rank == 4 else (rank * per_rank) % len(pool)
    idx = [(lo + i) % len(pool) for i in range(per_rank)]
    return pool[idx].copy()


def shard_round_robin(num_sources: int, rank: int, world: int) -> list[int]:
    """Strong-scaling split of one start file (config 3: 111 sources over 1/2/4/8 GPUs)."""
    return list(range(rank, num_sources, world))


def combine(dist, device, *, elapsed_ms: float, relaxations: int, sources: int, launches: int):
    """max(elapsed) / sum(work) over ranks.  `dist` is torch.distributed (or None at N=1)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(elapsed_ms=elapsed_ms, relaxations=relaxations, sources=sources, launches=launches)
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    w = torch.tensor([relaxations, sources, launches], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return dict(elapsed_ms=float(t.item()), relaxations=int(w[0].item()), sources=int(w[1].item()),
                launches=int(w[2].item()))
