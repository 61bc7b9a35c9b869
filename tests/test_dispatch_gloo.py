"""CPU, world_size 2 over gloo: the one-process-per-GPU dispatcher's host logic
(source assignment + max/sum combination of per-rank counters)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uoparallel_seismic_project_b200 import dispatch, workloads as W


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src = dispatch.sources_for_rank(rank, world)
    out = dispatch.combine(dist, torch.device("cpu"), elapsed_ms=10.0 * (rank + 1), relaxations=1000 + rank,
                           sources=len(src), launches=7)
    q.put((rank, src.tolist(), out))
    dist.destroy_process_group()


def test_weak_scaling_assignment_and_combine():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, o0), (r1, s1, o1) = res
    assert s0 == W.starts(4).tolist()                    # rank 0 = BASELINE config 2 exactly
    assert s1 == W.starts(111)[:4].tolist()              # rank 1 = next rows of start-111
    assert o0 == o1 == dict(elapsed_ms=20.0, relaxations=2001, sources=8, launches=14)  # max time, summed work


def test_sources_are_distinct_across_8_ranks_and_strong_split_covers_all():
    seen = set()
    for r in range(8):
        for p in dispatch.sources_for_rank(r, 8):
            seen.add(tuple(p))
    assert len(seen) == 32
    cover = sorted(i for r in range(8) for i in dispatch.shard_round_robin(111, r, 8))
    assert cover == list(range(111))
    assert max(len(dispatch.shard_round_robin(111, r, 8)) for r in range(8)) == 14  # ideal speed-up 111/14
