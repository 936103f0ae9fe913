"""CPU, world_size 2 over gloo: the N>1 layout (streams shard across ranks, no data-path
collective; only timing / event gathering crosses ranks)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from real_time_vad.engine import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert sharding.world() == (rank, world, rank)
    mine = sharding.local_streams(n_streams, world, rank)
    assert len(mine) == sharding.local_capacity(n_streams, world, rank)
    for slot, s in enumerate(mine):
        assert sharding.owner_of(int(s), world) == (rank, slot)
    # every rank "processes" its shard; the bench's reduction = max time, sum of units
    elapsed = 1.0 + rank
    t = sharding.reduce_max(elapsed)
    total = sharding.reduce_sum(float(len(mine)))
    local_events = [(slot, 10 * rank + slot, 1) for slot in range(min(3, len(mine)))]
    ev = sharding.gather_events(local_events, world, rank)
    q.put((rank, t, total, ev, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_streams_shard_across_two_ranks_without_overlap():
    world, n_streams = 2, 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_streams = sorted(s for r in res for s in r[4])
    assert all_streams == list(range(n_streams))                      # exact partition
    assert all(r[1] == 2.0 and r[2] == float(n_streams) for r in res)  # max over ranks, sum over ranks
    assert res[0][3] == res[1][3] == sorted(res[0][3])                 # same global, ordered event list everywhere
    assert {e[0] for e in res[0][3]} == {0, 2, 4, 1, 3, 5}


def test_single_process_helpers_need_no_process_group():
    assert sharding.reduce_max(3.5) == 3.5 and sharding.reduce_sum(2.0) == 2.0
    assert sharding.gather_events([(1, 2, 1)], 4, 3) == [(7, 2, 1)]
    assert sharding.local_streams(10, 4, 3).tolist() == [3, 7]
