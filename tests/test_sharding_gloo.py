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


# ---------------------------------------------------------------------------------------------------------------
# ShardedVADManager: the router built on this layout
# ---------------------------------------------------------------------------------------------------------------

class _FakeManager:
    """BatchedVADManager's surface without a GPU: a stream 'starts' when it has been pushed 3 messages."""

    def __init__(self, max_streams, device=None, **kw):
        from real_time_vad.core.batched_manager import StepOutput, StreamEvent   # noqa: F401
        self.max_streams, self.device, self.kw = max_streams, device, kw
        self.streams, self.pushed, self.fired = {}, {}, set()

    def open_stream(self, config=None, **cb):
        sid = len(self.streams)
        self.streams[sid] = cb
        self.pushed[sid] = 0
        return sid

    def close_stream(self, sid):
        self.streams.pop(sid, None)

    def push(self, sid, samples):
        self.pushed[sid] += 1

    def push_many(self, sids, block):
        for s in sids:
            self.pushed[int(s)] += 1

    def step(self):
        from real_time_vad.core.batched_manager import StepOutput, StreamEvent
        ids = sorted(s for s, k in self.pushed.items() if k > 0)
        ev = []
        for s in ids:
            if self.pushed[s] >= 3 and s not in self.fired:
                self.fired.add(s)
                ev.append(StreamEvent(s, "start", self.pushed[s] - 1, 0))
        return StepOutput(ev, np.array(ids, np.int64), np.ones(len(ids), np.int64), np.zeros((len(ids), 1), np.float32),
                          np.zeros((len(ids), 1), np.uint8))

    def close(self):
        pass


def _sharded_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from real_time_vad import ShardedVADManager, VADError
    mgr = ShardedVADManager(total, manager_factory=_FakeManager, frame_len=480)
    assert (mgr.rank, mgr.world_size) == (rank, world) and mgr.local.kw == {"frame_len": 480}
    assert mgr.local_capacity == sharding.local_capacity(total, world, rank)
    opened = [s for s in range(total) if mgr.open_stream(s)]           # every rank broadcasts every call ...
    assert opened == sharding.local_streams(total, world, rank).tolist() == mgr.open_streams   # ... the owner acts
    for _ in range(3):
        took = mgr.push_many(np.arange(total), np.zeros((total, 480), np.float32))
        assert took == len(opened)
    assert mgr.push(0, np.zeros(480, np.float32)) is (rank == 0)
    out = mgr.step()
    assert out.stream_ids.tolist() == opened                              # GLOBAL ids
    assert [e.stream_id for e in out.events] == opened
    everybody = mgr.all_events(out)
    try:
        mgr.open_stream(total)
        bad = False
    except VADError:
        bad = True
    assert mgr.close_stream(opened[0]) and not mgr.close_stream(opened[0])
    q.put((rank, everybody, bad))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_manager_routes_streams_to_their_owner_and_gathers_events():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                                       # the same global event list on every rank
    assert [e[0] for e in res[0][1]] == list(range(total)) and all(e[2] == "start" for e in res[0][1])
    assert res[0][2] and res[1][2]
