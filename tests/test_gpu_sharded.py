"""GPU: ShardedVADManager with two processes (gloo for the optional event gather; both ranks use the box's GPU -- two
engines on one device stand in for two GPUs, the layout is the same).  Every stream's probabilities and events must equal
what ONE BatchedVADManager holding all the streams produces: sharding is a partition, nothing more."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import synth_streams

pytestmark = pytest.mark.gpu
CFG = dict(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=2, voice_end_frame_count=3)
N, T = 37, 24


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from real_time_vad import ShardedVADManager, VADConfig
    audio = synth_streams(N, 512 * T, seed=404)
    fired = []
    mgr = ShardedVADManager(N)
    for s in range(N):
        mgr.open_stream(s, VADConfig(**CFG), on_voice_start=lambda s=s: fired.append(s))
    probs = {s: [] for s in mgr.open_streams}
    events = []
    for j in range(T):
        mgr.push_many(np.arange(N), audio[:, j * 512:(j + 1) * 512])
        out = mgr.step()
        for k, s in enumerate(out.stream_ids):
            probs[int(s)].append(float(out.probs[k, 0]))
        events += mgr.all_events(out)
    q.put((rank, mgr.open_streams, {s: np.array(p, np.float32) for s, p in probs.items()}, events, sorted(set(fired))))
    dist.barrier()
    mgr.close()
    dist.destroy_process_group()


def test_two_ranks_equal_one_manager():
    from real_time_vad import BatchedVADManager, VADConfig
    audio = synth_streams(N, 512 * T, seed=404)
    one = BatchedVADManager(max_streams=N)
    ids = [one.open_stream(VADConfig(**CFG)) for _ in range(N)]
    want_p = {s: [] for s in range(N)}
    want_ev = []
    for j in range(T):
        one.push_many(ids, audio[:, j * 512:(j + 1) * 512])
        out = one.step()
        for k, sid in enumerate(out.stream_ids):
            want_p[ids.index(int(sid))].append(float(out.probs[k, 0]))
        want_ev += [(ids.index(e.stream_id), e.frame_index, e.kind) for e in out.events]
    one.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] == list(range(0, N, 2)) and res[1][1] == list(range(1, N, 2))     # the partition
    for _, owned, probs, _, fired in res:
        for s in owned:
            assert np.array_equal(probs[s], np.array(want_p[s], np.float32)), s        # same kernels, same bits
        assert set(fired) <= set(owned)                                                # callbacks fire on the owner only
    assert sorted(res[0][3]) == sorted(res[1][3]) == sorted(want_ev) and len(want_ev) >= 10
