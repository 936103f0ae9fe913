"""The websocket service seam (real_time_vad/service/batched_server.py): protocol pieces on CPU, and on the GPU the
reference's end-to-end known answer -- SampleVoiceMono.wav as 30 ms int16 messages with the server's defaults gives
VOICE_START / VOICE_END x 4 with segment_index 0..3 (websocket_service/README.md:290,
examples/test_python_vad_client.py:201-204) -- through ClientSession slots of one shared manager, many clients at once,
with the same event frames the reference's own Python produced for that wire format (tests/golden mode A)."""
import asyncio
import json

import numpy as np
import pytest

from conftest import GOLDEN


def test_query_and_config_message_precedence():
    from real_time_vad.service.batched_server import create_client_config, parse_query_params
    q = parse_query_params("sample_rate=16000&start_probability=0.5&frame_duration_ms=20&timeout=3&mode=pcm")
    assert q == {"sample_rate": 16000, "start_probability": 0.5, "frame_duration_ms": 20, "timeout": 3.0, "mode": "pcm"}
    c = create_client_config({})
    assert (c.vad.start_probability, c.vad.end_probability, c.vad.start_frame_count, c.vad.end_frame_count) == (0.4, 0.3, 6, 12)
    assert (c.audio.sample_rate, c.audio.sample_width, c.audio.frame_duration_ms, c.timeout) == (16000, 2, 30, 0.0)
    c = create_client_config(q, {"end_frame_count": 20, "start_probability": 0.6, "mode": None})
    assert c.vad.end_frame_count == 20 and c.vad.start_probability == 0.6          # the CONFIG message wins
    assert c.audio.frame_duration_ms == 20 and c.timeout == 3.0


def test_app_has_the_reference_routes():
    from real_time_vad.service import BatchedVADService, create_app
    app = create_app(BatchedVADService(max_clients=4))
    paths = {r.path for r in app.routes}
    assert {"/vad", "/", "/health", "/clients"} <= paths


@pytest.mark.gpu
def test_sessions_reproduce_the_four_segment_answer_for_many_clients():
    from real_time_vad.service import BatchedVADService, create_client_config
    g = np.load(GOLDEN / "sample_voice.npz")
    q = g["q16k"]
    n_msg = len(q) // 480
    want = [tuple(e) for e in g["A_events"].tolist()]               # (frames processed when it fired, kind)
    svc = BatchedVADService(max_clients=64)
    outs, sessions = [], []

    async def main():
        for c in range(20):
            box = []
            outs.append(box)

            async def send(text, box=box):
                box.append(json.loads(text))
            sessions.append(svc.connect(send, create_client_config({})))
        bad = sessions[0]
        await bad.process_audio_frame(b"\x00" * 100)                    # wrong size: ERROR, nothing buffered
        assert outs[0][-1]["event"] == "ERROR" and "Invalid frame size: expected 960, got 100" in outs[0][-1]["message"]
        outs[0].clear()
        fired = [[] for _ in sessions]
        for m in range(n_msg):
            msg = q[m * 480:(m + 1) * 480].tobytes()
            for k, s in enumerate(sessions):
                if k % 2 == 0 or (m % 2 == 0 and m == n_msg - 1):
                    await s.process_audio_frame(msg)
                elif m % 2 == 1:                                         # odd clients deliver two messages every other tick
                    await s.process_audio_frame(q[(m - 1) * 480:m * 480].tobytes())
                    await s.process_audio_frame(msg)
            before = [len(o) for o in outs]
            await svc.tick()
            for k, o in enumerate(outs):
                fired[k] += [(m, e["event"]) for e in o[before[k]:] if e["event"] in ("VOICE_START", "VOICE_END")]
        return fired

    fired = asyncio.run(main())
    for k, (s, o) in enumerate(zip(sessions, outs)):
        kinds = [e["event"] for e in o if e["event"] != "VOICE_CONTINUE"]
        assert kinds == ["VOICE_START", "VOICE_END"] * 4, (k, kinds)
        assert [e["segment_index"] for e in o if e["event"] == "VOICE_START"] == [0, 1, 2, 3]
        assert [e["segment_index"] for e in o if e["event"] == "VOICE_END"] == [0, 1, 2, 3]
        assert all(e["duration_ms"] == e["segment_end_ms"] - e["segment_start_ms"] >= 0 for e in o if e["event"] == "VOICE_END")
        assert s.segment_index == 4
        n_cont = sum(e["event"] == "VOICE_CONTINUE" for e in o)
        assert n_cont == sum(e - st for (st, _), (e, _) in zip(want[0::2], want[1::2])), k   # one per frame while voice is active
    # even clients step one frame per tick: their events fall on the reference's frames (golden mode A)
    assert [(m, 1 if ev == "VOICE_START" else 2) for m, ev in fired[0]] == want
    for s in sessions:
        s.cleanup()
    assert not svc.sessions
    svc.close()


class _StubManager:
    """Stands in for BatchedVADManager on a box without a GPU: counts streams and pending samples."""
    made = 0

    def __init__(self, **kw):
        type(self).made += 1
        self.kw, self.streams, self.buffered, self.closed = kw, {}, {}, False

    def open_stream(self, cfg, **cb):
        sid = len(self.streams)
        self.streams[sid] = cfg
        self.buffered[sid] = 0
        return sid

    def close_stream(self, sid):
        self.streams.pop(sid, None)

    @property
    def open_streams(self):
        return sorted(self.streams)

    def pending(self, sid):
        return self.buffered[sid]

    def push_bytes(self, sid, data):
        self.buffered[sid] += len(data) // 2

    def step(self):
        raise RuntimeError("CUDA error: simulated")

    def close(self):
        self.closed = True


def test_remote_clients_cannot_grow_the_service_without_bound(monkeypatch):
    """A wire format costs a max_clients-slot engine and a pinned arena, and clients choose the format: at most
    `max_formats` managers exist at once, a manager is freed when its last stream closes, and a stream that runs more
    than `max_pending_frames` messages ahead of the tick has its audio dropped (with an ERROR) instead of buffered."""
    from real_time_vad.service import batched_server as bs
    monkeypatch.setattr(bs, "BatchedVADManager", _StubManager)
    svc = bs.BatchedVADService(max_clients=8, max_formats=2, max_pending_frames=3)
    sent = []

    async def send(text):
        sent.append(json.loads(text))

    async def main():
        a = svc.connect(send, bs.create_client_config({"frame_duration_ms": 30}))
        b = svc.connect(send, bs.create_client_config({"frame_duration_ms": 20}))
        with pytest.raises(ValueError, match="too many distinct audio formats"):
            svc.connect(send, bs.create_client_config({"frame_duration_ms": 40}))
        assert len(svc._managers) == 2
        mgr_b = b._manager
        b.cleanup()                                               # last stream of its format: the manager goes away
        assert len(svc._managers) == 1 and mgr_b.closed
        c = svc.connect(send, bs.create_client_config({"frame_duration_ms": 40}))   # now there is room again
        msg = b"\x00" * a.expected_frame_bytes
        for _ in range(6):
            await a.process_audio_frame(msg)
        assert a.frame_count == 3 and a.dropped_frames == 3       # three messages ahead, the rest dropped
        assert any(e["event"] == "ERROR" and "faster than real time" in e["message"] for e in sent)
        a.cleanup()
        c.cleanup()
    asyncio.run(main())
    assert not svc._managers


def test_a_failing_tick_is_reported_and_the_loop_goes_on(monkeypatch):
    from real_time_vad.service import batched_server as bs
    monkeypatch.setattr(bs, "BatchedVADManager", _StubManager)
    svc = bs.BatchedVADService(max_clients=4, tick_s=0.001)
    sent = []

    async def send(text):
        sent.append(json.loads(text))

    async def main():
        svc.connect(send, bs.create_client_config({}))
        task = asyncio.create_task(svc.run())
        await asyncio.sleep(0.05)
        assert not task.done()                                    # the ticker survived its failing steps
        task.cancel()
    asyncio.run(main())
    assert svc.failed_ticks >= 2 and "simulated" in svc.last_error
    assert any(e["event"] == "ERROR" and "VAD step failed" in e["message"] for e in sent)
