"""CPU: the native stream feeder's host logic (cvad_feeder_*, include/cutter_vad_b200.h) without a GPU --
framing / carry-over against AudioUtils.split_into_frames, and the callback side against the oracle's literal
restatement of VADProcessor's voice buffering (oracle/vad_oracle.py: StateMachine + VoiceBuffer).  The GPU step in the
middle is covered by tests/test_gpu_dropin_api.py (BatchedVADManager runs on the same feeder)."""
import numpy as np
import pytest

from conftest import synth_streams


def _feeder(**kw):
    from real_time_vad.engine.feeder import StreamFeeder
    return StreamFeeder(None, **kw)


@pytest.mark.parametrize("frame_len,hop", [(512, 512), (512, 256), (480, 480), (512, 301), (400, 640)])
def test_gather_equals_split_into_frames_with_carry_over(frame_len, hop):
    """Whatever the push sizes, the frames stepped for a stream are split_into_frames(whole audio) in order
    (audio.py:164-190), leftovers are carried, and a stream runs at most capacity_frames per step."""
    from vad_oracle import split_into_frames
    n, L = 7, 9000
    audio = synth_streams(n, L, seed=3)
    f = _feeder(max_streams=16, frame_len=frame_len, hop=hop, capacity_frames=4)
    slots = [2, 3, 5, 7, 11, 13, 15]
    for s in slots:
        f.open(s)
    rng = np.random.default_rng(1)
    pos = np.zeros(n, int)
    got = {s: [] for s in slots}
    for it in range(400):
        for k, s in enumerate(slots):
            if pos[k] < L:
                m = int(rng.integers(1, 1400))
                f.push(s, np.ascontiguousarray(audio[k, pos[k]:pos[k] + m]))
                pos[k] += m
        r = f.gather_only()
        assert r.counts.max(initial=0) <= 4
        for row, (s, c) in enumerate(zip(r.slots, r.counts)):
            for j in range(c):
                got[int(s)].append(r.raw[row, j * hop:j * hop + frame_len].copy())
        if (pos >= L).all() and r.slots.size == 0:
            break
    for k, s in enumerate(slots):
        want = split_into_frames(audio[k], frame_len, hop)
        assert len(got[s]) == len(want), (s, len(got[s]), len(want))
        assert np.array_equal(np.stack(got[s]), want)
    f.close()


def test_push_many_equals_push_and_rows_grow():
    n = 40
    block = synth_streams(n, 3000, seed=9)
    a = _feeder(max_streams=64, frame_len=512, hop=256, capacity_frames=2)   # rows start at 512 + 3*512 samples
    b = _feeder(max_streams=64, frame_len=512, hop=256, capacity_frames=2)
    ids = np.arange(n) * 1 + 5
    for s in ids:
        a.open(int(s))
        b.open(int(s))
    for lo in range(0, 3000, 750):                                           # 3,000 pending samples: rows must grow
        a.push_many(ids, np.ascontiguousarray(block[:, lo:lo + 750]))
        for k, s in enumerate(ids):
            b.push(int(s), np.ascontiguousarray(block[k, lo:lo + 750]))
    assert a.pending(int(ids[0])) == b.pending(int(ids[0])) == 3000
    for _ in range(6):
        ra, rb = a.gather_only(), b.gather_only()
        assert np.array_equal(ra.slots, rb.slots) and np.array_equal(ra.counts, rb.counts)
        assert np.array_equal(ra.raw, rb.raw)
    assert a.pending(int(ids[0])) == 3000 - 10 * 256                         # 10 frames consumed, tail carried
    a.close()
    b.close()


def test_nonfinite_and_bad_arguments_are_rejected_before_any_state_change():
    from real_time_vad.engine import capi
    from real_time_vad.engine.feeder import FeederError
    f = _feeder(max_streams=4)
    f.open(1)
    f.push(1, np.full(100, 0.25, np.float32))
    with pytest.raises(FeederError, match="infinite or NaN") as ei:
        f.push(1, np.array([0.1, np.nan, 0.2], np.float32))
    assert ei.value.code == capi.E_INVALID
    with pytest.raises(FeederError, match="infinite or NaN"):
        f.push_many([1], np.array([[0.1, np.inf]], np.float32))
    assert f.pending(1) == 100
    with pytest.raises(FeederError, match="not open"):
        f.push(2, np.zeros(4, np.float32))
    with pytest.raises(FeederError):
        f.push(9, np.zeros(4, np.float32))
    with pytest.raises(FeederError, match="one source rate"):
        f.open(3, src_rate=48000)
    with pytest.raises(FeederError):
        _feeder(max_streams=4, src_rate=44100)
    with pytest.raises(FeederError, match="no CPU fallback"):          # stepping needs the engine
        f.step()
    f.close()


def test_mixed_rates_frame_by_each_streams_own_chunk():
    f = _feeder(max_streams=8, src_rate=0, capacity_frames=8)
    plan = {0: 8000, 1: 16000, 2: 24000, 3: 48000}
    for s, r in plan.items():
        f.open(s, src_rate=r)
    x = {s: np.arange(r * 512 // 16000 * 3 + 17, dtype=np.float32) * 1e-4 + s for s, r in plan.items()}
    for s in plan:
        f.push(s, x[s])
    r = f.gather_only()
    assert list(r.slots) == [0, 1, 2, 3] and list(r.counts) == [3, 3, 3, 3]
    for row, s in enumerate(plan):
        n_in = plan[s] * 512 // 16000
        assert np.array_equal(r.raw[row, :3 * n_in], x[s][:3 * n_in])
        assert f.pending(s) == 17
    f.close()


def _run_deliver(pcm_format, payload, denoise, frame_len, hop, seed, n=6, T=120):
    """Scripted probabilities -> oracle state machine flags -> feeder.deliver_only, against the oracle's VoiceBuffer."""
    from real_time_vad.engine import capi
    from vad_oracle import StateMachine, VoiceBuffer, denoise as gate
    rng = np.random.default_rng(seed)
    start_p, end_p = 0.5, 0.35
    probs = np.clip(0.5 + 0.5 * np.sin(np.arange(T)[None, :] / rng.uniform(3, 9, (n, 1)) + rng.uniform(0, 6, (n, 1)))
                    + 0.15 * rng.standard_normal((n, T)), 0, 1).astype(np.float32)
    L = (T - 1) * hop + frame_len
    audio = synth_streams(n, L, seed=seed)
    if pcm_format == capi.PCM_F32:
        wire, as_float = audio, audio
    else:
        wire = np.clip(np.round(audio * 32767.0), -32768, 32767).astype(np.int16)
        as_float = wire.astype(np.float32) / np.float32(32767.0 if pcm_format == capi.PCM_S16_32767 else 32768.0)
    f = _feeder(max_streams=max(8, n), pcm_format=pcm_format, frame_len=frame_len, hop=hop, capacity_frames=5)
    sms = [StateMachine(start_p, end_p, 0.8, 0.95, 2 + s % 2, 3 + s % 3) for s in range(n)]
    vbs = [VoiceBuffer(start_p) for _ in range(n)]
    for s in range(n):
        f.open(s, payload=payload, vad_start_probability=start_p, enable_denoising=denoise)
    done = np.zeros(n, int)
    pos = np.zeros(n, int)
    n_seg = n_frames = n_start = 0
    while (done < T).any():
        for s in range(n):
            m = int(rng.integers(200, 2600))
            if pos[s] < L:
                f.push(s, np.ascontiguousarray(wire[s, pos[s]:pos[s] + m]))
            pos[s] += m
        g = f.gather_only()
        if g.slots.size == 0:
            continue
        Tm = int(g.counts.max())
        p = np.zeros((g.slots.size, Tm), np.float32)
        fl = np.zeros((g.slots.size, Tm), np.uint8)
        want = []
        for row, (s, c) in enumerate(zip(g.slots, g.counts)):
            for j in range(c):
                t = done[s] + j
                p[row, j] = probs[s, t]
                fl[row, j] = sms[s].step(float(probs[s, t]))
                frame = as_float[s, t * hop:t * hop + frame_len]
                frame = gate(frame).astype(np.float32) if denoise else frame
                seg, cont = vbs[s].step(frame, float(probs[s, t]), int(fl[row, j]))
                if fl[row, j] & 3 or (cont is not None and payload == capi.PAYLOAD_FRAMES):
                    want.append((int(s), j, int(fl[row, j]), seg, cont))
            done[s] += c
        r = f.deliver_only(p, fl)
        assert [(d.slot, d.step_frame, d.flags) for d in r.deliveries] == [(w[0], w[1], w[2]) for w in want]
        for d, (_, _, b, seg, cont) in zip(r.deliveries, want):
            if b & 1:
                n_start += 1
            if b & 2:
                n_seg += 1
                if payload >= capi.PAYLOAD_SEGMENTS:
                    assert np.array_equal(d.segment, seg)
                else:
                    assert d.segment is None or d.segment.size == 0
            if payload == capi.PAYLOAD_FRAMES and cont is not None:
                n_frames += 1
                assert np.array_equal(d.frame, cont)
            else:
                assert d.frame is None
        for s in range(n):
            assert f.is_active(s) == vbs[s].active == sms[s].active
    f.close()
    return n_start, n_seg, n_frames


@pytest.mark.parametrize("pcm_format", [0, 1, 2])
@pytest.mark.parametrize("payload,denoise,frame_len,hop", [(3, True, 480, 480), (2, True, 512, 256), (3, False, 512, 512),
                                                           (1, True, 512, 512)])
def test_deliver_equals_the_reference_voice_buffering(pcm_format, payload, denoise, frame_len, hop):
    n_start, n_seg, n_frames = _run_deliver(pcm_format, payload, denoise, frame_len, hop, seed=17 + payload)
    assert n_start >= 6 and n_seg >= 6
    assert (n_frames > 100) == (payload == 3)


def test_deliver_on_the_helper_pool_keeps_stream_order_and_payloads():
    """Enough payload-carrying streams (>= 128) for the segment assembly to run on the feeder's helper threads, each over its
    own range of streams with its own record / payload storage: the merged records are still in stream-then-frame order
    and every payload equals the oracle's."""
    n_start, n_seg, n_frames = _run_deliver(1, 3, True, 480, 480, seed=5, n=300, T=24)
    assert n_start >= 100 and n_seg >= 50 and n_frames > 1000


def test_streams_without_payloads_only_keep_the_active_mirror():
    from real_time_vad.engine import capi
    f = _feeder(max_streams=4, frame_len=512, hop=512)
    f.open(0)
    f.open(1, payload=capi.PAYLOAD_EVENTS, vad_start_probability=0.5)
    for s in (0, 1):
        f.push(s, np.zeros(1024, np.float32))
    f.gather_only()
    fl = np.array([[0, capi.FLAG_STARTED], [0, capi.FLAG_STARTED]], np.uint8)
    r = f.deliver_only(np.array([[0.1, 0.9], [0.1, 0.9]], np.float32), fl)
    assert [(d.slot, d.step_frame) for d in r.deliveries] == [(1, 1)]
    assert f.is_active(0) and f.is_active(1)
    f.clear(0)
    assert not f.is_active(0) and f.pending(0) == 0
    f.close()
