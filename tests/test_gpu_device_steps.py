"""GPU: device-pointer steps (cvad_step_device), the call bench.py's `value` leg times.

One-frame v5 steps on 16 kHz input run as ONE fused kernel, and consecutive device-pointer steps are chained kernel to
kernel by programmatic dependent launch: no memset and no event record between them (the kernel clears its own status
words, counts events in an engine-owned counter that the last CTA publishes, and prefetches weight tiles before
`griddepcontrol.wait`).  Everything below compares that path, bit for bit, with the host-buffer call `cvad_step` on a
second engine fed the same audio: probabilities, flags, event records and counts, for partial tiles, dead tiles
(n_frames == 0 for a whole tile), slot indirection, NaN input, with timing events between the kernels, with host
steps interleaved on the same engine, and on source-rate input (one rate, per-stream rates, an unknown rate), where the
resampler kernel runs ahead of the chained model kernel.
"""
import numpy as np
import pytest

from conftest import synth_streams

pytestmark = pytest.mark.gpu


def _dev_args(capi, torch, n, audio_d, probs_d, flags_d, events_d, nev_d, slots_d=None, nfr_d=None, max_events=0,
              src_rate=16000, rates_d=None):
    a = capi.StepArgs()
    a.n_streams = n
    a.audio = audio_d.data_ptr()
    a.pcm_format = capi.PCM_F32
    a.stream_stride = audio_d.shape[1]
    a.max_frames = 1
    # source-rate input: one chunk of 512 * rate / 16000 samples per model frame (with per-stream rates the row is
    # sized for the fastest one and frame_len / hop stay at the model's 512)
    a.frame_len = a.hop = 512 if rates_d is not None else src_rate * 512 // 16000
    a.src_rate = src_rate
    if rates_d is not None:
        a.src_rates = rates_d.data_ptr()
    if slots_d is not None:
        a.slots = slots_d.data_ptr()
    if nfr_d is not None:
        a.n_frames = nfr_d.data_ptr()
    a.probs_out = probs_d.data_ptr()
    a.flags_out = flags_d.data_ptr()
    a.events_out = events_d.data_ptr()
    a.max_events = max_events or 2 * n
    a.n_events_out = nev_d.data_ptr()
    return a


def _events_from(events_d, nev_d, max_events):
    k = min(int(nev_d.cpu()[0]), max_events)
    rec = events_d.cpu().numpy().view(np.int64)[:3 * k]
    r32 = rec.view(np.int32).reshape(k, 6)
    sf = rec.reshape(k, 3)[:, 2]
    return sorted((int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(f)) for r, f in zip(r32, sf))


def _run_pair(engine_factory, math, n, T, *, slots=None, n_frames=None, nan_at=None, timing=False, host_every=0,
              src_rate=16000, src_rates=None, dev_src_rates=None):
    """T one-frame steps on two engines: `dev` through cvad_step_device (all steps enqueued back to back, results read
    from per-step device buffers afterwards), `host` through cvad_step.  -> per-step (probs, flags, events, count)."""
    import torch
    from real_time_vad.engine import capi

    W = (max(src_rates) if src_rates is not None else src_rate) * 512 // 16000      # samples per stream and step
    audio = synth_streams(n, W * T, seed=77)
    if nan_at is not None:
        s, j = nan_at
        audio[s, W * j + 100] = np.nan
    rate_kw = ({"src_rates": np.asarray(src_rates, np.int32)} if src_rates is not None else {"src_rate": src_rate})
    cap = 512
    kw = dict(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=2, voice_end_frame_count=3,
              enable_denoising=True)
    dev = engine_factory(cap, math=math)
    host = engine_factory(cap, math=math)
    for e in (dev, host):
        e.reset()
        e.configure(**kw)
    dev.set_timing(timing)
    cu = "cuda:0"
    frames = [torch.from_numpy(np.ascontiguousarray(audio[:, W * j:W * (j + 1)])).to(cu) for j in range(T)]
    rates_d = None if src_rates is None else torch.from_numpy(
        np.asarray(src_rates if dev_src_rates is None else dev_src_rates, np.int32)).to(cu)
    slots_d = None if slots is None else torch.from_numpy(np.asarray(slots, np.int32)).to(cu)
    nfr = None if n_frames is None else [torch.from_numpy(np.asarray(n_frames(j), np.int32)).to(cu) for j in range(T)]
    probs_d = [torch.full((n, 1), -1.0, dtype=torch.float32, device=cu) for _ in range(T)]
    flags_d = [torch.full((n, 1), 255, dtype=torch.uint8, device=cu) for _ in range(T)]
    events_d = [torch.zeros((2 * n * 24,), dtype=torch.uint8, device=cu) for _ in range(T)]
    nev_d = [torch.full((1,), 12345, dtype=torch.int32, device=cu) for _ in range(T)]   # stale: the step must overwrite it
    torch.cuda.synchronize()
    got = [None] * T
    for j in range(T):
        if host_every and j % host_every == host_every - 1:
            # a host-buffer step on the SAME engine in the middle of a chain (its lane stream must wait for the chain)
            r = dev.step(np.ascontiguousarray(audio[:, W * j:W * (j + 1)]), slots=slots,
                         n_frames=None if n_frames is None else n_frames(j), max_frames=1, **rate_kw)
            got[j] = (r.probs.copy(), r.flags.copy(), sorted(r.events), len(r.events))
            continue
        a = _dev_args(capi, torch, n, frames[j], probs_d[j], flags_d[j], events_d[j], nev_d[j], slots_d,
                      None if nfr is None else nfr[j], src_rate=src_rate, rates_d=rates_d)
        dev.step_device(a)
    dev.sync()
    if timing:
        dev.read_timing()
        dev.set_timing(False)
    want = []
    for j in range(T):
        r = host.step(np.ascontiguousarray(audio[:, W * j:W * (j + 1)]), slots=slots,
                      n_frames=None if n_frames is None else n_frames(j), max_frames=1, **rate_kw)
        want.append((r.probs.copy(), r.flags.copy(), sorted(r.events), len(r.events), r.status.copy()))
        if got[j] is None:
            ev = _events_from(events_d[j], nev_d[j], 2 * n)
            got[j] = (probs_d[j].cpu().numpy(), flags_d[j].cpu().numpy(), ev, int(nev_d[j].cpu()[0]))
    return got, want, (dev, host)


def _compare(got, want, live=None):
    n_events = 0
    for j, (g, w) in enumerate(zip(got, want)):
        m = np.ones(len(w[0]), bool) if live is None else live(j)
        assert np.array_equal(g[0][m], w[0][m]), f"probabilities differ at step {j}"
        assert np.array_equal(g[1][m], w[1][m]), f"flags differ at step {j}"
        assert g[3] == w[3], f"event count differs at step {j}: {g[3]} vs {w[3]}"
        assert g[2] == w[2], f"event records differ at step {j}"
        n_events += w[3]
    return n_events


@pytest.mark.parametrize("math", ["tc16", "tc"])
def test_chained_device_steps_equal_host_steps(engine_factory, math):
    n, T = 300, 48                                   # 10 tiles, the last one partial
    got, want, (dev, host) = _run_pair(engine_factory, math, n, T)
    assert _compare(got, want) > 20                  # the state machine fired, and every record came through
    for s in (0, 137, 299):                          # resident state after the chain
        hd, cd, smd, fd = dev.get_state(s)
        hh, ch, smh, fh = host.get_state(s)
        assert np.array_equal(hd, hh) and np.array_equal(cd, ch) and np.array_equal(smd, smh) and fd == fh == T


def test_chained_steps_with_dead_tiles_and_slot_indirection(engine_factory):
    n, T = 200, 30
    rng = np.random.default_rng(3)
    slots = rng.permutation(512)[:n]

    def n_frames(j):
        nf = np.ones(n, np.int32)
        nf[32:64] = 0                                # a whole tile without work in every step: its CTA only drains
        nf[(np.arange(n) + j) % 7 == 0] = 0          # and streams that sit a step out
        if j % 5 == 4:
            nf[:] = 0                                # a step in which nothing is live at all
            nf[190] = 1
        return nf

    got, want, _ = _run_pair(engine_factory, "tc16", n, T, slots=slots, n_frames=n_frames)
    assert _compare(got, want, live=lambda j: n_frames(j) > 0) > 5


def test_chained_steps_skip_a_non_finite_stream_like_host_steps(engine_factory):
    n, T = 96, 20
    got, want, (dev, host) = _run_pair(engine_factory, "tc16", n, T, nan_at=(41, 6))
    assert want[6][4][41] != 0                       # the host call reports the stream ...
    keep = np.ones(n, bool)
    keep[41] = False
    _compare(got, want, live=lambda j: keep if j == 6 else np.ones(n, bool))
    # ... and in both paths it keeps its state for that step (the status word was cleared by the kernel itself in
    # the steps that followed: the stream runs on)
    assert dev.get_state(41)[3] == host.get_state(41)[3] == T - 1


def test_device_steps_with_timing_events_and_interleaved_host_steps(engine_factory):
    got, want, _ = _run_pair(engine_factory, "tc16", 130, 24, timing=True)
    _compare(got, want)
    got, want, _ = _run_pair(engine_factory, "tc16", 130, 24, host_every=4)
    _compare(got, want)


def test_chained_steps_on_resampled_input(engine_factory):
    """Source-rate input: the resampler kernel runs ahead of the fused kernel, still with nothing but kernels on the
    stream (the model kernel is scheduled while the resampler runs).  One rate for all streams, then per-stream rates
    in one step (rate_lists_kernel writes the status words there)."""
    got, want, _ = _run_pair(engine_factory, "tc16", 150, 24, src_rate=48000, host_every=5)
    assert _compare(got, want) >= 1
    got, want, _ = _run_pair(engine_factory, "tc16", 100, 16, src_rate=8000)
    _compare(got, want)
    n = 140
    rates = np.array([(24000, 48000, 16000, 8000)[i % 4] for i in range(n)], np.int32)
    got, want, _ = _run_pair(engine_factory, "tc16", n, 16, src_rates=rates)
    _compare(got, want)
    # a rate the engine does not have (the host call rejects the whole step for it; a device-pointer step cannot look
    # at the array): that stream is flagged on the device and sits the step out, its neighbours are not affected
    bad = rates.copy()
    bad[77] = 44100
    got, want, (dev, host) = _run_pair(engine_factory, "tc16", n, 16, src_rates=rates, dev_src_rates=bad)
    keep = np.ones(n, bool)
    keep[77] = False
    for j, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g[0][keep], w[0][keep]) and np.array_equal(g[1][keep], w[1][keep]), j
        assert g[0][77, 0] == -1.0                   # never written
        assert g[2] == [e for e in w[2] if e[0] != 77]
    assert dev.get_state(77)[3] == 0 and host.get_state(77)[3] == 16
