#!/usr/bin/env python
"""Service-mode benchmark (BASELINE.json configs[4]): N live client streams on one GPU through
`BatchedVADManager`, websocket defaults (30 ms int16 messages = 480 samples, thresholds 0.4/0.3/6/12,
reference websocket_service/server/vad_websocket_server.py:252,:273,:565-572), arrival jitter.

Every tick (= 30 ms of stream time) each client delivers 0, 1 or 2 messages (jitter); then one
`manager.step()` runs every complete frame of every client.  Reported: per-tick wall latency of
(pushes + step) and of step alone (p50/p99), i.e. message-in -> event-out for the batch, and the
real-time factor.  Only the manager's calls are timed; building the synthetic clients' messages is not.  Writes one JSON line.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=10000)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--callbacks", type=float, default=0.1, help="fraction of streams with start/end callbacks")
    ap.add_argument("--per-message-push", action="store_true", help="one Python push() per message (server-like)")
    ap.add_argument("--bytes", action="store_true", help="with --per-message-push: push_bytes(wire message) instead of a numpy array")
    args = ap.parse_args()

    from bench import synth_audio
    from real_time_vad import BatchedVADManager, VADConfig
    from real_time_vad.engine import capi

    n = args.streams
    mgr = BatchedVADManager(max_streams=n, frame_len=480, hop=480, pcm_format=capi.PCM_S16_32767)
    cfg = VADConfig(buffer_size=480, vad_start_probability=0.4, vad_end_probability=0.3,
                    voice_start_frame_count=6, voice_end_frame_count=12)
    ev_count = [0, 0]
    ids = []
    n_cb = int(n * args.callbacks)
    for s in range(n):
        if s < n_cb:
            ids.append(mgr.open_stream(cfg, on_voice_start=lambda: ev_count.__setitem__(0, ev_count[0] + 1),
                                       on_voice_end=lambda b: ev_count.__setitem__(1, ev_count[1] + 1)))
        else:
            ids.append(mgr.open_stream(cfg))
    ids = np.array(ids)
    rng = np.random.default_rng(0)
    sec = 4
    audio = np.clip(np.round(synth_audio(n, 16000 * sec, seed=1) * 32767.0), -32768, 32767).astype(np.int16)
    pos = np.zeros(n, np.int64)
    lat_tick, lat_step, frames, events, phases = [], [], 0, 0, []
    for t in range(args.ticks + 10):
        k = rng.choice([0, 1, 1, 1, 1, 1, 1, 2], size=n)               # jitter: late / on time / catching up
        t_push = 0.0
        for m in (1, 2):
            sel = np.flatnonzero(k >= m)
            if sel.size == 0:
                continue
            start = pos[sel] % (16000 * sec - 480)
            block = audio[sel[:, None], start[:, None] + np.arange(480)[None, :]]   # the clients' side: not timed
            sids = [int(x) for x in ids[sel]]
            if args.per_message_push and args.bytes:
                msgs = [block[i].tobytes() for i in range(len(sids))]
                t0 = time.perf_counter()
                for sid, msg in zip(sids, msgs):
                    mgr.push_bytes(sid, msg)
            elif args.per_message_push:
                t0 = time.perf_counter()
                for i, sid in enumerate(sids):
                    mgr.push(sid, block[i])
            else:
                t0 = time.perf_counter()
                mgr.push_many(ids[sel], block)
            t_push += time.perf_counter() - t0
            pos[sel] += 480
        t1 = time.perf_counter()
        out = mgr.step()
        t2 = time.perf_counter()
        if t >= 10:
            lat_tick.append(t_push + t2 - t1)
            lat_step.append(t2 - t1)
            phases.append(getattr(out, "phase_ms", (0.0, 0.0, 0.0)))
            frames += out.frames
            events += len(out.events)
    total = sum(lat_tick)
    line = {
        "bench": "service", "streams": n, "ticks": args.ticks, "message": "480 x int16 (30 ms)",
        "callbacks_fraction": args.callbacks, "push": ("per-message bytes" if args.bytes else "per-message ndarray") if args.per_message_push else "push_many",
        "timed": "manager calls only (push* + step); the synthetic clients' own array work is outside",
        "tick_ms_p50": 1e3 * float(np.percentile(lat_tick, 50)), "tick_ms_p99": 1e3 * float(np.percentile(lat_tick, 99)),
        "step_ms_p50": 1e3 * float(np.percentile(lat_step, 50)), "step_ms_p99": 1e3 * float(np.percentile(lat_step, 99)),
        "native_step_ms_p50": dict(zip(("gather", "gpu_step", "deliver"), [float(x) for x in np.percentile(np.array(phases), 50, axis=0)])),
        "frames": frames, "events": events, "callbacks_fired": ev_count,
        "audio_s_per_s": frames * 0.030 / total, "realtime_factor_per_stream": frames * 0.030 / total / n,
        "budget_ms": 30.0,
    }
    print(json.dumps(line))
    mgr.close()


if __name__ == "__main__":
    main()
