"""The websocket server's VAD seam on the batched engine.

The reference server (websocket_service/server/vad_websocket_server.py) builds one `ClientState` per connection,
each with its own `VADWrapper` + onnxruntime session (:248-290), and runs one model call per binary message on the
event loop (:326-369).  Here a connection is a `ClientSession` holding one SLOT of a shared `BatchedVADManager`:

    message in   -> ClientSession.process_audio_frame(bytes)   size check as :334-337, then one memcpy into the
                                                                pinned stream arena (manager.push_bytes); no model call
    every tick   -> BatchedVADService.tick()                    ONE GPU step for every client that has a whole frame
    events out   -> the same JSON the reference sends (:83-124, :454-500): INFO / ERROR / VOICE_START /
                    VOICE_CONTINUE / VOICE_END {segment_start_ms, segment_end_ms, duration_ms} / TIMEOUT, with the
                    same `segment_index` bookkeeping (incremented after VOICE_END, :489-492)

Protocol, query parameters, CONFIG / HEARTBEAT control messages, defaults (0.4 / 0.3 / 6 / 12, 30 ms int16 frames,
:565-585) are the reference's.  Out of scope, as in SURVEY.md section 2 row 11: Opus / AAC decoding (PyAV).
`create_app()` wires the sessions into a FastAPI app with the reference's routes (`/vad`, `/`, `/health`, `/clients`).
"""
from __future__ import annotations

import asyncio
import json
import logging
import threading
import time
import uuid
from typing import Any, Awaitable, Callable, Dict, List, Optional, Tuple
from urllib.parse import parse_qs

from pydantic import BaseModel, ValidationError

from ..core.batched_manager import BatchedVADManager
from ..core.config import SampleRate, SileroModelVersion, VADConfig
from ..engine import capi

_INT_KEYS = ("sample_rate", "channels", "sample_width", "frame_duration_ms", "start_frame_count", "end_frame_count")
_FLOAT_KEYS = ("start_probability", "end_probability", "start_ratio", "end_ratio", "timeout")
_AUDIO_KEYS = ("mode", "sample_rate", "channels", "sample_width", "frame_duration_ms")
_VAD_KEYS = ("start_probability", "end_probability", "start_frame_count", "end_frame_count", "start_ratio", "end_ratio")


class AudioMode(BaseModel):                       # vad_websocket_server.py:41-47
    mode: str = "pcm"
    sample_rate: int = 16000
    channels: int = 1
    sample_width: int = 2
    frame_duration_ms: int = 30


class VADParameters(BaseModel):                   # :50-57
    start_probability: float = 0.4
    end_probability: float = 0.3
    start_frame_count: int = 6
    end_frame_count: int = 12
    start_ratio: float = 0.8
    end_ratio: float = 0.95


class ClientConfig(BaseModel):                    # :60-64
    audio: AudioMode = AudioMode()
    vad: VADParameters = VADParameters()
    timeout: float = 0.0


def parse_query_params(query_string: str) -> Dict[str, Any]:
    """Connection URL parameters, typed as the reference types them (:512-531)."""
    out: Dict[str, Any] = {}
    for key, values in parse_qs(query_string or "").items():
        if not values:
            continue
        v = values[0]
        out[key] = int(v) if key in _INT_KEYS else float(v) if key in ("start_probability", "end_probability", "timeout") else v
    return out


def create_client_config(query_params: Dict[str, Any], config_message: Optional[Dict[str, Any]] = None) -> ClientConfig:
    """Defaults <- query parameters <- CONFIG message (:534-611)."""
    audio, vad, timeout = AudioMode().model_dump(), VADParameters().model_dump(), 0.0
    for source in (query_params, {k: v for k, v in (config_message or {}).items() if v is not None}):
        for key, value in source.items():
            if key in _AUDIO_KEYS:
                audio[key] = value
            elif key in _VAD_KEYS:
                vad[key] = value
            elif key == "timeout":
                timeout = value
    return ClientConfig(audio=AudioMode(**audio), vad=VADParameters(**vad), timeout=timeout)


def _now_ms() -> int:
    return int(time.time() * 1000)


class ClientSession:
    """One websocket client = one slot of a shared manager (the reference's ClientState, :209-505)."""

    def __init__(self, service: "BatchedVADService", client_id: str, send_text: Callable[[str], Awaitable[None]],
                 config: ClientConfig):
        self.service = service
        self.client_id = client_id
        self.send_text = send_text
        self.config = config
        self.start_time = time.time()
        self.segment_index = 0
        self.voice_start_time: Optional[float] = None
        self.last_frame_time: Optional[float] = None
        self.timeout_sent = False
        self.frame_count = 0
        self.dropped_frames = 0
        self._frame_samples = 1
        self.expected_frame_bytes = self._frame_bytes()
        self._manager: Optional[BatchedVADManager] = None
        self._stream: Optional[int] = None
        self._pending: List[Tuple[str, float]] = []     # (event kind, time it was detected), filled by the step thread
        self._attach()

    # ------------------------------------------------------------------ configuration
    def _frame_bytes(self) -> int:
        a = self.config.audio
        return int(a.sample_rate * (a.frame_duration_ms / 1000) * a.channels * a.sample_width)     # :237-246

    def _attach(self) -> None:
        a, v = self.config.audio, self.config.vad
        if a.mode != "pcm":
            raise ValueError(f"PyAV is required for {a.mode} decoding but not installed")
        if a.channels != 1:
            raise ValueError("the batched service takes mono streams (channels=1)")
        if a.sample_width not in (2, 4):
            raise ValueError(f"Unsupported sample width: {a.sample_width}")
        frame_samples = int(a.sample_rate * (a.frame_duration_ms / 1000))
        self._frame_samples = max(1, frame_samples)
        cfg = VADConfig(sample_rate=SampleRate(a.sample_rate), model_version=SileroModelVersion.V5,
                        vad_start_probability=v.start_probability, vad_end_probability=v.end_probability,
                        voice_start_ratio=v.start_ratio, voice_end_ratio=v.end_ratio,
                        voice_start_frame_count=v.start_frame_count, voice_end_frame_count=v.end_frame_count,
                        enable_denoising=True, auto_convert_sample_rate=True, buffer_size=frame_samples)   # :259-272
        self._manager = self.service.manager_for(a.sample_rate, frame_samples, a.sample_width)
        self._stream = self._manager.open_stream(cfg, on_voice_start=self._on_voice_start, on_voice_end=self._on_voice_end,
                                                 on_voice_continue=self._on_voice_continue)

    def _detach(self) -> None:
        if self._manager is not None and self._stream is not None:
            self._manager.close_stream(self._stream)
            self.service.release_if_idle(self._manager)
        self._manager, self._stream = None, None

    def update_config(self, new_config: ClientConfig) -> None:
        """:306-324: the VAD side is rebuilt when its parameters or the audio format change."""
        old, self.config = self.config, new_config
        if old.audio != new_config.audio:
            self.expected_frame_bytes = self._frame_bytes()
        if old.vad != new_config.vad or old.audio != new_config.audio:
            with self.service.step_lock:
                self._detach()
                self._attach()

    # ------------------------------------------------------------------ data path
    async def process_audio_frame(self, frame_data: bytes) -> None:
        """:326-383 without the model call: validate, append to the stream's pending audio."""
        try:
            if len(frame_data) != self.expected_frame_bytes:
                await self._send_error(f"Invalid frame size: expected {self.expected_frame_bytes}, got {len(frame_data)}")
                return
            # backpressure: the reference ran the model inside this call, so a client could never be more than one
            # message ahead; here audio waits for the next tick, and a flooding client must not grow the shared arena
            if self._manager.pending(self._stream) >= self.service.max_pending_frames * self._frame_samples:
                self.dropped_frames += 1
                if self.dropped_frames == 1 or self.dropped_frames % 100 == 0:
                    await self._send_error(f"Audio arrives faster than real time: {self.dropped_frames} frame(s) dropped")
                return
            self.frame_count += 1
            self._manager.push_bytes(self._stream, frame_data)
            self.last_frame_time = time.time()
            self.timeout_sent = False
        except Exception as exc:
            await self._send_error(f"Audio processing error: {exc}")

    # callbacks: fired by manager.step() on the step thread; the JSON goes out on the event loop afterwards
    def _on_voice_start(self) -> None:
        self._pending.append(("start", time.time()))

    def _on_voice_end(self, wav_data: bytes) -> None:
        self._pending.append(("end", time.time()))

    def _on_voice_continue(self, pcm_data: bytes) -> None:
        self._pending.append(("continue", time.time()))

    async def flush_events(self) -> None:
        """Send what the last step detected, in order (start, end, continue per frame: vad_wrapper.py:498-519)."""
        pending, self._pending = self._pending, []
        for kind, at in pending:
            if kind == "start":
                self.voice_start_time = at
                await self._send({"event": "VOICE_START", "timestamp_ms": int(at * 1000), "segment_index": self.segment_index})
            elif kind == "continue":
                await self._send({"event": "VOICE_CONTINUE", "timestamp_ms": int(at * 1000), "segment_index": self.segment_index})
            else:
                start_ms = int((self.voice_start_time if self.voice_start_time else at) * 1000)
                end_ms = int(at * 1000)
                await self._send({"event": "VOICE_END", "timestamp_ms": end_ms, "segment_index": self.segment_index,
                                  "segment_start_ms": start_ms, "segment_end_ms": end_ms, "duration_ms": end_ms - start_ms})
                self.segment_index += 1
        if (self.config.timeout > 0 and self.last_frame_time and not self.timeout_sent
                and time.time() - self.last_frame_time >= self.config.timeout):
            self.timeout_sent = True
            await self._send({"event": "TIMEOUT", "timestamp_ms": _now_ms(), "segment_index": None,
                              "message": "no voice detected in configured timeout"})

    # ------------------------------------------------------------------ events
    async def _send(self, event: Dict[str, Any]) -> None:
        try:
            await self.send_text(json.dumps(event))
        except Exception:
            pass                                   # a closed socket must not stop the tick (the reference logs and goes on, :642-649)

    async def _send_error(self, message: str) -> None:
        await self._send({"event": "ERROR", "timestamp_ms": _now_ms(), "segment_index": None, "message": message})

    async def _send_info(self, message: str) -> None:
        await self._send({"event": "INFO", "timestamp_ms": _now_ms(), "segment_index": None, "message": message})

    def cleanup(self) -> None:
        with self.service.step_lock:
            self._detach()
        self.service.sessions.pop(self.client_id, None)


class BatchedVADService:
    """All clients of one process: managers keyed by wire format, one GPU step per tick for everybody."""

    def __init__(self, max_clients: int = 10_000, device: Optional[int] = None, tick_s: float = 0.010,
                 max_formats: int = 4, max_pending_frames: int = 64) -> None:
        self.max_clients = max_clients
        self.device = device
        self.tick_s = tick_s
        # a manager owns a max_clients-slot engine and a pinned arena: remote clients choose the wire format, so the number
        # of distinct formats alive at once is capped, idle managers are freed, and a stream may be at most
        # max_pending_frames messages ahead of the tick
        self.max_formats = max_formats
        self.max_pending_frames = max_pending_frames
        self.last_error: Optional[str] = None
        self.failed_ticks = 0
        self.sessions: Dict[str, ClientSession] = {}
        self._managers: Dict[Tuple[int, int, int], BatchedVADManager] = {}
        self.step_lock = threading.RLock()         # open / close / reconfigure vs the step thread
        self.ticks = 0
        self.frames = 0

    def manager_for(self, sample_rate: int, frame_samples: int, sample_width: int) -> BatchedVADManager:
        key = (sample_rate, frame_samples, sample_width)
        with self.step_lock:
            m = self._managers.get(key)
            if m is None:
                if len(self._managers) >= self.max_formats:
                    raise ValueError(f"too many distinct audio formats in use ({len(self._managers)}); "
                                     f"supported at once: {self.max_formats}")
                pcm = capi.PCM_S16_32767 if sample_width == 2 else capi.PCM_F32          # server.py:341 divides by 32767
                if sample_rate == 16000:
                    m = BatchedVADManager(max_streams=self.max_clients, device=self.device, frame_len=frame_samples,
                                          hop=frame_samples, pcm_format=pcm)
                else:                                                                    # resampled on the GPU, 32 ms chunks
                    m = BatchedVADManager(max_streams=self.max_clients, device=self.device, pcm_format=pcm,
                                          source_rate=sample_rate)
                self._managers[key] = m
            return m

    def release_if_idle(self, manager: BatchedVADManager) -> None:
        """Free a manager (engine, device state, pinned arena) once its last stream has closed."""
        with self.step_lock:
            if manager.open_streams:
                return
            for key, m in list(self._managers.items()):
                if m is manager:
                    del self._managers[key]
                    m.close()

    def connect(self, send_text: Callable[[str], Awaitable[None]], config: ClientConfig,
                client_id: Optional[str] = None) -> ClientSession:
        cid = client_id or str(uuid.uuid4())
        with self.step_lock:
            s = ClientSession(self, cid, send_text, config)
        self.sessions[cid] = s
        return s

    def step(self) -> int:
        """One GPU step per wire format (normally one); callbacks fill the sessions' pending lists.  Thread-safe
        against connects / disconnects; pushes never wait for it."""
        n = 0
        with self.step_lock:
            for m in list(self._managers.values()):
                n += m.step().frames
        self.ticks += 1
        self.frames += n
        return n

    async def tick(self) -> int:
        """Step on a worker thread, then send the detected events from the event loop."""
        n = await asyncio.get_running_loop().run_in_executor(None, self.step)
        for s in list(self.sessions.values()):
            if s._pending or s.config.timeout > 0:
                await s.flush_events()
        return n

    async def run(self) -> None:
        """The process's one step loop.  A failing tick (a CUDA error, an allocation failure, a callback error) must not
        end the loop silently while clients keep sending: it is logged, reported to every session and on /health, and
        the loop goes on."""
        while True:
            t0 = time.perf_counter()
            try:
                await self.tick()
                self.last_error = None
            except asyncio.CancelledError:
                raise
            except Exception as exc:
                self.failed_ticks += 1
                self.last_error = f"{type(exc).__name__}: {exc}"
                logging.getLogger(__name__).exception("VAD tick failed")
                for s in list(self.sessions.values()):
                    await s._send_error(f"VAD step failed: {exc}")
            await asyncio.sleep(max(0.0, self.tick_s - (time.perf_counter() - t0)))

    def close(self) -> None:
        with self.step_lock:
            for m in self._managers.values():
                m.close()
            self._managers.clear()
        self.sessions.clear()


def create_app(service: Optional[BatchedVADService] = None):
    """FastAPI app with the reference's routes (:617-797) on a BatchedVADService."""
    from fastapi import FastAPI, WebSocket, WebSocketDisconnect

    from contextlib import asynccontextmanager

    svc = service or BatchedVADService()

    @asynccontextmanager
    async def lifespan(app):
        ticker = asyncio.create_task(svc.run())       # the one step loop of the process

        def _ticker_done(task: "asyncio.Task") -> None:
            if not task.cancelled() and task.exception() is not None:
                svc.last_error = f"ticker stopped: {task.exception()!r}"
                logging.getLogger(__name__).error("VAD ticker stopped: %r", task.exception())
        ticker.add_done_callback(_ticker_done)
        try:
            yield
        finally:
            ticker.cancel()
            svc.close()

    app = FastAPI(title="VAD WebSocket Server", description="Real-time Voice Activity Detection WebSocket Server (batched B200 engine)",
                  version="1.0.0", lifespan=lifespan)
    app.state.service = svc

    async def _refuse(ws, message: str) -> None:
        await ws.send_text(json.dumps({"event": "ERROR", "message": message, "timestamp_ms": _now_ms()}))
        await ws.close()

    @app.websocket("/vad")
    async def vad(ws: WebSocket) -> None:
        await ws.accept()
        query = parse_query_params(str(ws.query_params))
        try:
            config = create_client_config(query)
        except ValidationError as exc:
            return await _refuse(ws, f"Invalid configuration: {exc}")
        a = config.audio
        if a.mode not in ("pcm", "opus", "aac"):
            return await _refuse(ws, f"Unsupported audio mode: {a.mode}")
        if a.mode != "pcm":
            return await _refuse(ws, f"PyAV is required for {a.mode} decoding but not installed")
        frame_bytes = a.sample_rate * (a.frame_duration_ms / 1000) * a.channels * a.sample_width
        if frame_bytes != int(frame_bytes):
            return await _refuse(ws, f"Frame duration {a.frame_duration_ms}ms produces non-integer bytes ({frame_bytes})")
        try:
            session = svc.connect(ws.send_text, config)
        except Exception as exc:
            return await _refuse(ws, f"Invalid configuration: {exc}")
        await session._send_info("VAD WebSocket server ready")
        try:
            while True:
                message = await ws.receive()
                if message.get("type") == "websocket.disconnect":
                    break
                if message.get("bytes") is not None:
                    await session.process_audio_frame(message["bytes"])
                elif message.get("text") is not None:
                    try:
                        data = json.loads(message["text"])
                        if data.get("type") == "CONFIG":
                            fields = {k: data.get(k) for k in _AUDIO_KEYS + _VAD_KEYS + ("timeout",)}
                            session.update_config(create_client_config(query, fields))
                            await session._send_info("Configuration updated")
                        elif data.get("type") == "HEARTBEAT":
                            await session._send_info("Heartbeat received")
                        else:
                            await session._send_error(f"Unknown message type: {data.get('type')}")
                    except json.JSONDecodeError as exc:
                        await session._send_error(f"Invalid JSON: {exc}")
                    except (ValidationError, ValueError) as exc:
                        await session._send_error(f"Invalid message format: {exc}")
        except WebSocketDisconnect:
            pass
        finally:
            session.cleanup()

    @app.get("/")
    async def root():
        return {"message": "VAD WebSocket Server", "status": "running", "connected_clients": len(svc.sessions),
                "timestamp": _now_ms()}

    @app.get("/health")
    async def health():
        return {"status": "healthy" if svc.last_error is None else "degraded", "last_error": svc.last_error,
                "failed_ticks": svc.failed_ticks, "connected_clients": len(svc.sessions), "timestamp": _now_ms()}

    @app.get("/clients")
    async def clients():
        info = {cid: {"connected_at": s.start_time, "config": s.config.model_dump(), "segment_index": s.segment_index}
                for cid, s in svc.sessions.items()}
        return {"connected_clients": len(info), "clients": info, "timestamp": _now_ms()}

    return app
