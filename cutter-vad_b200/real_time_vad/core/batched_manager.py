"""BatchedVADManager: thousands of independent audio streams stepped per GPU launch.

New entry point named by the north star (SURVEY.md section 8b "New entry"); the
reference has no equivalent -- it builds one `VADWrapper` + one onnxruntime session per
stream (websocket_service/server/vad_websocket_server.py:248-290) and runs one model
call per frame on the event loop (:369).  Here a stream is a *slot* of resident state
in HBM (LSTM h/c, the state-machine words, thresholds); `step()` advances every stream
that has at least one complete frame buffered, in ONE `cvad_step`:

    mgr = BatchedVADManager(max_streams=10_000)
    sid = mgr.open_stream(VADConfig(...), on_voice_start=..., on_voice_end=..., on_voice_continue=...)
    mgr.push(sid, samples)            # float32 in [-1,1] or int16 PCM, any length
    out = mgr.step()                  # events in stream-then-frame order, callbacks fired

Per stream the results equal what a `VADWrapper` fed the same frames would produce:
same probabilities (FP32, 1e-4), same start/end frame indices, same callback payloads
(`voice_end` WAV of the gated frames buffered since the first above-threshold frame,
`voice_continue` float32 bytes of the gated frame; silero_model.py:839-869, :891-895, :932-949).
Framing is configurable: `frame_len` samples per frame, `hop` between frames
(hop == frame_len for the websocket service's one-frame-per-message mode, frame_len/2 for
the wrapper's overlapped mode); leftovers shorter than a frame are carried to the next step.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Callable, Dict, List, NamedTuple, Optional, Sequence

import numpy as np

from ..engine import capi
from ..engine.stream_engine import StreamEngine
from ..utils.wav_writer import WAVWriter
from .config import SileroModelVersion, VADConfig
from .exceptions import AudioProcessingError, CallbackError, ConfigurationError, VADError


class StreamEvent(NamedTuple):
    stream_id: int        # slot id returned by open_stream
    kind: str             # "start" | "end"
    frame_index: int      # frame index since the stream was opened / last reset
    step_frame: int       # frame index inside the step that produced it


class StepOutput:
    """Result of one `step()`.  Arrays are per stepped stream (row k <-> stream_ids[k])."""

    def __init__(self, events: List[StreamEvent], stream_ids: np.ndarray, counts: np.ndarray,
                 probs: np.ndarray, flags: np.ndarray):
        self.events = events                 # stream-then-frame order
        self.stream_ids = stream_ids         # [n] slot ids that ran at least one frame
        self.counts = counts                 # [n] frames run per stream
        self.probs = probs                   # [n, max(counts)] float32 (columns >= counts[k] are 0)
        self.flags = flags                   # [n, max(counts)] uint8, FLAG_* bits
        self.frames = int(counts.sum()) if counts.size else 0

    @property
    def probabilities(self) -> Dict[int, np.ndarray]:
        """stream id -> probabilities of the frames run in this step (built on demand)."""
        return {int(s): self.probs[k, :int(c)].copy() for k, (s, c) in enumerate(zip(self.stream_ids, self.counts))}


@dataclass
class _Stream:
    config: VADConfig
    writer: WAVWriter
    on_start: Optional[Callable[[], None]] = None
    on_end: Optional[Callable[[bytes], None]] = None
    on_continue: Optional[Callable[[bytes], None]] = None
    active: bool = False
    pre_roll: List[np.ndarray] = field(default_factory=list)
    segment: List[np.ndarray] = field(default_factory=list)

    @property
    def wants_audio(self) -> bool:
        return self.on_end is not None or self.on_continue is not None


class BatchedVADManager:
    def __init__(self, max_streams: int = 4096, model_version: SileroModelVersion = SileroModelVersion.V5,
                 device: Optional[int] = None, frame_len: int = 512, hop: Optional[int] = None,
                 pcm_format: int = capi.PCM_F32, source_rate: Optional[int] = 16000) -> None:
        """`source_rate` 8000 / 24000 / 48000: streams deliver audio at that rate and every chunk of
        512*rate/16000 samples is resampled to one 16 kHz model frame on the GPU (frame_len and hop
        are then fixed to that chunk size).  `source_rate=None`: every stream has its own rate
        (`VADConfig.sample_rate` at open_stream, any of 8000 / 16000 / 24000 / 48000) and all of them
        advance in ONE engine step (cvad_step_args.src_rates), the way a websocket server meets them."""
        self.mixed = source_rate is None
        if self.mixed:
            source_rate, frame_len, hop = 48000, 1536, 1536        # arena sizing; per-stream values in _n_in
        if source_rate not in (8000, 16000, 24000, 48000):
            raise ConfigurationError("source_rate", str(source_rate))
        self.source_rate = int(source_rate)
        if source_rate != 16000:
            frame_len = hop = source_rate * 512 // 16000
        if not (1 <= frame_len <= 2048):
            raise ConfigurationError("frame_len", str(frame_len))
        self.frame_len = int(frame_len)
        self.hop = int(hop) if hop is not None else int(frame_len)
        if self.hop < 1:
            raise ConfigurationError("hop", str(hop))
        self.pcm_format = pcm_format
        self._dtype = np.float32 if pcm_format == capi.PCM_F32 else np.int16
        from ..engine.pool import default_device
        try:
            self._engine = StreamEngine(model_version.value, max_streams=max_streams,
                                        device=default_device() if device is None else device)
        except Exception as exc:
            raise VADError(f"Failed to initialize VAD processor: {exc}")
        self.max_streams = max_streams
        self._free = list(range(max_streams - 1, -1, -1))
        self._streams: Dict[int, _Stream] = {}
        self._lock = threading.Lock()
        # pending audio lives in one arena [slot][sample] with a fill count per slot, so that a step
        # gathers / compacts all streams with a handful of numpy calls instead of a Python loop
        self._cap = self.frame_len + 8 * max(self.hop, self.frame_len)
        self._buf = np.zeros((max_streams, self._cap), self._dtype)
        self._fill = np.zeros(max_streams, np.int64)
        self._open = np.zeros(max_streams, bool)
        self._has_cb = np.zeros(max_streams, bool)
        self._active = np.zeros(max_streams, bool)
        self._n_in = np.full(max_streams, self.frame_len, np.int64)   # source samples per model frame, per stream
        self._rate = np.full(max_streams, self.source_rate, np.int32)

    # ------------------------------------------------------------------ stream lifecycle
    @property
    def engine(self) -> StreamEngine:
        return self._engine

    def open_stream(self, config: Optional[VADConfig] = None, on_voice_start: Optional[Callable[[], None]] = None,
                    on_voice_end: Optional[Callable[[bytes], None]] = None,
                    on_voice_continue: Optional[Callable[[bytes], None]] = None) -> int:
        cfg = config if config is not None else VADConfig()
        if self.mixed:
            if int(cfg.sample_rate) not in (8000, 16000, 24000, 48000):
                raise ConfigurationError("sample_rate", str(cfg.sample_rate), "8000, 16000, 24000 or 48000")
        elif int(cfg.sample_rate) != self.source_rate:
            raise ConfigurationError("sample_rate", str(cfg.sample_rate),
                                     f"this manager takes {self.source_rate} Hz streams")
        with self._lock:
            if not self._free:
                raise VADError(f"all {self.max_streams} stream slots are in use")
            sid = self._free.pop()
            self._engine.reset([sid])
            self._apply_config(sid, cfg)
            self._streams[sid] = _Stream(
                config=cfg, writer=WAVWriter(cfg.output_wav_sample_rate, cfg.output_wav_bit_depth, 1),
                on_start=on_voice_start, on_end=on_voice_end, on_continue=on_voice_continue)
            self._fill[sid] = 0
            self._active[sid] = False
            if self.mixed:
                self._rate[sid] = int(cfg.sample_rate)
                self._n_in[sid] = int(cfg.sample_rate) * 512 // 16000
            self._open[sid] = True
            self._has_cb[sid] = any(cb is not None for cb in (on_voice_start, on_voice_end, on_voice_continue))
            return sid

    def _apply_config(self, sid: int, cfg: VADConfig) -> None:
        self._engine.configure([sid], vad_start_probability=cfg.vad_start_probability,
                               vad_end_probability=cfg.vad_end_probability,
                               voice_start_frame_count=cfg.voice_start_frame_count,
                               voice_end_frame_count=cfg.voice_end_frame_count,
                               enable_denoising=cfg.enable_denoising)

    def close_stream(self, stream_id: int) -> None:
        with self._lock:
            if self._streams.pop(stream_id, None) is not None:
                self._engine.reset([stream_id])
                self._open[stream_id] = False
                self._fill[stream_id] = 0
                self._free.append(stream_id)

    def reset_stream(self, stream_id: int) -> None:
        """VADProcessor.reset (silero_model.py:951-968) for one stream; buffered samples are dropped."""
        with self._lock:
            st = self._stream(stream_id)
            self._engine.reset([stream_id])
            self._fill[stream_id] = 0
            self._active[stream_id] = False
            st.active = False
            st.pre_roll.clear()
            st.segment.clear()

    def update_config(self, stream_id: int, config: VADConfig) -> None:
        with self._lock:
            st = self._stream(stream_id)
            st.config = config
            st.writer = WAVWriter(config.output_wav_sample_rate, config.output_wav_bit_depth, 1)
            self._apply_config(stream_id, config)
        self.reset_stream(stream_id)

    def _stream(self, stream_id: int) -> _Stream:
        try:
            return self._streams[stream_id]
        except KeyError:
            raise VADError(f"stream {stream_id} is not open")

    def is_voice_active(self, stream_id: int) -> bool:
        self._stream(stream_id)
        return bool(self._active[stream_id])

    @property
    def open_streams(self) -> List[int]:
        return sorted(self._streams)

    # ------------------------------------------------------------------ data path
    def _check_samples(self, samples) -> np.ndarray:
        x = np.asarray(samples)
        if x.size == 0:
            raise AudioProcessingError("Audio data is empty")
        if self._dtype == np.float32:
            x = x.astype(np.float32, copy=False)
            if not np.isfinite(x).all():
                raise AudioProcessingError("Audio data contains infinite or NaN values")
        else:
            x = x.astype(np.int16, copy=False)
        return x

    def _ensure_cap(self, need: int) -> None:
        if need <= self._cap:
            return
        cap = max(need, 2 * self._cap)
        buf = np.zeros((self.max_streams, cap), self._dtype)
        buf[:, :self._cap] = self._buf
        self._buf, self._cap = buf, cap

    def push(self, stream_id: int, samples) -> None:
        """Append audio to a stream's buffer (no GPU work).  NaN/Inf is rejected here, before any
        state changes, as `AudioUtils.validate_audio_data` does (audio.py:227-228)."""
        self._stream(stream_id)
        x = self._check_samples(samples)
        if x.ndim == 2:
            x = x.mean(axis=1).astype(self._dtype)
        f = int(self._fill[stream_id])
        self._ensure_cap(f + x.size)
        self._buf[stream_id, f:f + x.size] = x
        self._fill[stream_id] = f + x.size

    def push_many(self, stream_ids: Sequence[int], block: np.ndarray) -> None:
        """Lock-step producers: `block[k]` is appended to stream `stream_ids[k]` (all rows the same length)."""
        ids = np.asarray(stream_ids, np.int64)
        x = self._check_samples(block)
        if x.ndim != 2 or x.shape[0] != ids.size:
            raise AudioProcessingError("push_many expects block[len(stream_ids), samples]")
        if not self._open[ids].all():
            raise VADError("push_many: a stream is not open")
        m = x.shape[1]
        fills = self._fill[ids]
        self._ensure_cap(int(fills.max()) + m)
        f0 = int(fills[0])
        if (fills == f0).all():
            self._buf[ids, f0:f0 + m] = x
        else:
            cols = fills[:, None] + np.arange(m)[None, :]
            self._buf[ids[:, None], cols] = x
        self._fill[ids] = fills + m

    def _gate(self, st: _Stream, frame: np.ndarray) -> np.ndarray:
        f = frame.astype(np.float32)
        if self.pcm_format == capi.PCM_S16_32767:
            f = f / np.float32(32767.0)
        elif self.pcm_format == capi.PCM_S16_32768:
            f = f / np.float32(32768.0)
        rate = int(st.config.sample_rate) if self.mixed else self.source_rate
        if rate != 16000:
            # callback payloads are 16 kHz audio: same operator as the GPU resampler, host side,
            # only for the frames that actually reach a payload
            from ..utils.audio import AudioUtils
            f = AudioUtils.resample_audio(f, rate, 16000)
        if st.config.enable_denoising:
            f = np.where(np.abs(f) > 0.01, f, 0.0).astype(np.float32)
        return f

    def step(self) -> StepOutput:
        """Run every complete buffered frame of every open stream in one GPU step."""
        with self._lock:
            fill = self._fill
            if self.mixed:
                return self._step_mixed()
            counts_all = np.where(self._open & (fill >= self.frame_len), (fill - self.frame_len) // self.hop + 1, 0)
            ids = np.flatnonzero(counts_all)
            if ids.size == 0:
                z = np.zeros(0, np.int64)
                return StepOutput([], z, z, np.zeros((0, 0), np.float32), np.zeros((0, 0), np.uint8))
            counts = counts_all[ids]
            tmax = int(counts.max())
            row = (tmax - 1) * self.hop + self.frame_len
            row += (-row) % 4
            self._ensure_cap(row)
            block = self._buf[ids, :row]                      # one gather: [n, row], contiguous
            r = self._engine.step(block, slots=ids.astype(np.int32), n_frames=counts.astype(np.int32),
                                  max_frames=tmax, frame_len=self.frame_len, hop=self.hop,
                                  pcm_format=self.pcm_format, src_rate=self.source_rate)
            if r.status.any():
                self._fill[ids[r.status != 0]] = 0
                raise AudioProcessingError("Audio data contains infinite or NaN values")
            events = [StreamEvent(slot, "start" if kind == capi.FLAG_STARTED else "end", stream_frame, j)
                      for (k, slot, j, kind, stream_frame) in r.events]
            self._deliver(ids, counts, block, r.probs, r.flags)
            # compact: drop the consumed hop-multiples, keep the tail (vectorised per distinct amount)
            used = counts * self.hop
            rem = fill[ids] - used
            keep = rem > 0
            if keep.any():
                for u in np.unique(used[keep]):
                    sel = ids[keep & (used == u)]
                    w = int((fill[sel] - u).max())
                    self._buf[sel, :w] = self._buf[sel, int(u):int(u) + w]
            self._fill[ids] = rem
            return StepOutput(events, ids, counts, r.probs, r.flags)

    def _step_mixed(self) -> StepOutput:
        """step() for per-stream source rates: stream i consumes whole chunks of n_in[i] samples (hop = chunk)."""
        fill, n_in = self._fill, self._n_in
        counts_all = np.where(self._open, fill // n_in, 0)
        ids = np.flatnonzero(counts_all)
        if ids.size == 0:
            z = np.zeros(0, np.int64)
            return StepOutput([], z, z, np.zeros((0, 0), np.float32), np.zeros((0, 0), np.uint8))
        counts = counts_all[ids]
        tmax = int(counts.max())
        row = tmax * int(n_in[ids].max())
        row += (-row) % 4
        self._ensure_cap(row)
        block = self._buf[ids, :row]
        r = self._engine.step(block, slots=ids.astype(np.int32), n_frames=counts.astype(np.int32), max_frames=tmax,
                              pcm_format=self.pcm_format, src_rates=self._rate[ids])
        if r.status.any():
            self._fill[ids[r.status != 0]] = 0
            raise AudioProcessingError("Audio data contains infinite or NaN values")
        events = [StreamEvent(slot, "start" if kind == capi.FLAG_STARTED else "end", stream_frame, j)
                  for (k, slot, j, kind, stream_frame) in r.events]
        self._deliver(ids, counts, block, r.probs, r.flags)
        used = counts * n_in[ids]
        rem = fill[ids] - used
        keep = rem > 0
        if keep.any():
            for u in np.unique(used[keep]):
                sel = ids[keep & (used == u)]
                w = int((fill[sel] - u).max())
                self._buf[sel, :w] = self._buf[sel, int(u):int(u) + w]
        self._fill[ids] = rem
        return StepOutput(events, ids, counts, r.probs, r.flags)

    def _deliver(self, ids: np.ndarray, counts: np.ndarray, block: np.ndarray, probs: np.ndarray,
                 flags: np.ndarray) -> None:
        """Host side of the callbacks: segment assembly from the device's per-frame flags.  Streams
        without callbacks only get their `active` mirror refreshed, and only when something happened."""
        # every stream: voice-active mirror from the flags of its last frame (ended -> off, started or
        # continuing -> on), one vectorised pass
        last = flags[np.arange(ids.size), counts - 1]
        self._active[ids] = ((last & 2) == 0) & ((last & 5) != 0)
        # streams with callbacks: replay their frames on the host to assemble payloads
        for k in np.flatnonzero(self._has_cb[ids]):
            sid = int(ids[k])
            st = self._streams[sid]
            fl_row = flags[k]
            n_k = int(counts[k])
            start_p = st.config.vad_start_probability
            pending = block[k]
            for j in range(n_k):
                fl = int(fl_row[j])
                frame = None
                if st.wants_audio:
                    if st.active or float(probs[k, j]) >= start_p:
                        if self.mixed:
                            ni = int(self._n_in[sid])
                            frame = self._gate(st, pending[j * ni:(j + 1) * ni])
                        else:
                            frame = self._gate(st, pending[j * self.hop: j * self.hop + self.frame_len])
                if not st.active:
                    if float(probs[k, j]) >= start_p:
                        if frame is not None:
                            st.pre_roll.append(frame)
                        if fl & 1:
                            st.segment = st.pre_roll
                            st.pre_roll = []
                            st.active = True
                            self._call(st.on_start, "voice_start")
                    else:
                        st.pre_roll = []
                else:
                    if frame is not None:
                        st.segment.append(frame)
                    wav = None
                    if fl & 2:
                        if st.on_end is not None and st.segment:
                            wav = st.writer.write_wav_data(np.concatenate(st.segment))
                        st.segment = []
                        st.active = False
                    if wav:
                        self._call(st.on_end, "voice_end", wav)
                    if frame is not None:
                        self._call(st.on_continue, "voice_continue", frame.tobytes())

    @staticmethod
    def _call(cb: Optional[Callable], name: str, *args) -> None:
        if cb is None:
            return
        try:
            cb(*args)
        except Exception as exc:
            raise CallbackError(name, exc)

    # ------------------------------------------------------------------ lock-step block API
    def step_block(self, audio: np.ndarray, stream_ids: Optional[Sequence[int]] = None):
        """Lock-step fast path: `audio[k]` holds whole frames for stream `stream_ids[k]` (no buffering,
        no callbacks).  -> engine StepResult (probs, flags, status, events)."""
        ids = list(self.open_streams) if stream_ids is None else list(stream_ids)
        if self.mixed:
            raise ConfigurationError("source_rate", "None", "step_block needs one source rate for all streams")
        with self._lock:
            return self._engine.step(audio, slots=ids, frame_len=self.frame_len, hop=self.hop,
                                     pcm_format=self.pcm_format, src_rate=self.source_rate)

    def close(self) -> None:
        with self._lock:
            self._streams.clear()
            self._open[:] = False
            self._engine.close()

    def __enter__(self) -> "BatchedVADManager":
        return self

    def __exit__(self, *exc) -> None:
        self.close()
