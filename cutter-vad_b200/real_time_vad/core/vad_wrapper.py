"""VADWrapper: the synchronous public API, unchanged for callers, engine-backed inside.

Signatures, defaults, exception classes and message prefixes follow the reference's
`src/real_time_vad/core/vad_wrapper.py` (helper models :25-199, VADWrapper :202-984).
The one structural change is in `_process_audio_frames` (:610-647 in the reference):
instead of one onnxruntime call per overlapped frame, every frame of the call goes to
the GPU in a single batched step (`VADProcessor.process_audio_batched`); the per-frame
loop over `processor.process_frame` is kept as the compatibility path for processors
that are not engine-backed (test doubles, subclasses).
"""
from __future__ import annotations

import logging
import threading
import time
import warnings
from contextlib import contextmanager
from pathlib import Path  # noqa: F401
from typing import Any, Callable, ClassVar, Dict, List, Optional, Union

import numpy as np
from pydantic import BaseModel, ConfigDict, Field, ValidationError, field_validator, model_validator

from ..utils.audio import AudioUtils
from .config import SampleRate, SileroModelVersion, VADConfig
from .exceptions import AudioProcessingError, CallbackError, ConfigurationError, VADError
from .silero_model import ProcessingResult, ProcessingStatistics, VADProcessor

VoiceStartCallback = Callable[[], None]
VoiceEndCallback = Callable[[bytes], None]
VoiceContinueCallback = Callable[[bytes], None]


class VADWrapperState(BaseModel):
    """Counters and last-error string of one wrapper."""
    model_config = ConfigDict(arbitrary_types_allowed=True, validate_assignment=True, extra="forbid", frozen=False)

    is_initialized: bool = Field(default=False)
    total_frames_processed: int = Field(default=0, ge=0)
    total_processing_time: float = Field(default=0.0, ge=0.0)
    last_error: Optional[str] = Field(default=None)

    @property
    def average_processing_time_per_frame(self) -> float:
        n = self.total_frames_processed
        return self.total_processing_time / n if n else 0.0

    def reset_statistics(self) -> None:
        self.total_frames_processed = 0
        self.total_processing_time = 0.0

    def record_error(self, error: Exception) -> None:
        self.last_error = str(error)

    def clear_error(self) -> None:
        self.last_error = None


class CallbackConfiguration(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True, validate_assignment=True, extra="forbid")

    voice_start_callback: Optional[VoiceStartCallback] = Field(default=None)
    voice_end_callback: Optional[VoiceEndCallback] = Field(default=None)
    voice_continue_callback: Optional[VoiceContinueCallback] = Field(default=None)

    @field_validator("voice_start_callback", "voice_end_callback", "voice_continue_callback")
    @classmethod
    def _must_be_callable(cls, v):
        if v is not None and not callable(v):
            raise ValueError("Callback must be a callable function")
        return v

    def has_any_callback(self) -> bool:
        return (self.voice_start_callback is not None or self.voice_end_callback is not None
                or self.voice_continue_callback is not None)


class ThresholdConfiguration(BaseModel):
    """Argument check of `set_thresholds` (defaults differ from VADConfig: end count 57)."""
    model_config = ConfigDict(validate_assignment=True, extra="forbid")

    vad_start_probability: float = Field(default=0.7, ge=0.0, le=1.0)
    vad_end_probability: float = Field(default=0.7, ge=0.0, le=1.0)
    voice_start_ratio: float = Field(default=0.8, ge=0.0, le=1.0)
    voice_end_ratio: float = Field(default=0.95, ge=0.0, le=1.0)
    voice_start_frame_count: int = Field(default=10, ge=1)
    voice_end_frame_count: int = Field(default=57, ge=1)

    @model_validator(mode="after")
    def _sane_ranges(self):
        if self.vad_start_probability < 0.1:
            raise ValueError("Start probability should be at least 0.1 for reliable detection")
        if self.vad_end_probability < 0.1:
            raise ValueError("End probability should be at least 0.1 for reliable detection")
        if self.voice_start_frame_count > 100:
            raise ValueError("Voice start frame count should not exceed 100 for responsive detection")
        if self.voice_end_frame_count > 200:
            raise ValueError("Voice end frame count should not exceed 200 for responsive detection")
        return self


class VADWrapper:
    """High-level voice activity detector for one audio stream."""

    _DEFAULT_FRAME_OVERLAP_RATIO: ClassVar[float] = 0.5
    _MAX_PROCESSING_TIME_WARNING: ClassVar[float] = 1.0

    def __init__(self, config: Optional[VADConfig] = None) -> None:
        try:
            self._config = config if config is not None else VADConfig()
            self._validate_initial_config()
            self._state = VADWrapperState()
            self._callbacks = CallbackConfiguration()
            self._lock = threading.Lock()
            self._processor: Optional[VADProcessor] = None
            self._initialize_processor()
        except ValidationError as exc:
            raise VADError(f"Invalid configuration provided: {exc}")
        except Exception as exc:
            raise VADError(f"Failed to initialize VAD wrapper: {exc}")

    def _validate_initial_config(self) -> None:
        try:
            if not isinstance(self._config, VADConfig):
                raise ValueError("Config must be a VADConfig instance")
            if self._config.buffer_size <= 0:
                raise ValueError("Buffer size must be positive")
        except ValidationError as exc:
            raise ConfigurationError("config", str(self._config), str(exc))

    def _initialize_processor(self) -> None:
        try:
            self._processor = VADProcessor(self._config)
            self._state.is_initialized = True
            self._state.clear_error()
        except Exception as exc:
            self._state.record_error(exc)
            self._state.is_initialized = False
            raise VADError(f"Failed to initialize VAD processor: {exc}")

    # ------------------------------------------------------------------ configuration
    def _swap_setting(self, field: str, value: Any, kind: type, label: str) -> None:
        with self._lock:
            try:
                if not isinstance(value, kind):
                    raise ValueError(f"Invalid {label} type: {type(value)}")
                previous = getattr(self._config, field)
                setattr(self._config, field, value)
                if previous != value and self._state.is_initialized:
                    self._initialize_processor()
            except ValidationError as exc:
                raise ConfigurationError(field, str(value), str(exc))
            except Exception as exc:
                self._state.record_error(exc)
                raise ConfigurationError(field, str(value), str(exc))

    def set_sample_rate(self, sample_rate: SampleRate) -> None:
        self._swap_setting("sample_rate", sample_rate, SampleRate, "sample rate")

    def set_silero_model(self, model_version: SileroModelVersion) -> None:
        self._swap_setting("model_version", model_version, SileroModelVersion, "model version")

    def set_thresholds(self, vad_start_probability: float = 0.7, vad_end_probability: float = 0.7,
                       voice_start_ratio: float = 0.8, voice_end_ratio: float = 0.95,
                       voice_start_frame_count: int = 10, voice_end_frame_count: int = 57) -> None:
        with self._lock:
            try:
                checked = ThresholdConfiguration(
                    vad_start_probability=vad_start_probability, vad_end_probability=vad_end_probability,
                    voice_start_ratio=voice_start_ratio, voice_end_ratio=voice_end_ratio,
                    voice_start_frame_count=voice_start_frame_count, voice_end_frame_count=voice_end_frame_count)
                for name in ThresholdConfiguration.model_fields:
                    setattr(self._config, name, getattr(checked, name))
                if self._processor:
                    self._processor.reset()
            except ValidationError as exc:
                raise ConfigurationError("thresholds", "multiple", str(exc))
            except Exception as exc:
                self._state.record_error(exc)
                raise ConfigurationError("thresholds", "multiple", str(exc))

    # ------------------------------------------------------------------ callbacks
    def set_callbacks(self, voice_start_callback: Optional[VoiceStartCallback] = None,
                      voice_end_callback: Optional[VoiceEndCallback] = None,
                      voice_continue_callback: Optional[VoiceContinueCallback] = None) -> None:
        try:
            self._callbacks = CallbackConfiguration(voice_start_callback=voice_start_callback,
                                                    voice_end_callback=voice_end_callback,
                                                    voice_continue_callback=voice_continue_callback)
        except ValidationError as exc:
            raise VADError(f"Invalid callback configuration: {exc}")

    def _execute_callback_safely(self, callback: Optional[Callable], callback_name: str, *args, **kwargs) -> None:
        if callback is None:
            return
        try:
            callback(*args, **kwargs)
        except Exception as exc:
            self._state.record_error(exc)
            raise CallbackError(callback_name, exc)

    def _handle_callbacks(self, result: ProcessingResult) -> None:
        """start, then end (only with WAV bytes), then continue -- the reference's order (:498-519)."""
        try:
            if not isinstance(result, ProcessingResult):
                raise ValueError("Invalid processing result type")
            cb = self._callbacks
            if result.voice_started:
                logging.info("VAD: Voice started detected!")
                self._execute_callback_safely(cb.voice_start_callback, "voice_start")
            if result.voice_ended and result.wav_data:
                logging.info("VAD: Voice ended detected!")
                self._execute_callback_safely(cb.voice_end_callback, "voice_end", result.wav_data)
            if result.voice_continuing and result.pcm_data:
                self._execute_callback_safely(cb.voice_continue_callback, "voice_continue", result.pcm_data)
        except ValidationError as exc:
            raise CallbackError("result_validation", exc)

    # ------------------------------------------------------------------ audio
    @contextmanager
    def _processing_context(self):
        if not self._state.is_initialized or self._processor is None:
            raise VADError("VAD processor not initialized")
        began = time.time()
        try:
            yield
        finally:
            spent = time.time() - began
            self._state.total_processing_time += spent
            if spent > self._MAX_PROCESSING_TIME_WARNING:
                warnings.warn(f"Audio processing took {spent:.3f}s, which may indicate performance issues")

    def process_audio_data(self, audio_data: Union[np.ndarray, List[float]], input_sample_rate: Optional[int] = None) -> None:
        """The reference's call (vad_wrapper.py:547).  `input_sample_rate` is this package's opt-in to the GPU resampler:
        8000 / 24000 / 48000 = `audio_data` is at that rate and is resampled to 16 kHz in chunks of one model frame on the
        device (what the reference leaves as a placeholder under `auto_convert_sample_rate`, :619-624); None or 16000 =
        the reference's behaviour, overlapped 16 kHz frames."""
        with self._lock:
            with self._processing_context():
                try:
                    self._process_audio_frames(self._validate_and_prepare_audio(audio_data), input_sample_rate)
                except ValidationError as exc:
                    self._state.record_error(exc)
                    raise AudioProcessingError(f"Audio validation failed: {exc}")
                except Exception as exc:
                    self._state.record_error(exc)
                    raise AudioProcessingError(f"Audio processing failed: {exc}")

    def _validate_and_prepare_audio(self, audio_data: Union[np.ndarray, List[float]]) -> np.ndarray:
        if isinstance(audio_data, list):
            if not audio_data:
                raise AudioProcessingError("Audio data cannot be empty")
            audio = np.array(audio_data, dtype=np.float32)
        elif isinstance(audio_data, np.ndarray):
            audio = audio_data.astype(np.float32)
        else:
            raise AudioProcessingError(f"Unsupported audio data type: {type(audio_data)}")
        AudioUtils.validate_audio_data(audio)
        return AudioUtils.convert_to_mono(audio)

    def _process_audio_frames(self, audio_data: np.ndarray, input_sample_rate: Optional[int] = None) -> None:
        try:
            frame_size = self._config.buffer_size
            hop_size = int(frame_size * self._DEFAULT_FRAME_OVERLAP_RATIO)
            proc = self._processor
            batched = getattr(proc, "supports_batched", None)
            if input_sample_rate not in (None, 16000):
                if input_sample_rate not in (8000, 24000, 48000):
                    raise AudioProcessingError("input_sample_rate must be 8000, 16000, 24000 or 48000")
                if not (callable(batched) and batched() is True):
                    raise AudioProcessingError("input_sample_rate needs the engine-backed 16 kHz model")
                chunk = 512 * input_sample_rate // 16000       # one model frame of source samples, no overlap
                results = proc.process_audio_batched(audio_data, chunk, chunk, src_rate=input_sample_rate)
                try:
                    for result in results:
                        self._handle_callbacks(result)
                        self._state.__dict__["total_frames_processed"] += 1
                finally:
                    results.close()
                return
            if callable(batched) and batched() is True:  # `is True`: a test double answers with a Mock
                # one GPU step per batch of frames.  The generator is closed HERE, deterministically: its `finally`
                # puts the stream back to the frame a raising callback stopped at, and must not wait for the garbage
                # collector (the exception's traceback would keep it alive past the next call)
                results = proc.process_audio_batched(audio_data, frame_size, hop_size)
                try:
                    for result in results:
                        self._handle_callbacks(result)
                        self._state.__dict__["total_frames_processed"] += 1
                finally:
                    results.close()
                return
            # compatibility path: one model call per frame
            for frame in AudioUtils.split_into_frames(audio_data, frame_size, hop_size):
                if len(frame) < frame_size:
                    frame = np.pad(frame, (0, frame_size - len(frame)))
                self._handle_callbacks(proc.process_frame(frame))
                self._state.total_frames_processed += 1
        except Exception as exc:
            raise AudioProcessingError(f"Frame processing failed: {exc}")

    def process_audio_data_with_buffer(self, audio_buffer: np.ndarray, count: int) -> None:
        try:
            if not isinstance(audio_buffer, np.ndarray):
                raise AudioProcessingError("Audio buffer must be a numpy array")
            if count < 0:
                raise AudioProcessingError("Count must be non-negative")
            if count > len(audio_buffer):
                raise AudioProcessingError(f"Count {count} exceeds buffer size {len(audio_buffer)}")
            self.process_audio_data(audio_buffer[:count])
        except (AudioProcessingError, ValidationError):
            raise
        except Exception as exc:
            raise AudioProcessingError(f"Buffer processing failed: {exc}")

    # ------------------------------------------------------------------ accessors
    @property
    def processor(self) -> Optional[VADProcessor]:
        return self._processor

    @property
    def config(self) -> VADConfig:
        return self._config

    @config.setter
    def config(self, value: VADConfig) -> None:
        self.update_config(value)

    # ------------------------------------------------------------------ state
    def reset(self) -> None:
        with self._lock:
            try:
                if self._processor:
                    self._processor.reset()
                self._state.reset_statistics()
                self._state.clear_error()
            except Exception as exc:
                self._state.record_error(exc)
                raise VADError(f"Failed to reset VAD state: {exc}")

    def cleanup(self) -> None:
        with self._lock:
            try:
                self._processor = None
                self._state.is_initialized = False
                self._state.clear_error()
            except Exception as exc:
                self._state.record_error(exc)

    # ------------------------------------------------------------------ information
    def get_statistics(self) -> Dict[str, Any]:
        with self._lock:
            st = self._state
            try:
                stats: Dict[str, Any] = {
                    "total_frames_processed": st.total_frames_processed,
                    "total_processing_time": st.total_processing_time,
                    "average_processing_time_per_frame": st.average_processing_time_per_frame,
                    "is_initialized": st.is_initialized,
                    "last_error": st.last_error,
                    "has_callbacks": self._callbacks.has_any_callback(),
                    "config": self._serialize_config_for_json(),
                }
                if self._processor:
                    ps = self._processor.get_statistics()
                    stats.update(ps.model_dump() if isinstance(ps, ProcessingStatistics) else ps)
                return stats
            except Exception as exc:
                st.record_error(exc)
                return {"error": str(exc), "is_initialized": st.is_initialized,
                        "total_frames_processed": st.total_frames_processed}

    def _serialize_config_for_json(self) -> Dict[str, Any]:
        try:
            data = self._config.model_dump()
            mv = data.get("model_version")
            if mv is not None:
                data["model_version"] = mv.value if hasattr(mv, "value") else str(mv)
            if data.get("model_path") is not None:
                data["model_path"] = str(data["model_path"])
            return data
        except Exception as exc:
            sr, mv = self._config.sample_rate, self._config.model_version
            return {"sample_rate": sr.value if hasattr(sr, "value") else int(sr),
                    "model_version": mv.value if hasattr(mv, "value") else str(mv),
                    "error": f"Serialization error: {exc}"}

    def get_config(self) -> VADConfig:
        return self._config

    def update_config(self, config: VADConfig) -> None:
        with self._lock:
            previous = self._config
            try:
                if not isinstance(config, VADConfig):
                    raise ValueError("Config must be a VADConfig instance")
                config.model_validate(config.model_dump())
                if self._processor:
                    self._processor.update_config(config)
                else:
                    self._config = config
                    self._initialize_processor()
                self._config = config
                self._state.clear_error()
            except ValidationError as exc:
                self._state.record_error(exc)
                raise VADError(f"Invalid configuration: {exc}")
            except Exception as exc:
                self._config = previous
                self._state.record_error(exc)
                raise VADError(f"Failed to update configuration: {exc}")

    def is_voice_active(self) -> bool:
        try:
            return self._processor.is_voice_active if self._processor else False
        except Exception as exc:
            self._state.record_error(exc)
            return False

    def get_last_error(self) -> Optional[str]:
        return self._state.last_error

    def get_last_error_details(self) -> Dict[str, Any]:
        return {"last_error": self._state.last_error, "is_initialized": self._state.is_initialized,
                "total_frames_processed": self._state.total_frames_processed,
                "has_processor": self._processor is not None}

    # ------------------------------------------------------------------ protocol support
    def __enter__(self) -> "VADWrapper":
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        try:
            self.cleanup()
        except Exception:
            pass

    def __del__(self) -> None:
        try:
            self.cleanup()
        except Exception:
            pass

    def __repr__(self) -> str:
        return (f"VADWrapper(initialized={self._state.is_initialized}, sample_rate={self._config.sample_rate}, "
                f"model_version={self._config.model_version}, "
                f"frames_processed={self._state.total_frames_processed})")

    def __str__(self) -> str:
        status = "Initialized" if self._state.is_initialized else "Not Initialized"
        return (f"VAD Wrapper - {status}\n"
                f"Sample Rate: {self._config.sample_rate.value} Hz\n"
                f"Model Version: {self._config.model_version.value}\n"
                f"Frames Processed: {self._state.total_frames_processed}\n"
                f"Has Callbacks: {self._callbacks.has_any_callback()}")
