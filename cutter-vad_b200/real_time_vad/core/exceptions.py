"""Exception family of the drop-in API.

Same class names, constructor signatures, `error_code` strings and message texts as
the reference's `src/real_time_vad/core/exceptions.py:8-68`, because callers and the
reference's tests match on them (`pytest.raises(..., match=...)`).  The engine's C ABI
return codes become `EngineError` in `engine/stream_engine.py` and are mapped onto these classes where the mirror calls the engine
(`core/silero_model.py`, `core/batched_manager.py`).
"""
from __future__ import annotations

from typing import Optional


class VADError(Exception):
    """Root of every error raised by this package."""

    def __init__(self, message: str, error_code: Optional[str] = None) -> None:
        super().__init__(message)
        self.message = message
        self.error_code = error_code

    def __str__(self) -> str:
        return f"[{self.error_code}] {self.message}" if self.error_code else self.message


class ModelNotFoundError(VADError):
    def __init__(self, model_path: str, message: Optional[str] = None) -> None:
        super().__init__(message or f"Silero model not found at path: {model_path}", "MODEL_NOT_FOUND")
        self.model_path = model_path


class ConfigurationError(VADError):
    def __init__(self, parameter: str, value: str, message: Optional[str] = None) -> None:
        super().__init__(message or f"Invalid configuration for parameter '{parameter}': {value}",
                         "CONFIGURATION_ERROR")
        self.parameter = parameter
        self.value = value


class AudioProcessingError(VADError):
    def __init__(self, message: str, audio_data_info: Optional[str] = None) -> None:
        super().__init__(message, "AUDIO_PROCESSING_ERROR")
        self.audio_data_info = audio_data_info


class ModelInitializationError(VADError):
    def __init__(self, model_version: str, message: Optional[str] = None) -> None:
        super().__init__(message or f"Failed to initialize Silero model version: {model_version}",
                         "MODEL_INITIALIZATION_ERROR")
        self.model_version = model_version


class CallbackError(VADError):
    def __init__(self, callback_name: str, original_error: Exception) -> None:
        super().__init__(f"Error in callback '{callback_name}': {original_error}", "CALLBACK_ERROR")
        self.callback_name = callback_name
        self.original_error = original_error
