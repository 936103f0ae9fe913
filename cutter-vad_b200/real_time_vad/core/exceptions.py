"""Exception family of the drop-in API.

Same class names, constructor signatures, `error_code` strings and message texts as
the reference's `src/real_time_vad/core/exceptions.py:8-68`, because callers and the
reference's tests match on them (`pytest.raises(..., match=...)`).  The engine's C ABI
return codes become `EngineError` in `engine/stream_engine.py` and are mapped onto these classes where the mirror calls the engine
(`core/silero_model.py`, `core/batched_manager.py`).
"""
from __future__ import annotations

from typing import Any, Optional


class VADError(Exception):
    """Root of every error raised by this package."""

    def __init__(self, message: str, error_code: Optional[str] = None) -> None:
        Exception.__init__(self, message)
        self.message, self.error_code = message, error_code

    def __str__(self) -> str:
        code = self.error_code
        return self.message if not code else "[%s] %s" % (code, self.message)


class _Coded(VADError):
    """A VADError with a fixed code; `_set` stores the text and the attributes the reference's classes expose."""

    CODE = ""

    def _set(self, text: str, **attributes: Any) -> None:
        VADError.__init__(self, text, self.CODE)
        self.__dict__.update(attributes)


class ModelNotFoundError(_Coded):
    CODE = "MODEL_NOT_FOUND"

    def __init__(self, model_path: str, message: Optional[str] = None) -> None:
        self._set(message or "Silero model not found at path: " + str(model_path), model_path=model_path)


class ConfigurationError(_Coded):
    CODE = "CONFIGURATION_ERROR"

    def __init__(self, parameter: str, value: str, message: Optional[str] = None) -> None:
        self._set(message or "Invalid configuration for parameter '%s': %s" % (parameter, value),
                  parameter=parameter, value=value)


class AudioProcessingError(_Coded):
    CODE = "AUDIO_PROCESSING_ERROR"

    def __init__(self, message: str, audio_data_info: Optional[str] = None) -> None:
        self._set(message, audio_data_info=audio_data_info)


class ModelInitializationError(_Coded):
    CODE = "MODEL_INITIALIZATION_ERROR"

    def __init__(self, model_version: str, message: Optional[str] = None) -> None:
        self._set(message or "Failed to initialize Silero model version: " + str(model_version), model_version=model_version)


class CallbackError(_Coded):
    CODE = "CALLBACK_ERROR"

    def __init__(self, callback_name: str, original_error: Exception) -> None:
        self._set("Error in callback '%s': %s" % (callback_name, original_error),
                  callback_name=callback_name, original_error=original_error)
