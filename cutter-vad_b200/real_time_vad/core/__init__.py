"""Core classes of the drop-in API (same export list as the reference's core/__init__.py)."""
from .async_vad_wrapper import AsyncVADWrapper
from .batched_manager import BatchedVADManager, StreamEvent
from .sharded_manager import ShardedVADManager
from .config import SampleRate, SileroModelVersion, VADConfig
from .exceptions import (AudioProcessingError, CallbackError, ConfigurationError, ModelInitializationError,
                         ModelNotFoundError, VADError)
from .vad_wrapper import VADWrapper

__all__ = ["VADConfig", "SampleRate", "SileroModelVersion", "VADWrapper", "AsyncVADWrapper",
           "BatchedVADManager", "ShardedVADManager", "StreamEvent", "VADError", "ModelNotFoundError", "ConfigurationError",
           "AudioProcessingError", "ModelInitializationError", "CallbackError"]
