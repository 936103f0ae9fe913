"""Host-side utilities of the drop-in API."""
from .audio import AudioUtils
from .wav_writer import WAVWriter

__all__ = ["AudioUtils", "WAVWriter"]
