"""In-memory RIFF/WAVE encoder for the voice-end callback payload.

Host-side, runs once per voice segment (out of the GPU hot
path).  Output bytes are identical to the reference's `WAVWriter.write_wav_data`
(`src/real_time_vad/utils/wav_writer.py:40-99`): 44-byte PCM header, samples
`clip(x * 32767, -32768, 32767)` truncated to int16 (or the int32 analogue).
"""
from __future__ import annotations

import struct
from typing import BinaryIO

import numpy as np

from ..core.exceptions import AudioProcessingError

_SCALE = {16: (32767, -32768, 32767, np.int16), 32: (2147483647, -2147483648, 2147483647, np.int32)}


def _riff_header(sample_rate: int, bit_depth: int, channels: int, data_size: int) -> bytes:
    width = bit_depth // 8
    return b"".join((
        b"RIFF", struct.pack("<I", 36 + data_size), b"WAVE",
        b"fmt ", struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * channels * width,
                             channels * width, bit_depth),
        b"data", struct.pack("<I", data_size),
    ))


class WAVWriter:
    def __init__(self, sample_rate: int = 16000, bit_depth: int = 16, channels: int = 1) -> None:
        if bit_depth not in _SCALE:
            raise ValueError(f"Unsupported bit depth: {bit_depth}. Must be 16 or 32.")
        if channels not in (1, 2):
            raise ValueError(f"Unsupported channel count: {channels}. Must be 1 or 2.")
        self.sample_rate = sample_rate
        self.bit_depth = bit_depth
        self.channels = channels

    def _quantise(self, audio_data: np.ndarray) -> np.ndarray:
        if not isinstance(audio_data, np.ndarray):
            raise ValueError("Audio data must be a numpy array")
        x = audio_data if audio_data.dtype == np.float32 else audio_data.astype(np.float32)
        if self.channels == 1 and x.ndim > 1:
            x = np.mean(x, axis=1)
        gain, lo, hi, dtype = _SCALE[self.bit_depth]
        return np.clip(x * gain, lo, hi).astype(dtype)

    def write_wav_data(self, audio_data: np.ndarray) -> bytes:
        try:
            pcm = self._quantise(audio_data)
            body = pcm.tobytes()
            return _riff_header(self.sample_rate, self.bit_depth, self.channels,
                                len(pcm) * (self.bit_depth // 8)) + body
        except Exception as exc:
            raise AudioProcessingError(f"Failed to create WAV data: {exc}")

    def _write_wav_header(self, buffer: BinaryIO, file_size: int, data_size: int, format_code: int,
                          bytes_per_sample: int) -> None:
        buffer.write(_riff_header(self.sample_rate, bytes_per_sample * 8, self.channels, data_size))

    def write_wav_file(self, filename: str, audio_data: np.ndarray) -> None:
        try:
            payload = self.write_wav_data(audio_data)
            with open(filename, "wb") as fh:
                fh.write(payload)
        except Exception as exc:
            raise AudioProcessingError(f"Failed to write WAV file {filename}: {exc}")

    @staticmethod
    def create_wav_header(sample_rate: int, bit_depth: int, channels: int, data_size: int) -> bytes:
        return _riff_header(sample_rate, bit_depth, channels, data_size)

    @staticmethod
    def validate_wav_parameters(sample_rate: int, bit_depth: int, channels: int) -> None:
        if sample_rate <= 0:
            raise ValueError(f"Invalid sample rate: {sample_rate}")
        if bit_depth not in (16, 32):
            raise ValueError(f"Unsupported bit depth: {bit_depth}")
        if channels not in (1, 2):
            raise ValueError(f"Unsupported channel count: {channels}")

    def get_format_info(self) -> dict:
        return {"sample_rate": self.sample_rate, "bit_depth": self.bit_depth, "channels": self.channels,
                "bytes_per_sample": self.bit_depth // 8, "format": "PCM"}
