"""VADConfig and its enums: the configuration contract of the drop-in API.

Field names, defaults, bounds and loader behaviour follow the reference's
`src/real_time_vad/core/config.py:15-269` (SampleRate :15, SileroModelVersion :23,
VADConfig :29, from_dict/from_yaml/from_env :144-214, get_model_filename :242).
The thresholds declared here are what `cvad_configure()` uploads per slot.
"""
from __future__ import annotations

import os
from enum import Enum, IntEnum
from pathlib import Path
from typing import Any, Dict, Optional, Union

import yaml
from pydantic import BaseModel, ConfigDict, Field, field_validator


class SampleRate(IntEnum):
    SAMPLERATE_8 = 8000
    SAMPLERATE_16 = 16000
    SAMPLERATE_24 = 24000
    SAMPLERATE_48 = 48000


class SileroModelVersion(Enum):
    V4 = "v4"
    V5 = "v5"


_INT_KEYS = {"sample_rate", "voice_start_frame_count", "voice_end_frame_count", "buffer_size",
             "output_wav_sample_rate", "output_wav_bit_depth"}
_FLOAT_KEYS = {"vad_start_probability", "vad_end_probability", "voice_start_ratio", "voice_end_ratio"}
_BOOL_KEYS = {"enable_denoising", "auto_convert_sample_rate"}
_ENV_SUFFIX = {
    "SAMPLE_RATE": "sample_rate", "MODEL_VERSION": "model_version", "MODEL_PATH": "model_path",
    "START_PROBABILITY": "vad_start_probability", "END_PROBABILITY": "vad_end_probability",
    "VOICE_START_RATIO": "voice_start_ratio", "VOICE_END_RATIO": "voice_end_ratio",
    "VOICE_START_FRAME_COUNT": "voice_start_frame_count", "VOICE_END_FRAME_COUNT": "voice_end_frame_count",
    "ENABLE_DENOISING": "enable_denoising", "AUTO_CONVERT_SAMPLE_RATE": "auto_convert_sample_rate",
    "BUFFER_SIZE": "buffer_size",
}


class VADConfig(BaseModel):
    """All tunables of one VAD stream."""

    model_config = ConfigDict(use_enum_values=False, validate_assignment=True, extra="forbid")

    # model selection
    sample_rate: SampleRate = Field(default=SampleRate.SAMPLERATE_16, description="Audio sample rate")
    model_version: SileroModelVersion = Field(default=SileroModelVersion.V5, description="Silero model version")
    model_path: Optional[Path] = Field(default=None, description="Directory holding the .onnx files")
    # probability thresholds
    vad_start_probability: float = Field(default=0.7, ge=0.0, le=1.0)
    vad_end_probability: float = Field(default=0.7, ge=0.0, le=1.0)
    # ratio tests of the start/end detector
    voice_start_ratio: float = Field(default=0.8, ge=0.0, le=1.0)
    voice_end_ratio: float = Field(default=0.95, ge=0.0, le=1.0)
    # consecutive-frame counts
    voice_start_frame_count: int = Field(default=10, ge=1)
    voice_end_frame_count: int = Field(default=50, ge=1)
    # processing switches
    enable_denoising: bool = Field(default=True)
    auto_convert_sample_rate: bool = Field(default=True)
    buffer_size: int = Field(default=512, ge=256, le=2048)
    # WAV emitted on voice end
    output_wav_sample_rate: int = Field(default=16000)
    output_wav_bit_depth: int = Field(default=16)

    @field_validator("model_path")
    @classmethod
    def _model_dir_must_exist(cls, v: Optional[Path]) -> Optional[Path]:
        if v is None:
            return v
        if not v.exists():
            raise ValueError(f"Model path does not exist: {v}")
        if not v.is_dir():
            raise ValueError(f"Model path must be a directory: {v}")
        return v

    # ------------------------------------------------------------------ loaders
    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "VADConfig":
        sr = config_dict.get("sample_rate")
        if isinstance(sr, str):
            config_dict["sample_rate"] = getattr(SampleRate, f"SAMPLERATE_{sr}")
        elif isinstance(sr, int) and "sample_rate" in config_dict:
            config_dict["sample_rate"] = SampleRate(sr)
        mv = config_dict.get("model_version")
        if isinstance(mv, str):
            config_dict["model_version"] = SileroModelVersion(mv.lower())
        if config_dict.get("model_path"):
            config_dict["model_path"] = Path(config_dict["model_path"])
        return cls(**config_dict)

    @classmethod
    def from_yaml(cls, yaml_path: Union[str, Path]) -> "VADConfig":
        yaml_path = Path(yaml_path)
        if not yaml_path.exists():
            raise FileNotFoundError(f"Configuration file not found: {yaml_path}")
        with open(yaml_path, "r", encoding="utf-8") as fh:
            return cls.from_dict(yaml.safe_load(fh))

    @classmethod
    def from_env(cls, prefix: str = "VAD_") -> "VADConfig":
        found: Dict[str, Any] = {}
        for suffix, key in _ENV_SUFFIX.items():
            raw = os.getenv(prefix + suffix)
            if raw is None:
                continue
            if key in _INT_KEYS:
                found[key] = int(raw)
            elif key in _FLOAT_KEYS:
                found[key] = float(raw)
            elif key in _BOOL_KEYS:
                found[key] = raw.lower() in ("true", "1", "yes", "on")
            else:
                found[key] = raw
        return cls.from_dict(found) if found else cls()

    # ------------------------------------------------------------------ dumpers
    def to_dict(self) -> Dict[str, Any]:
        return self.model_dump()

    def _to_serializable_dict(self) -> Dict[str, Any]:
        data = self.model_dump()
        if isinstance(data.get("sample_rate"), SampleRate):
            data["sample_rate"] = data["sample_rate"].value
        if isinstance(data.get("model_version"), SileroModelVersion):
            data["model_version"] = data["model_version"].value
        if data.get("model_path") is not None:
            data["model_path"] = str(data["model_path"])
        return data

    def to_yaml(self, yaml_path: Union[str, Path]) -> None:
        yaml_path = Path(yaml_path)
        yaml_path.parent.mkdir(parents=True, exist_ok=True)
        with open(yaml_path, "w", encoding="utf-8") as fh:
            yaml.dump(self._to_serializable_dict(), fh, default_flow_style=False)

    # ------------------------------------------------------------------ helpers
    def get_model_filename(self) -> str:
        return "silero_vad_v5.onnx" if self.model_version == SileroModelVersion.V5 else "silero_vad.onnx"

    def get_frame_duration_ms(self) -> float:
        return (self.buffer_size / self.sample_rate) * 1000

    def __str__(self) -> str:
        return (f"VADConfig(\tsample_rate={self.sample_rate}Hz, \tmodel={self.model_version.value}, "
                f"\tstart_prob={self.vad_start_probability}, \tend_prob={self.vad_end_probability})")

    __repr__ = __str__
