"""Process-wide engine pool: one StreamEngine per (model version, device), slots leased
to per-stream objects (VADProcessor, compat sessions, BatchedVADManager).

The reference builds one onnxruntime session per VADWrapper
(/root/reference/src/real_time_vad/core/silero_model.py:321-325; one per websocket client,
websocket_service/server/vad_websocket_server.py:277).  Here every stream of a process
shares the weights and the launch machinery of one engine and only owns a slot of
resident state.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Tuple

from .stream_engine import StreamEngine

_POOL_LOCK = threading.Lock()


@dataclass
class PooledEngine:
    engine: StreamEngine
    lock: threading.RLock = field(default_factory=threading.RLock)
    free: List[int] = field(default_factory=list)

    def lease(self) -> Optional[int]:
        with self.lock:
            return self.free.pop() if self.free else None

    def release(self, slot: int) -> None:
        with self.lock:
            self.engine.reset([slot])
            self.free.append(slot)


_ENGINES: Dict[Tuple[str, int, str], List[PooledEngine]] = {}


def default_device() -> int:
    return int(os.environ.get("CVAD_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def pool_slots() -> int:
    return int(os.environ.get("CVAD_POOL_SLOTS", "1024"))


def lease_slot(model_version: str = "v5", device: Optional[int] = None,
               model_path: Optional[Path] = None) -> Tuple[PooledEngine, int]:
    """-> (pooled engine, slot id); creates a further engine when every slot is taken."""
    dev = default_device() if device is None else device
    key = (model_version, dev, str(model_path) if model_path else "")
    with _POOL_LOCK:
        engines = _ENGINES.setdefault(key, [])
        for pe in engines:
            slot = pe.lease()
            if slot is not None:
                return pe, slot
        n = pool_slots()
        eng = StreamEngine(model_version, max_streams=n, device=dev, model_path=model_path)
        pe = PooledEngine(eng, free=list(range(n - 1, -1, -1)))
        engines.append(pe)
        slot = pe.lease()
        assert slot is not None
        return pe, slot


def shutdown() -> None:
    with _POOL_LOCK:
        for engines in _ENGINES.values():
            for pe in engines:
                pe.engine.close()
        _ENGINES.clear()
