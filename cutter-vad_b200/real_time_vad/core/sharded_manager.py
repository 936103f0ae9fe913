"""ShardedVADManager: the stream -> GPU router of the multi-GPU layout.

Streams are independent, so the 65,536-stream configuration (BASELINE.json configs[3]) is eight 8,192-stream
`BatchedVADManager`s, one per process and GPU, with NO collective on the data path.  This class is the piece that makes
that an API instead of a launch recipe: every rank constructs it with the same arguments; GLOBAL stream `s` lives on rank
`s % world_size` in local slot order (`engine/sharding.py`).  Calls that name a stream act on the owning rank and are
no-ops elsewhere, so a front end may broadcast a call or route it, as it prefers:

    mgr = ShardedVADManager(total_streams=65_536)        # under torchrun: rank / world size / device from the environment
    if mgr.owns(s): mgr.open_stream(s, VADConfig(...), on_voice_start=...)
    mgr.push(s, samples)                                  # False on ranks that do not own s
    out = mgr.step()                                      # this rank's streams, ids are GLOBAL
    everybody = mgr.all_events(out)                       # optional: every rank's events, gathered (the only collective)

There is no reference counterpart: the reference runs one process with one onnxruntime session per client
(websocket_service/server/vad_websocket_server.py:248-290).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from ..engine import sharding
from .batched_manager import BatchedVADManager, StepOutput, StreamEvent
from .config import VADConfig
from .exceptions import VADError


class ShardedVADManager:
    def __init__(self, total_streams: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 device: Optional[int] = None, manager_factory: Optional[Callable[..., BatchedVADManager]] = None,
                 **manager_kw) -> None:
        """`total_streams`: capacity over ALL ranks.  rank / world_size default to the torchrun environment (RANK,
        WORLD_SIZE); `device` defaults to LOCAL_RANK modulo the visible devices.  `manager_kw` goes to the local
        `BatchedVADManager` (frame_len, hop, pcm_format, source_rate, ...)."""
        env_rank, env_world, env_local = sharding.world()
        self.rank = env_rank if rank is None else int(rank)
        self.world_size = env_world if world_size is None else int(world_size)
        if not (0 <= self.rank < self.world_size):
            raise VADError(f"rank {self.rank} outside world of {self.world_size}")
        self.total_streams = int(total_streams)
        self.local_capacity = sharding.local_capacity(self.total_streams, self.world_size, self.rank)
        if device is None:
            from ..engine import capi
            n_dev = max(1, capi.lib().cvad_device_count()) if manager_factory is None else 1
            device = env_local % n_dev
        make = manager_factory or BatchedVADManager
        self._mgr = make(max_streams=max(1, self.local_capacity), device=device, **manager_kw)
        self._local_of: Dict[int, int] = {}        # global stream id -> local manager id
        self._global_of: Dict[int, int] = {}       # local manager id -> global stream id

    # ------------------------------------------------------------------ routing
    def owner(self, stream: int) -> int:
        return sharding.owner_of(int(stream), self.world_size)[0]

    def owns(self, stream: int) -> bool:
        return 0 <= stream < self.total_streams and self.owner(stream) == self.rank

    @property
    def local(self) -> BatchedVADManager:
        return self._mgr

    @property
    def open_streams(self) -> List[int]:
        """GLOBAL ids of the streams open on this rank."""
        return sorted(self._local_of)

    # ------------------------------------------------------------------ stream lifecycle (owner acts, others ignore)
    def open_stream(self, stream: int, config: Optional[VADConfig] = None, **callbacks) -> bool:
        if not (0 <= stream < self.total_streams):
            raise VADError(f"stream id {stream} outside [0, {self.total_streams})")
        if not self.owns(stream):
            return False
        if stream in self._local_of:
            raise VADError(f"stream {stream} is already open")
        lid = self._mgr.open_stream(config, **callbacks)
        self._local_of[stream] = lid
        self._global_of[lid] = stream
        return True

    def close_stream(self, stream: int) -> bool:
        lid = self._local_of.pop(stream, None)
        if lid is None:
            return False
        self._global_of.pop(lid, None)
        self._mgr.close_stream(lid)
        return True

    def push(self, stream: int, samples) -> bool:
        lid = self._local_of.get(stream)
        if lid is None:
            if self.owns(stream):
                raise VADError(f"stream {stream} is not open")
            return False
        self._mgr.push(lid, samples)
        return True

    def push_bytes(self, stream: int, data: bytes) -> bool:
        lid = self._local_of.get(stream)
        if lid is None:
            if self.owns(stream):
                raise VADError(f"stream {stream} is not open")
            return False
        self._mgr.push_bytes(lid, data)
        return True

    def push_many(self, streams, block: np.ndarray) -> int:
        """Rows of `block` whose stream this rank owns are appended; returns how many that were."""
        streams = np.asarray(streams, np.int64)
        mine = np.flatnonzero(streams % self.world_size == self.rank)
        if mine.size == 0:
            return 0
        try:
            lids = np.array([self._local_of[int(s)] for s in streams[mine]], np.int32)
        except KeyError as exc:
            raise VADError(f"stream {exc.args[0]} is not open")
        self._mgr.push_many(lids, np.ascontiguousarray(np.asarray(block)[mine]))
        return int(mine.size)

    # ------------------------------------------------------------------ the step
    def step(self) -> StepOutput:
        """One GPU step over this rank's streams; stream ids in the result are GLOBAL."""
        out = self._mgr.step()
        g = self._global_of
        events = [StreamEvent(g[e.stream_id], e.kind, e.frame_index, e.step_frame) for e in out.events]
        ids = np.array([g[int(s)] for s in out.stream_ids], np.int64)
        res = StepOutput(events, ids, out.counts, out.probs, out.flags)
        for name in ("phase_ms", "host_ms"):
            if hasattr(out, name):
                setattr(res, name, getattr(out, name))
        return res

    def all_events(self, out: StepOutput) -> List[Tuple[int, int, str]]:
        """Every rank's (global stream, frame index, kind) of this step, stream-then-frame order.  The only collective of
        the layout, and optional: callbacks have already fired on the owning ranks."""
        import torch.distributed as dist
        mine = [(e.stream_id, e.frame_index, e.kind) for e in out.events]
        if not (dist.is_available() and dist.is_initialized()) or self.world_size == 1:
            return sorted(mine)
        bucket = [None] * self.world_size
        dist.all_gather_object(bucket, mine)
        return sorted(e for part in bucket for e in part)

    def close(self) -> None:
        self._local_of.clear()
        self._global_of.clear()
        self._mgr.close()

    def __enter__(self) -> "ShardedVADManager":
        return self

    def __exit__(self, *exc) -> None:
        self.close()
