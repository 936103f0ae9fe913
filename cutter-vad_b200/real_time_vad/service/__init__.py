"""Batched websocket service seam: the reference server's per-client VADWrapper replaced by manager slots."""
from .batched_server import BatchedVADService, ClientConfig, ClientSession, create_app, create_client_config  # noqa: F401
