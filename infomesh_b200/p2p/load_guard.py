"""Inbound query admission control: 30 queries / minute and 5 concurrent per node
(reference infomesh/p2p/load_guard.py:23-178).  Unlike the reference — where it exists but is not wired into the
peer-search handler — the router here consults it for every inbound SEARCH_REQUEST."""
from __future__ import annotations

import threading
import time
from collections import deque
from dataclasses import dataclass

MAX_QUERIES_PER_MINUTE = 30
MAX_CONCURRENT_QUERIES = 5
OVERLOAD_RETRY_MS = 5000
_MAX_TRACKED_PEERS = 10_000


@dataclass
class LoadGuardStats:
    accepted: int = 0
    rejected: int = 0
    concurrent: int = 0
    queries_this_minute: int = 0
    is_overloaded: bool = False


class NodeLoadGuard:
    def __init__(self, max_queries_per_minute: int = MAX_QUERIES_PER_MINUTE,
                 max_concurrent: int = MAX_CONCURRENT_QUERIES):
        self._qpm, self._cap = max_queries_per_minute, max_concurrent
        self._running = 0
        self._stamps: deque[float] = deque()
        self._stats = LoadGuardStats()
        self._lock = threading.RLock()
        self._per_peer: dict[str, int] = {}

    def _trim(self) -> None:
        cutoff = time.monotonic() - 60.0
        while self._stamps and self._stamps[0] < cutoff:
            self._stamps.popleft()

    @property
    def is_overloaded(self) -> bool:
        with self._lock:
            self._trim()
            return self._running >= self._cap or len(self._stamps) >= self._qpm

    def try_acquire(self, peer_id: str = "") -> bool:
        with self._lock:
            self._trim()
            if len(self._stamps) >= self._qpm or self._running >= self._cap:
                self._stats.rejected += 1
                return False
            self._stamps.append(time.monotonic())
            self._running += 1
            self._stats.accepted += 1
            if peer_id in self._per_peer or len(self._per_peer) < _MAX_TRACKED_PEERS:
                self._per_peer[peer_id] = self._per_peer.get(peer_id, 0) + 1
            return True

    def release(self, peer_id: str = "") -> None:
        with self._lock:
            self._running = max(0, self._running - 1)

    @property
    def stats(self) -> LoadGuardStats:
        with self._lock:
            self._trim()
            self._stats.concurrent = self._running
            self._stats.queries_this_minute = len(self._stamps)
            self._stats.is_overloaded = self._running >= self._cap or len(self._stamps) >= self._qpm
            return self._stats

    def get_reject_info(self) -> dict[str, object]:
        with self._lock:
            return {"status": "OVERLOADED", "retry_after_ms": OVERLOAD_RETRY_MS, "concurrent": self._running,
                    "qpm": len(self._stamps)}

    def peer_query_count(self, peer_id: str) -> int:
        with self._lock:
            return self._per_peer.get(peer_id, 0)

    def reset(self) -> None:
        with self._lock:
            self._running = 0
            self._stamps.clear()
            self._per_peer.clear()
            self._stats = LoadGuardStats()
