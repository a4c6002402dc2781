"""Admission control for queries arriving from other peers.

Contract (SURVEY §2.1 p2p/ "load guard"; reference infomesh/p2p/load_guard.py): a node accepts at most 30 queries in any
60-second window and runs at most 5 at once; a query over either limit is refused with ``{"status": "OVERLOADED",
"retry_after_ms": 5000, ...}``; accepted / rejected totals and a per-peer tally (bounded to 10 000 peers) are kept.
Here the router consults the guard for every inbound SEARCH_REQUEST (in the reference the class exists but is not wired
in).

Implementation: the two limits are two small admission objects -- a ``_SlidingWindow`` of monotonic timestamps and a
``_Slots`` counter -- and the guard admits a query when both say yes; the public stats record is rebuilt from them on
demand, so there is no second copy of the counters to keep in sync."""
from __future__ import annotations

import threading
import time
from collections import Counter, deque
from dataclasses import dataclass

MAX_QUERIES_PER_MINUTE = 30
MAX_CONCURRENT_QUERIES = 5
OVERLOAD_RETRY_MS = 5000
_MAX_TRACKED_PEERS = 10_000
_WINDOW_SECONDS = 60.0


@dataclass
class LoadGuardStats:
    accepted: int = 0
    rejected: int = 0
    concurrent: int = 0
    queries_this_minute: int = 0
    is_overloaded: bool = False


class _SlidingWindow:
    """How many events happened in the last ``span`` seconds."""

    def __init__(self, limit: int, span: float = _WINDOW_SECONDS):
        self.limit, self._span = limit, span
        self._marks: deque[float] = deque()

    def count(self) -> int:
        horizon = time.monotonic() - self._span
        while self._marks and self._marks[0] < horizon:
            self._marks.popleft()
        return len(self._marks)

    def has_room(self) -> bool:
        return self.count() < self.limit

    def mark(self) -> None:
        self._marks.append(time.monotonic())

    def clear(self) -> None:
        self._marks.clear()


class _Slots:
    """Concurrently running queries."""

    def __init__(self, limit: int):
        self.limit, self.busy = limit, 0

    def has_room(self) -> bool:
        return self.busy < self.limit

    def take(self) -> None:
        self.busy += 1

    def give_back(self) -> None:
        self.busy = max(self.busy - 1, 0)


class NodeLoadGuard:
    def __init__(self, max_queries_per_minute: int = MAX_QUERIES_PER_MINUTE, max_concurrent: int = MAX_CONCURRENT_QUERIES):
        self._mutex = threading.RLock()
        self._window = _SlidingWindow(max_queries_per_minute)
        self._slots = _Slots(max_concurrent)
        self._outcomes: Counter[str] = Counter()
        self._by_peer: Counter[str] = Counter()

    def _saturated(self) -> bool:
        return not (self._window.has_room() and self._slots.has_room())

    # ---- admission
    def try_acquire(self, peer_id: str = "") -> bool:
        with self._mutex:
            if self._saturated():
                self._outcomes["rejected"] += 1
                return False
            self._window.mark()
            self._slots.take()
            self._outcomes["accepted"] += 1
            if peer_id in self._by_peer or len(self._by_peer) < _MAX_TRACKED_PEERS:
                self._by_peer[peer_id] += 1
            return True

    def release(self, peer_id: str = "") -> None:
        with self._mutex:
            self._slots.give_back()

    # ---- reporting
    @property
    def is_overloaded(self) -> bool:
        with self._mutex:
            return self._saturated()

    @property
    def stats(self) -> LoadGuardStats:
        with self._mutex:
            return LoadGuardStats(accepted=self._outcomes["accepted"], rejected=self._outcomes["rejected"], concurrent=self._slots.busy,
                                  queries_this_minute=self._window.count(), is_overloaded=self._saturated())

    def get_reject_info(self) -> dict[str, object]:
        with self._mutex:
            return {"status": "OVERLOADED", "retry_after_ms": OVERLOAD_RETRY_MS, "concurrent": self._slots.busy,
                    "qpm": self._window.count()}

    def peer_query_count(self, peer_id: str) -> int:
        with self._mutex:
            return self._by_peer[peer_id] if peer_id in self._by_peer else 0

    def reset(self) -> None:
        with self._mutex:
            self._window.clear()
            self._slots.busy = 0
            self._outcomes.clear()
            self._by_peer.clear()
