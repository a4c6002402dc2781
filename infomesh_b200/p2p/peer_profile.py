"""What this node has learned about how quickly each peer answers.

Contract (SURVEY §2.1 p2p/ "peer profiles"; reference infomesh/p2p/peer_profile.py): per peer an exponentially weighted
mean latency (alpha 0.3, seeded by the first sample), the 95th percentile and the success rate over the last 100
interactions (failures count towards the rate but not the latency), and -- after three interactions -- a bandwidth class
(< 100 ms fast, < 500 ms medium, else slow).  Ranking orders peers by mean latency with unknown peers last and, for
diversity, lets each peer of the slower half jump ahead of its half with probability 0.2.  The per-peer timeout scales a
base timeout by ``latency / 200 ms`` and is clamped to 0.5-5 s.  Peers silent for an hour are forgotten.

Implementation: sliding windows are bounded deques; class boundaries are a bisected table; the percentile is a single
interpolating helper; pruning is triggered by a countdown rather than a modulo on a creation counter."""
from __future__ import annotations

import random
import time
from bisect import bisect_right
from collections import deque
from dataclasses import dataclass, field
from enum import StrEnum

EMA_ALPHA = 0.3
MAX_HISTORY = 100
STALE_TIMEOUT = 3600
DIVERSITY_RATIO = 0.2
REFERENCE_LATENCY_MS = 200.0
_TIMEOUT_RANGE_MS = (500.0, 5000.0)
_UNRANKED_MS = 9999.0
_CLASS_AFTER = 3                      # interactions before a class is assigned
_PRUNE_EVERY_NEW_PEERS = 500


class BandwidthClass(StrEnum):
    FAST = "fast"
    MEDIUM = "medium"
    SLOW = "slow"
    UNKNOWN = "unknown"


_CLASS_EDGES_MS = (100.0, 500.0)
_CLASS_BY_BUCKET = (BandwidthClass.FAST, BandwidthClass.MEDIUM, BandwidthClass.SLOW)


def _classify_bandwidth(avg_ms: float) -> BandwidthClass:
    return _CLASS_BY_BUCKET[bisect_right(_CLASS_EDGES_MS, avg_ms)]


def _percentile(values, pct: float) -> float:
    """Linear interpolation between closest ranks; 0.0 for no data."""
    ordered = sorted(values)
    if not ordered:
        return 0.0
    whole, frac = divmod(pct / 100.0 * (len(ordered) - 1), 1.0)
    low = ordered[int(whole)]
    return low if frac == 0.0 else low + (ordered[int(whole) + 1] - low) * frac


@dataclass
class PeerProfile:
    peer_id: str
    avg_latency_ms: float = 0.0
    p95_latency_ms: float = 0.0
    success_rate: float = 1.0
    last_seen: float = 0.0
    bandwidth_class: BandwidthClass = BandwidthClass.UNKNOWN
    total_interactions: int = 0
    _latency_history: deque = field(default_factory=lambda: deque(maxlen=MAX_HISTORY), repr=False)
    _success_history: deque = field(default_factory=lambda: deque(maxlen=MAX_HISTORY), repr=False)

    def observe(self, elapsed_ms: float, ok: bool, stamp: float) -> None:
        self.total_interactions += 1
        self.last_seen = stamp
        self._success_history.append(bool(ok))
        self.success_rate = sum(self._success_history) / len(self._success_history)
        if ok:
            seeded = self.avg_latency_ms != 0.0
            self.avg_latency_ms = self.avg_latency_ms + EMA_ALPHA * (elapsed_ms - self.avg_latency_ms) if seeded else elapsed_ms
            self._latency_history.append(elapsed_ms)
            self.p95_latency_ms = _percentile(self._latency_history, 95)
        if self.total_interactions >= _CLASS_AFTER:
            self.bandwidth_class = _classify_bandwidth(self.avg_latency_ms)

    @property
    def rank_key(self) -> float:
        return _UNRANKED_MS if self.bandwidth_class is BandwidthClass.UNKNOWN else self.avg_latency_ms


class PeerProfileTracker:
    def __init__(self, *, max_peers: int = 10_000):
        self._book: dict[str, PeerProfile] = {}
        self._capacity = max_peers
        self._until_prune = _PRUNE_EVERY_NEW_PEERS

    # ---- recording
    def _admit(self, peer_id: str) -> PeerProfile:
        self._until_prune -= 1
        if self._until_prune <= 0 or len(self._book) >= self._capacity:
            self._until_prune = _PRUNE_EVERY_NEW_PEERS
            self.prune_stale()
        profile = self._book[peer_id] = PeerProfile(peer_id)
        return profile

    def record(self, peer_id: str, elapsed_ms: float, *, success: bool = True) -> PeerProfile:
        profile = self._book.get(peer_id) or self._admit(peer_id)
        profile.observe(elapsed_ms, success, time.time())
        return profile

    # ---- lookup
    def get(self, peer_id: str) -> PeerProfile | None:
        return self._book.get(peer_id)

    def get_or_default(self, peer_id: str) -> PeerProfile:
        return self._book.get(peer_id) or PeerProfile(peer_id)

    @property
    def known_peers(self) -> int:
        return len(self._book)

    # ---- decisions
    def rank_by_latency(self, peer_ids: list[str], *, diversity: bool = True) -> list[str]:
        by_speed = sorted(peer_ids, key=lambda pid: self.get_or_default(pid).rank_key)
        if not diversity or len(by_speed) < 3:
            return by_speed
        cut = max(len(by_speed) // 2, 1)
        quick, slow = by_speed[:cut], by_speed[cut:]
        promoted = {pid for pid in slow if random.random() < DIVERSITY_RATIO}
        return quick + [p for p in slow if p in promoted] + [p for p in slow if p not in promoted]

    def adaptive_timeout(self, peer_id: str, *, base_ms: float = 2000.0) -> float:
        profile = self._book.get(peer_id)
        if profile is None or not profile.avg_latency_ms:
            return base_ms
        low, high = _TIMEOUT_RANGE_MS
        return min(high, max(low, base_ms * (profile.avg_latency_ms / REFERENCE_LATENCY_MS)))

    # ---- housekeeping
    def prune_stale(self, *, max_age: float = STALE_TIMEOUT) -> int:
        horizon = time.time() - max_age
        silent = [pid for pid, prof in self._book.items() if 0 < prof.last_seen < horizon]
        for pid in silent:
            self._book.pop(pid)
        return len(silent)

    def reset(self) -> None:
        self._book.clear()
