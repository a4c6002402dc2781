"""Per-peer latency profiles: EMA (alpha 0.3) average, p95 over the last 100 samples, success rate, bandwidth class,
latency ranking with 20 % slow-peer diversity, adaptive per-peer timeouts clamped to 500..5000 ms
(reference infomesh/p2p/peer_profile.py:27-263)."""
from __future__ import annotations

import math
import random
import time
from dataclasses import dataclass, field
from enum import StrEnum

EMA_ALPHA = 0.3
MAX_HISTORY = 100
STALE_TIMEOUT = 3600
DIVERSITY_RATIO = 0.2
REFERENCE_LATENCY_MS = 200.0


class BandwidthClass(StrEnum):
    FAST = "fast"
    MEDIUM = "medium"
    SLOW = "slow"
    UNKNOWN = "unknown"


@dataclass
class PeerProfile:
    peer_id: str
    avg_latency_ms: float = 0.0
    p95_latency_ms: float = 0.0
    success_rate: float = 1.0
    last_seen: float = 0.0
    bandwidth_class: BandwidthClass = BandwidthClass.UNKNOWN
    total_interactions: int = 0
    _latency_history: list[float] = field(default_factory=list, repr=False)
    _success_history: list[bool] = field(default_factory=list, repr=False)


def _classify_bandwidth(avg_ms: float) -> BandwidthClass:
    return BandwidthClass.FAST if avg_ms < 100 else BandwidthClass.MEDIUM if avg_ms < 500 else BandwidthClass.SLOW


def _percentile(values: list[float], pct: float) -> float:
    if not values:
        return 0.0
    s = sorted(values)
    pos = pct / 100 * (len(s) - 1)
    lo, hi = math.floor(pos), math.ceil(pos)
    return s[lo] if lo == hi else s[lo] * (hi - pos) + s[hi] * (pos - lo)


class PeerProfileTracker:
    def __init__(self, *, max_peers: int = 10_000):
        self._profiles: dict[str, PeerProfile] = {}
        self._max = max_peers
        self._new = 0

    def record(self, peer_id: str, elapsed_ms: float, *, success: bool = True) -> PeerProfile:
        p = self._profiles.get(peer_id)
        if p is None:
            self._new += 1
            if len(self._profiles) >= self._max or self._new % 500 == 0:
                self.prune_stale()
            p = self._profiles[peer_id] = PeerProfile(peer_id)
        p.total_interactions += 1
        p.last_seen = time.time()
        if success:
            p.avg_latency_ms = elapsed_ms if p.avg_latency_ms == 0.0 else EMA_ALPHA * elapsed_ms + (1 - EMA_ALPHA) * p.avg_latency_ms
            p._latency_history = (p._latency_history + [elapsed_ms])[-MAX_HISTORY:]
            p.p95_latency_ms = _percentile(p._latency_history, 95)
        p._success_history = (p._success_history + [success])[-MAX_HISTORY:]
        p.success_rate = sum(p._success_history) / len(p._success_history)
        if p.total_interactions >= 3:
            p.bandwidth_class = _classify_bandwidth(p.avg_latency_ms)
        return p

    def get(self, peer_id: str) -> PeerProfile | None:
        return self._profiles.get(peer_id)

    def get_or_default(self, peer_id: str) -> PeerProfile:
        return self._profiles.get(peer_id) or PeerProfile(peer_id)

    @property
    def known_peers(self) -> int:
        return len(self._profiles)

    def rank_by_latency(self, peer_ids: list[str], *, diversity: bool = True) -> list[str]:
        """Fast half first; each slow-half peer is promoted right behind it with probability 0.2."""
        def key(pid: str) -> float:
            p = self.get_or_default(pid)
            return 9999.0 if p.bandwidth_class == BandwidthClass.UNKNOWN else p.avg_latency_ms

        ordered = sorted(peer_ids, key=key)
        if not diversity or len(ordered) <= 2:
            return ordered
        mid = max(1, len(ordered) // 2)
        fast, slow = ordered[:mid], ordered[mid:]
        lucky = [pid for pid in slow if random.random() < DIVERSITY_RATIO]
        return fast + lucky + [pid for pid in slow if pid not in lucky]

    def adaptive_timeout(self, peer_id: str, *, base_ms: float = 2000.0) -> float:
        p = self.get(peer_id)
        if p is None or p.avg_latency_ms == 0.0:
            return base_ms
        return max(500.0, min(base_ms * p.avg_latency_ms / REFERENCE_LATENCY_MS, 5000.0))

    def prune_stale(self, *, max_age: float = STALE_TIMEOUT) -> int:
        now = time.time()
        dead = [pid for pid, p in self._profiles.items() if p.last_seen > 0 and now - p.last_seen > max_age]
        for pid in dead:
            del self._profiles[pid]
        return len(dead)

    def reset(self) -> None:
        self._profiles.clear()
