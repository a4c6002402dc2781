"""Credit-farming detection: 24 h probation for new nodes, per-action hourly caps, bot-like regular intervals
(coefficient of variation < 0.15 over >= 10 samples), bursts (>= 30 actions / 5 min); 3 anomalies => blocked
(reference infomesh/credits/farming.py:30-480)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from enum import StrEnum
from pathlib import Path

from infomesh_b200.db import SQLiteStore
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

PROBATION_HOURS: float = 24.0
PROBATION_CREDIT_MULTIPLIER: float = 0.5
MAX_CRAWLS_PER_HOUR: int = 120
MAX_QUERIES_PER_HOUR: int = 300
MAX_LLM_PER_HOUR: int = 60
MIN_INTERVAL_CV: float = 0.15
BURST_WINDOW_MINUTES: float = 5.0
BURST_THRESHOLD: int = 30
ANOMALY_FLAG_THRESHOLD: int = 3
_HOURLY_LIMITS = {"crawl": MAX_CRAWLS_PER_HOUR, "query_process": MAX_QUERIES_PER_HOUR, "llm_own": MAX_LLM_PER_HOUR,
                  "llm_peer": MAX_LLM_PER_HOUR}


class FarmingVerdict(StrEnum):
    CLEAN = "clean"
    PROBATION = "probation"
    RATE_LIMITED = "rate_limited"
    SUSPICIOUS = "suspicious"
    BLOCKED = "blocked"


@dataclass(frozen=True)
class FarmingCheck:
    peer_id: str
    verdict: FarmingVerdict
    probation_remaining_hours: float
    rate_limit_exceeded: bool
    anomaly_count: int
    detail: str


@dataclass(frozen=True)
class AnomalyEvent:
    event_id: int
    peer_id: str
    anomaly_type: str
    detail: str
    timestamp: float


class FarmingDetector(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS node_registry (peer_id TEXT PRIMARY KEY, registered_at REAL NOT NULL,
            blocked INTEGER NOT NULL DEFAULT 0, anomaly_count INTEGER NOT NULL DEFAULT 0);
        CREATE TABLE IF NOT EXISTS action_log (log_id INTEGER PRIMARY KEY AUTOINCREMENT, peer_id TEXT NOT NULL,
            action TEXT NOT NULL, timestamp REAL NOT NULL);
        CREATE TABLE IF NOT EXISTS anomaly_events (event_id INTEGER PRIMARY KEY AUTOINCREMENT, peer_id TEXT NOT NULL,
            anomaly_type TEXT NOT NULL, detail TEXT NOT NULL DEFAULT '', timestamp REAL NOT NULL);
        CREATE INDEX IF NOT EXISTS idx_action_log_peer ON action_log(peer_id, action, timestamp);
        CREATE INDEX IF NOT EXISTS idx_anomaly_peer ON anomaly_events(peer_id);
    """

    def __init__(self, db_path: Path | str | None = None):
        super().__init__(db_path)

    def _registered_at(self, peer_id: str) -> float | None:
        row = self._conn.execute("SELECT registered_at FROM node_registry WHERE peer_id = ?", (peer_id,)).fetchone()
        return float(row[0]) if row else None

    def register_node(self, peer_id: str, *, now: float | None = None) -> None:
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO node_registry (peer_id, registered_at) VALUES (?, ?)",
                               (peer_id, now or time.time()))
            self._conn.commit()

    def probation_remaining(self, peer_id: str, *, now: float | None = None) -> float:
        reg = self._registered_at(peer_id)
        if reg is None:
            return PROBATION_HOURS
        return max(0.0, PROBATION_HOURS - ((now or time.time()) - reg) / 3600.0)

    def is_on_probation(self, peer_id: str, *, now: float | None = None) -> bool:
        return self.probation_remaining(peer_id, now=now) > 0

    def log_action(self, peer_id: str, action: str, *, now: float | None = None) -> None:
        with self._lock:
            self._conn.execute("INSERT INTO action_log (peer_id, action, timestamp) VALUES (?, ?, ?)",
                               (peer_id, action, now or time.time()))
            self._conn.commit()

    def prune_old_actions(self, *, max_age_seconds: float = 7 * 24 * 3600.0) -> int:
        with self._lock:
            cur = self._conn.execute("DELETE FROM action_log WHERE timestamp < ?", (time.time() - max_age_seconds,))
            self._conn.commit()
        return cur.rowcount

    def _count_since(self, peer_id: str, action: str, cutoff: float) -> int:
        return int(self._conn.execute("SELECT COUNT(*) FROM action_log WHERE peer_id = ? AND action = ? AND "
                                      "timestamp >= ?", (peer_id, action, cutoff)).fetchone()[0])

    def actions_in_last_hour(self, peer_id: str, action: str, *, now: float | None = None) -> int:
        return self._count_since(peer_id, action, (now or time.time()) - 3600.0)

    def is_rate_limited(self, peer_id: str, action: str, *, now: float | None = None) -> bool:
        return self.actions_in_last_hour(peer_id, action, now=now) >= _HOURLY_LIMITS.get(action, MAX_CRAWLS_PER_HOUR)

    def detect_regular_intervals(self, peer_id: str, action: str, *, window_hours: float = 1.0,
                                 now: float | None = None) -> bool:
        now = now or time.time()
        ts = [r[0] for r in self._conn.execute(
            "SELECT timestamp FROM action_log WHERE peer_id = ? AND action = ? AND timestamp >= ? ORDER BY timestamp",
            (peer_id, action, now - window_hours * 3600.0))]
        if len(ts) < 10:
            return False
        gaps = [b - a for a, b in zip(ts, ts[1:])]
        mean = sum(gaps) / len(gaps)
        if mean <= 0:
            return True
        cv = (sum((g - mean) ** 2 for g in gaps) / len(gaps)) ** 0.5 / mean
        return cv < MIN_INTERVAL_CV

    def detect_burst(self, peer_id: str, action: str, *, now: float | None = None) -> bool:
        return self._count_since(peer_id, action, (now or time.time()) - BURST_WINDOW_MINUTES * 60.0) >= BURST_THRESHOLD

    def _anomaly_count(self, peer_id: str) -> int:
        row = self._conn.execute("SELECT anomaly_count FROM node_registry WHERE peer_id = ?", (peer_id,)).fetchone()
        return int(row[0]) if row else 0

    def record_anomaly(self, peer_id: str, anomaly_type: str, detail: str = "", *, now: float | None = None) -> int:
        with self._lock:
            self._conn.execute("INSERT INTO anomaly_events (peer_id, anomaly_type, detail, timestamp) VALUES (?, ?, ?, ?)",
                               (peer_id, anomaly_type, detail, now or time.time()))
            self._conn.execute("UPDATE node_registry SET anomaly_count = anomaly_count + 1 WHERE peer_id = ?",
                               (peer_id,))
            count = self._anomaly_count(peer_id) or 1
            if count >= ANOMALY_FLAG_THRESHOLD:
                self._conn.execute("UPDATE node_registry SET blocked = 1 WHERE peer_id = ?", (peer_id,))
                logger.warning("farming_node_blocked", peer_id=peer_id[:12], anomaly_count=count)
            self._conn.commit()
        return count

    def is_blocked(self, peer_id: str) -> bool:
        row = self._conn.execute("SELECT blocked FROM node_registry WHERE peer_id = ?", (peer_id,)).fetchone()
        return bool(row[0]) if row else False

    def unblock(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("UPDATE node_registry SET blocked = 0, anomaly_count = 0 WHERE peer_id = ?", (peer_id,))
            self._conn.commit()

    def check(self, peer_id: str, action: str, *, now: float | None = None) -> FarmingCheck:
        now = now or time.time()
        self.register_node(peer_id, now=now)
        if self.is_blocked(peer_id):
            return FarmingCheck(peer_id, FarmingVerdict.BLOCKED, 0.0, False, self._anomaly_count(peer_id),
                                "node blocked for credit farming")
        limited = self.is_rate_limited(peer_id, action, now=now)
        remaining = self.probation_remaining(peer_id, now=now)
        found = [name for name, hit in (("regular_intervals", self.detect_regular_intervals(peer_id, action, now=now)),
                                        ("burst", self.detect_burst(peer_id, action, now=now))) if hit]
        for name in found:
            self.record_anomaly(peer_id, name, detail=f"action={action}", now=now)
        if self.is_blocked(peer_id):
            verdict, detail = FarmingVerdict.BLOCKED, "blocked after anomaly detection"
        elif found:
            verdict, detail = FarmingVerdict.SUSPICIOUS, f"anomalies: {', '.join(found)}"
        elif limited:
            verdict, detail = FarmingVerdict.RATE_LIMITED, f"rate limit exceeded for {action}"
        elif remaining > 0:
            verdict, detail = FarmingVerdict.PROBATION, f"probation: {remaining:.1f}h remaining"
        else:
            verdict, detail = FarmingVerdict.CLEAN, "ok"
        return FarmingCheck(peer_id, verdict, round(remaining, 2), limited, self._anomaly_count(peer_id), detail)

    def get_anomaly_history(self, peer_id: str, *, limit: int = 50) -> list[AnomalyEvent]:
        rows = self._conn.execute("SELECT event_id, peer_id, anomaly_type, detail, timestamp FROM anomaly_events "
                                  "WHERE peer_id = ? ORDER BY timestamp DESC LIMIT ?", (peer_id, limit)).fetchall()
        return [AnomalyEvent(*tuple(r)) for r in rows]

    def credit_multiplier(self, peer_id: str, *, now: float | None = None) -> float:
        """0 when blocked, 0.5 during probation, else 1."""
        if self.is_blocked(peer_id):
            return 0.0
        return PROBATION_CREDIT_MULTIPLIER if self.is_on_probation(peer_id, now=now) else 1.0
