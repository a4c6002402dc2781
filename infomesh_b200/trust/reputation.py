"""LLM summary-quality reputation per peer: running mean + EMA (alpha 0.3), 7-day recent window, letter-style grades
with an UNKNOWN band below 5 samples (reference infomesh/trust/reputation.py:25-279)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from enum import StrEnum
from pathlib import Path

from infomesh_b200.db import SQLiteStore

MIN_SAMPLES = 5
EMA_ALPHA = 0.3
RECENT_WINDOW = 7 * 24 * 3600


class ReputationGrade(StrEnum):
    EXCELLENT = "excellent"
    GOOD = "good"
    ACCEPTABLE = "acceptable"
    POOR = "poor"
    UNRELIABLE = "unreliable"
    UNKNOWN = "unknown"


GRADE_THRESHOLDS: list[tuple[float, ReputationGrade]] = [
    (0.85, ReputationGrade.EXCELLENT), (0.70, ReputationGrade.GOOD), (0.50, ReputationGrade.ACCEPTABLE),
    (0.30, ReputationGrade.POOR), (0.0, ReputationGrade.UNRELIABLE)]


@dataclass(frozen=True)
class PeerReputation:
    peer_id: str
    total_ratings: int
    recent_ratings: int
    avg_quality: float
    ema_quality: float
    recent_avg: float
    grade: ReputationGrade
    last_rated: float


def _grade_from_score(score: float, total: int) -> ReputationGrade:
    if total < MIN_SAMPLES:
        return ReputationGrade.UNKNOWN
    return next((g for floor, g in GRADE_THRESHOLDS if score >= floor), ReputationGrade.UNRELIABLE)


class LLMReputationTracker(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS llm_reputation (peer_id TEXT PRIMARY KEY, total_ratings INTEGER NOT NULL DEFAULT 0,
            quality_sum REAL NOT NULL DEFAULT 0.0, ema_quality REAL NOT NULL DEFAULT 0.5,
            last_rated REAL NOT NULL DEFAULT 0);
        CREATE TABLE IF NOT EXISTS llm_quality_log (log_id INTEGER PRIMARY KEY AUTOINCREMENT, peer_id TEXT NOT NULL,
            quality REAL NOT NULL, url TEXT NOT NULL DEFAULT '', content_hash TEXT NOT NULL DEFAULT '',
            timestamp REAL NOT NULL);
        CREATE INDEX IF NOT EXISTS idx_llm_log_peer ON llm_quality_log(peer_id);
        CREATE INDEX IF NOT EXISTS idx_llm_log_ts ON llm_quality_log(timestamp);
    """

    def __init__(self, db_path: Path | str | None = None):
        super().__init__(db_path)

    def record_quality(self, peer_id: str, quality: float, *, url: str = "", content_hash: str = "",
                       now: float | None = None) -> None:
        q = min(1.0, max(0.0, quality))
        now = now or time.time()
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO llm_reputation (peer_id, last_rated) VALUES (?, ?)", (peer_id, now))
            ema = self._conn.execute("SELECT ema_quality FROM llm_reputation WHERE peer_id = ?", (peer_id,)).fetchone()[0]
            self._conn.execute("UPDATE llm_reputation SET total_ratings = total_ratings + 1, quality_sum = quality_sum + ?, "
                               "ema_quality = ?, last_rated = ? WHERE peer_id = ?",
                               (q, EMA_ALPHA * q + (1 - EMA_ALPHA) * ema, now, peer_id))
            self._conn.execute("INSERT INTO llm_quality_log (peer_id, quality, url, content_hash, timestamp) "
                               "VALUES (?, ?, ?, ?, ?)", (peer_id, q, url, content_hash, now))
            self._conn.commit()

    def _build(self, row, *, now: float) -> PeerReputation:
        pid, total, qsum, ema, last = row
        rc, ra = self._conn.execute("SELECT COUNT(*), COALESCE(AVG(quality), 0) FROM llm_quality_log WHERE peer_id = ? "
                                    "AND timestamp >= ?", (pid, now - RECENT_WINDOW)).fetchone()
        return PeerReputation(pid, total, rc, round(qsum / total if total else 0.0, 4), round(ema, 4),
                              round(ra if rc else 0.0, 4), _grade_from_score(ema, total), last)

    def get_reputation(self, peer_id: str, *, now: float | None = None) -> PeerReputation | None:
        row = self._conn.execute("SELECT peer_id, total_ratings, quality_sum, ema_quality, last_rated FROM "
                                 "llm_reputation WHERE peer_id = ?", (peer_id,)).fetchone()
        return self._build(tuple(row), now=now or time.time()) if row else None

    def get_quality_score(self, peer_id: str) -> float:
        rep = self.get_reputation(peer_id)
        return rep.ema_quality if rep else 0.5

    def list_peers(self, *, min_ratings: int = 0, grade: ReputationGrade | None = None) -> list[PeerReputation]:
        rows = self._conn.execute("SELECT peer_id, total_ratings, quality_sum, ema_quality, last_rated FROM "
                                  "llm_reputation WHERE total_ratings >= ? ORDER BY ema_quality DESC",
                                  (min_ratings,)).fetchall()
        reps = [self._build(tuple(r), now=time.time()) for r in rows]
        return [r for r in reps if grade is None or r.grade == grade]

    def top_peers(self, n: int = 10) -> list[PeerReputation]:
        """The ``n`` best peers by EMA quality among those with at least ``MIN_SAMPLES`` ratings."""
        return self.list_peers(min_ratings=MIN_SAMPLES)[:n]

    def best_peers(self, *, limit: int = 5, min_ratings: int = MIN_SAMPLES) -> list[str]:
        return [r.peer_id for r in self.list_peers(min_ratings=min_ratings)[:limit]]

    def prune_log(self, *, max_age_seconds: float = 30 * 24 * 3600.0) -> int:
        with self._lock:
            cur = self._conn.execute("DELETE FROM llm_quality_log WHERE timestamp < ?", (time.time() - max_age_seconds,))
            self._conn.commit()
        return cur.rowcount
