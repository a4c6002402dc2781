"""Peer trust: 0.15 uptime + 0.25 contribution + 0.40 audit pass rate + 0.20 summary quality; tiers at .8/.5/.3;
three consecutive audit failures isolate a peer (reference infomesh/trust/scoring.py:32-453)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from enum import StrEnum
from pathlib import Path
from typing import Any

from infomesh_b200.db import SQLiteStore
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

W_UPTIME = 0.15
W_CONTRIBUTION = 0.25
W_AUDIT = 0.40
W_SUMMARY = 0.20
AUDIT_FAILURE_ISOLATION_THRESHOLD = 3
MAX_UPTIME_HOURS: float = 30 * 24
MAX_CONTRIBUTION_SCORE: float = 5000.0


class TrustTier(StrEnum):
    TRUSTED = "trusted"
    NORMAL = "normal"
    SUSPECT = "suspect"
    UNTRUSTED = "untrusted"


TIER_THRESHOLDS: list[tuple[float, TrustTier]] = [(0.8, TrustTier.TRUSTED), (0.5, TrustTier.NORMAL),
                                                  (0.3, TrustTier.SUSPECT), (0.0, TrustTier.UNTRUSTED)]


@dataclass(frozen=True)
class PeerTrust:
    peer_id: str
    uptime_score: float
    contribution_score: float
    audit_pass_rate: float
    summary_quality: float
    trust_score: float
    tier: TrustTier
    consecutive_audit_failures: int
    isolated: bool
    last_updated: float


@dataclass(frozen=True)
class TrustUpdate:
    peer_id: str
    field: str
    value: float
    timestamp: float


def compute_trust_score(uptime_hours: float, contribution_raw: float, audit_total: int, audit_passed: int,
                        summary_avg: float, has_summary_data: bool | None = None) -> float:
    """Unknown signals default to the neutral 0.5 (no audits yet / no summary ratings yet)."""
    audit = audit_passed / audit_total if audit_total > 0 else 0.5
    if has_summary_data is True:
        summary = summary_avg
    elif has_summary_data is False:
        summary = 0.5
    else:
        summary = summary_avg if summary_avg > 0 else 0.5
    return (W_UPTIME * min(1.0, uptime_hours / MAX_UPTIME_HOURS)
            + W_CONTRIBUTION * min(1.0, contribution_raw / MAX_CONTRIBUTION_SCORE) + W_AUDIT * audit
            + W_SUMMARY * summary)


def trust_tier(score: float) -> TrustTier:
    return next((t for floor, t in TIER_THRESHOLDS if score >= floor), TrustTier.UNTRUSTED)


_COLS = ("peer_id, uptime_hours, contribution_raw, audit_total, audit_passed, summary_ratings_sum, "
         "summary_ratings_count, consecutive_audit_failures, isolated, last_updated")


def _compute_trust(row: tuple[Any, ...]) -> PeerTrust:
    pid, up, contrib, a_tot, a_ok, s_sum, s_cnt, fails, iso, last = row
    s_avg = s_sum / s_cnt if s_cnt > 0 else 0.0
    score = compute_trust_score(up, contrib, a_tot, a_ok, s_avg, has_summary_data=s_cnt > 0)
    # the record reports the raw component values (unrounded; a component without data reads 0.0 for summaries and the
    # neutral 0.5 for audits, as in the reference) -- the neutral defaults live in compute_trust_score
    return PeerTrust(pid, min(1.0, up / MAX_UPTIME_HOURS), min(1.0, contrib / MAX_CONTRIBUTION_SCORE),
                     a_ok / a_tot if a_tot else 0.5, s_avg if s_cnt else 0.0, score, trust_tier(score), int(fails), bool(iso), last)


class TrustStore(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS peer_trust (peer_id TEXT PRIMARY KEY, uptime_hours REAL NOT NULL DEFAULT 0,
            contribution_raw REAL NOT NULL DEFAULT 0, audit_total INTEGER NOT NULL DEFAULT 0,
            audit_passed INTEGER NOT NULL DEFAULT 0, summary_ratings_sum REAL NOT NULL DEFAULT 0,
            summary_ratings_count INTEGER NOT NULL DEFAULT 0, consecutive_audit_failures INTEGER NOT NULL DEFAULT 0,
            isolated INTEGER NOT NULL DEFAULT 0, last_updated REAL NOT NULL DEFAULT 0);
        CREATE TABLE IF NOT EXISTS trust_events (event_id INTEGER PRIMARY KEY AUTOINCREMENT, peer_id TEXT NOT NULL,
            field TEXT NOT NULL, value REAL NOT NULL, timestamp REAL NOT NULL);
        CREATE INDEX IF NOT EXISTS idx_trust_events_peer ON trust_events(peer_id);
    """

    def __init__(self, db_path: Path | str | None = None, *, reputation_tracker: Any | None = None):
        self._reputation = reputation_tracker
        super().__init__(db_path)

    def _touch(self, peer_id: str, set_sql: str, args: tuple, field: str, value: float) -> None:
        now = time.time()
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO peer_trust (peer_id, last_updated) VALUES (?, ?)", (peer_id, now))
            self._conn.execute(f"UPDATE peer_trust SET {set_sql}, last_updated = ? WHERE peer_id = ?",
                               (*args, now, peer_id))
            self._conn.execute("INSERT INTO trust_events (peer_id, field, value, timestamp) VALUES (?, ?, ?, ?)",
                               (peer_id, field, value, now))
            self._conn.commit()

    def update_uptime(self, peer_id: str, hours: float) -> None:
        self._touch(peer_id, "uptime_hours = ?", (hours,), "uptime", hours)

    def update_contribution(self, peer_id: str, score: float) -> None:
        self._touch(peer_id, "contribution_raw = ?", (score,), "contribution", score)

    def record_audit(self, peer_id: str, *, passed: bool) -> None:
        if passed:
            self._touch(peer_id, "audit_total = audit_total + 1, audit_passed = audit_passed + 1, "
                        "consecutive_audit_failures = 0", (), "audit", 1.0)
            return
        self._touch(peer_id, "audit_total = audit_total + 1, consecutive_audit_failures = consecutive_audit_failures + 1",
                    (), "audit", 0.0)
        row = self._conn.execute("SELECT consecutive_audit_failures FROM peer_trust WHERE peer_id = ?",
                                 (peer_id,)).fetchone()
        if row and row[0] >= AUDIT_FAILURE_ISOLATION_THRESHOLD:
            self.isolate_peer(peer_id)
            logger.warning("peer_isolated", peer_id=peer_id, failures=row[0])

    def record_summary_rating(self, peer_id: str, quality: float) -> None:
        q = min(1.0, max(0.0, quality))
        self._touch(peer_id, "summary_ratings_sum = summary_ratings_sum + ?, summary_ratings_count = "
                    "summary_ratings_count + 1", (q,), "summary", q)
        if self._reputation is not None and hasattr(self._reputation, "record_quality"):
            try:
                self._reputation.record_quality(peer_id, q)
            except Exception as exc:  # noqa: BLE001
                logger.debug("reputation_forward_failed", peer_id=peer_id, error=str(exc))

    def isolate_peer(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO peer_trust (peer_id, last_updated) VALUES (?, ?)",
                               (peer_id, time.time()))
            self._conn.execute("UPDATE peer_trust SET isolated = 1 WHERE peer_id = ?", (peer_id,))
            self._conn.commit()

    def unisolate(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("UPDATE peer_trust SET isolated = 0, consecutive_audit_failures = 0 WHERE peer_id = ?",
                               (peer_id,))
            self._conn.commit()

    def is_isolated(self, peer_id: str) -> bool:
        row = self._conn.execute("SELECT isolated FROM peer_trust WHERE peer_id = ?", (peer_id,)).fetchone()
        return bool(row[0]) if row else False

    def get_trust(self, peer_id: str) -> PeerTrust | None:
        row = self._conn.execute(f"SELECT {_COLS} FROM peer_trust WHERE peer_id = ?", (peer_id,)).fetchone()
        return _compute_trust(tuple(row)) if row else None

    def get_trust_score(self, peer_id: str) -> float:
        t = self.get_trust(peer_id)
        return t.trust_score if t else 0.5

    def list_peers(self, *, include_isolated: bool = False) -> list[PeerTrust]:
        where = "" if include_isolated else " WHERE isolated = 0"
        return [_compute_trust(tuple(r)) for r in self._conn.execute(f"SELECT {_COLS} FROM peer_trust{where}")]

    def list_isolated(self) -> list[PeerTrust]:
        return [_compute_trust(tuple(r)) for r in self._conn.execute(f"SELECT {_COLS} FROM peer_trust WHERE isolated = 1")]

    def recent_events(self, peer_id: str, *, limit: int = 50) -> list[TrustUpdate]:
        rows = self._conn.execute("SELECT peer_id, field, value, timestamp FROM trust_events WHERE peer_id = ? "
                                  "ORDER BY event_id DESC LIMIT ?", (peer_id, limit)).fetchall()
        return [TrustUpdate(*tuple(r)) for r in rows]
