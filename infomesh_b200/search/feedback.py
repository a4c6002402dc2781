"""Implicit search-quality feedback kept strictly local: fetch / skip / reformulate / cite events keyed by the
SHA-256 of the query (no plaintext), rolled into a per-URL boost (``0.95 * old + 1.0 fetch - 0.3 skip + 2.0 cite``)
(reference infomesh/search/feedback.py:26-300)."""
from __future__ import annotations

import hashlib
import sqlite3
import threading
import time
from dataclasses import dataclass

_REFORMULATION_WINDOW = 60.0
_BOOST_DECAY = 0.95
_MAX_SIGNALS = 100_000
_W_FETCH, _W_SKIP, _W_CITE = 1.0, -0.3, 2.0


@dataclass(frozen=True)
class FeedbackSignal:
    query_hash: str
    result_url: str
    action: str            # fetched | skipped | reformulated | cited
    result_rank: int
    timestamp: float


@dataclass(frozen=True)
class URLBoost:
    url: str
    boost_score: float
    fetch_count: int
    skip_count: int
    cite_count: int


class FeedbackStore:
    def __init__(self, db_path: str | None = None):
        self._db_path = db_path or ":memory:"
        self._conn = sqlite3.connect(self._db_path, check_same_thread=False)
        self._lock = threading.Lock()
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA busy_timeout=3000")
        self._conn.executescript("""
            CREATE TABLE IF NOT EXISTS feedback_signals (
                id INTEGER PRIMARY KEY AUTOINCREMENT, query_hash TEXT NOT NULL, result_url TEXT NOT NULL,
                action TEXT NOT NULL, result_rank INTEGER NOT NULL DEFAULT 0, created_at REAL NOT NULL);
            CREATE INDEX IF NOT EXISTS idx_feedback_url ON feedback_signals(result_url);
            CREATE INDEX IF NOT EXISTS idx_feedback_query ON feedback_signals(query_hash);
            CREATE INDEX IF NOT EXISTS idx_feedback_time ON feedback_signals(created_at);
            CREATE TABLE IF NOT EXISTS url_boosts (
                url TEXT PRIMARY KEY, boost_score REAL NOT NULL DEFAULT 0.0, fetch_count INTEGER NOT NULL DEFAULT 0,
                skip_count INTEGER NOT NULL DEFAULT 0, cite_count INTEGER NOT NULL DEFAULT 0, updated_at REAL NOT NULL);
        """)
        self._conn.commit()

    @staticmethod
    def hash_query(query: str) -> str:
        return hashlib.sha256(query.strip().lower().encode()).hexdigest()

    def _signal(self, query: str, url: str, action: str, rank: int = 0) -> None:
        self._conn.execute("INSERT INTO feedback_signals (query_hash, result_url, action, result_rank, created_at) "
                           "VALUES (?, ?, ?, ?, ?)", (self.hash_query(query), url, action, rank, time.time()))

    def _update_boost(self, url: str, fetch_delta: int = 0, skip_delta: int = 0, cite_delta: int = 0) -> None:
        gain = _W_FETCH * fetch_delta + _W_SKIP * skip_delta + _W_CITE * cite_delta
        self._conn.execute(
            "INSERT INTO url_boosts (url, boost_score, fetch_count, skip_count, cite_count, updated_at) VALUES (?, ?, ?, ?, ?, ?) "
            "ON CONFLICT(url) DO UPDATE SET boost_score = boost_score * ? + ?, fetch_count = fetch_count + ?, "
            "skip_count = skip_count + ?, cite_count = cite_count + ?, updated_at = excluded.updated_at",
            (url, gain, fetch_delta, skip_delta, cite_delta, time.time(), _BOOST_DECAY, gain, fetch_delta, skip_delta, cite_delta))

    def record_fetch(self, query: str, fetched_url: str, result_rank: int) -> None:
        with self._lock:
            self._signal(query, fetched_url, "fetched", result_rank)
            self._update_boost(fetched_url, fetch_delta=1)
            self._conn.commit()
            self._maybe_prune()

    def record_skip(self, query: str, skipped_urls: list[str]) -> None:
        with self._lock:
            for url in skipped_urls:
                self._signal(query, url, "skipped")
                self._update_boost(url, skip_delta=1)
            self._conn.commit()

    def record_reformulation(self, query: str) -> None:
        with self._lock:
            self._signal(query, "", "reformulated")
            self._conn.commit()

    def record_citation(self, query: str, cited_url: str) -> None:
        with self._lock:
            self._signal(query, cited_url, "cited")
            self._update_boost(cited_url, cite_delta=1)
            self._conn.commit()

    def get_boost(self, url: str) -> float:
        row = self._conn.execute("SELECT boost_score FROM url_boosts WHERE url = ?", (url,)).fetchone()
        return float(row[0]) if row else 0.0

    def get_url_stats(self, url: str) -> URLBoost | None:
        row = self._conn.execute("SELECT url, boost_score, fetch_count, skip_count, cite_count FROM url_boosts WHERE url = ?",
                                 (url,)).fetchone()
        return URLBoost(*row) if row else None

    def is_reformulation(self, query: str, window: float = _REFORMULATION_WINDOW) -> bool:
        row = self._conn.execute("SELECT COUNT(*) FROM feedback_signals WHERE query_hash = ? AND created_at > ?",
                                 (self.hash_query(query), time.time() - window)).fetchone()
        return bool(row and row[0] > 0)

    def top_boosted_urls(self, limit: int = 50) -> list[URLBoost]:
        rows = self._conn.execute("SELECT url, boost_score, fetch_count, skip_count, cite_count FROM url_boosts "
                                  "WHERE boost_score > 0 ORDER BY boost_score DESC LIMIT ?", (limit,)).fetchall()
        return [URLBoost(*r) for r in rows]

    def signal_count(self) -> int:
        return int(self._conn.execute("SELECT COUNT(*) FROM feedback_signals").fetchone()[0])

    def _maybe_prune(self, max_signals: int = _MAX_SIGNALS) -> None:
        if self.signal_count() <= max_signals:
            return
        row = self._conn.execute("SELECT created_at FROM feedback_signals ORDER BY created_at DESC LIMIT 1 OFFSET ?",
                                 (max_signals // 2,)).fetchone()
        if row:
            self._conn.execute("DELETE FROM feedback_signals WHERE created_at < ?", (row[0],))
            self._conn.commit()

    def close(self) -> None:
        self._conn.close()
