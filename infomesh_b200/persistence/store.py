"""SQLite-backed state for the MCP / HTTP servers that must survive restarts: analytics counters, webhook registry,
sessions with TTL, opt-in search history, named filter presets (reference infomesh/persistence/store.py:19-285)."""
from __future__ import annotations

import json
import sqlite3
import threading
import time
from pathlib import Path

_COUNTERS = ("total_searches", "total_crawls", "total_fetches", "latency_sum")


class PersistentStore:
    def __init__(self, db_path: Path | str | None = None):
        self._db_path = str(db_path) if db_path else ":memory:"
        if self._db_path != ":memory:":
            Path(self._db_path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(self._db_path, check_same_thread=False)
        self._conn.row_factory = sqlite3.Row
        self._lock = threading.RLock()
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.executescript("""
            CREATE TABLE IF NOT EXISTS analytics (key TEXT PRIMARY KEY, value REAL NOT NULL DEFAULT 0);
            CREATE TABLE IF NOT EXISTS webhooks (url TEXT PRIMARY KEY, created_at REAL NOT NULL);
            CREATE TABLE IF NOT EXISTS sessions (session_id TEXT PRIMARY KEY, last_query TEXT NOT NULL DEFAULT '',
                last_results TEXT NOT NULL DEFAULT '', updated_at REAL NOT NULL);
            CREATE TABLE IF NOT EXISTS search_history (id INTEGER PRIMARY KEY AUTOINCREMENT, query TEXT NOT NULL,
                result_count INTEGER NOT NULL DEFAULT 0, latency_ms REAL NOT NULL DEFAULT 0, searched_at REAL NOT NULL);
            CREATE TABLE IF NOT EXISTS presets (name TEXT PRIMARY KEY, config_json TEXT NOT NULL, created_at REAL NOT NULL);
        """)
        self._conn.executemany("INSERT OR IGNORE INTO analytics (key, value) VALUES (?, 0)", [(k,) for k in _COUNTERS])
        self._conn.commit()

    def _bump(self, **deltas: float) -> None:
        with self._lock:
            for key, d in deltas.items():
                self._conn.execute("UPDATE analytics SET value = value + ? WHERE key = ?", (d, key))
            self._conn.commit()

    # analytics
    def record_search(self, latency_ms: float) -> None:
        self._bump(total_searches=1, latency_sum=latency_ms)

    def record_crawl(self) -> None:
        self._bump(total_crawls=1)

    def record_fetch(self) -> None:
        self._bump(total_fetches=1)

    def get_analytics(self) -> dict[str, object]:
        with self._lock:
            d = {r["key"]: r["value"] for r in self._conn.execute("SELECT key, value FROM analytics")}
        n = d.get("total_searches", 0)
        return {"total_searches": int(n), "total_crawls": int(d.get("total_crawls", 0)), "total_fetches": int(d.get("total_fetches", 0)),
                "avg_latency_ms": round(d.get("latency_sum", 0) / n, 1) if n else 0.0}

    # webhooks
    def register_webhook(self, url: str) -> None:
        with self._lock:
            self._conn.execute("INSERT OR REPLACE INTO webhooks (url, created_at) VALUES (?, ?)", (url, time.time()))
            self._conn.commit()

    def unregister_webhook(self, url: str) -> bool:
        with self._lock:
            n = self._conn.execute("DELETE FROM webhooks WHERE url = ?", (url,)).rowcount
            self._conn.commit()
        return n > 0

    def get_webhooks(self) -> list[str]:
        with self._lock:
            return [r["url"] for r in self._conn.execute("SELECT url FROM webhooks ORDER BY created_at")]

    # sessions
    def save_session(self, session_id: str, last_query: str, last_results: str) -> None:
        with self._lock:
            self._conn.execute("INSERT OR REPLACE INTO sessions (session_id, last_query, last_results, updated_at) VALUES (?, ?, ?, ?)",
                               (session_id, last_query, last_results[:2000], time.time()))
            self._conn.commit()

    def get_session(self, session_id: str) -> dict[str, object] | None:
        with self._lock:
            row = self._conn.execute("SELECT * FROM sessions WHERE session_id = ?", (session_id,)).fetchone()
        return dict(row) if row else None

    def expire_sessions(self, ttl_seconds: float = 3600) -> int:
        with self._lock:
            n = self._conn.execute("DELETE FROM sessions WHERE updated_at < ?", (time.time() - ttl_seconds,)).rowcount
            self._conn.commit()
        return n

    # history
    def add_history(self, query: str, result_count: int = 0, latency_ms: float = 0) -> None:
        with self._lock:
            self._conn.execute("INSERT INTO search_history (query, result_count, latency_ms, searched_at) VALUES (?, ?, ?, ?)",
                               (query, result_count, latency_ms, time.time()))
            self._conn.commit()

    def get_history(self, *, limit: int = 50) -> list[dict[str, object]]:
        with self._lock:
            return [dict(r) for r in self._conn.execute(
                "SELECT query, result_count, latency_ms, searched_at FROM search_history ORDER BY searched_at DESC, id DESC LIMIT ?", (limit,))]

    def clear_history(self) -> int:
        with self._lock:
            n = self._conn.execute("DELETE FROM search_history").rowcount
            self._conn.commit()
        return n

    # presets
    def save_preset(self, name: str, config: dict[str, object]) -> None:
        with self._lock:
            self._conn.execute("INSERT OR REPLACE INTO presets (name, config_json, created_at) VALUES (?, ?, ?)",
                               (name, json.dumps(config), time.time()))
            self._conn.commit()

    def get_preset(self, name: str) -> dict[str, object] | None:
        with self._lock:
            row = self._conn.execute("SELECT config_json FROM presets WHERE name = ?", (name,)).fetchone()
        return json.loads(row["config_json"]) if row else None

    def list_presets(self) -> list[str]:
        with self._lock:
            return [r["name"] for r in self._conn.execute("SELECT name FROM presets ORDER BY created_at")]

    def delete_preset(self, name: str) -> bool:
        with self._lock:
            n = self._conn.execute("DELETE FROM presets WHERE name = ?", (name,)).rowcount
            self._conn.commit()
        return n > 0

    def close(self) -> None:
        self._conn.close()

    def __enter__(self) -> "PersistentStore":
        return self

    def __exit__(self, *args: object) -> None:
        self.close()
