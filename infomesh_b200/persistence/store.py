"""Server-side state that has to outlive a restart of the MCP / HTTP front ends.

Contract (SURVEY §2.1 persistence/; reference infomesh/persistence/store.py): one SQLite file (WAL) with five tables --
``analytics`` (named counters; the average search latency is ``latency_sum / total_searches`` rounded to 0.1 ms),
``webhooks`` (URL registry, oldest first), ``sessions`` (last query + at most 2000 characters of results per session,
expired by age), ``search_history`` (opt-in log, newest first) and ``presets`` (named JSON filter sets, oldest first).
The table and column names are the on-disk format shared with the reference, so they are kept.

Implementation: every statement goes through one of two helpers -- ``_write`` (locked, committed, returns the affected row
count) and ``_read`` (locked, returns plain dicts) -- so the public methods are one-liners over SQL constants; the schema
is data (table -> column DDL) and is also what seeds the counter rows."""
from __future__ import annotations

import json
import sqlite3
import threading
import time
from pathlib import Path
from typing import Any, Iterable

_SCHEMA: dict[str, str] = {
    "analytics": "key TEXT PRIMARY KEY, value REAL NOT NULL DEFAULT 0",
    "webhooks": "url TEXT PRIMARY KEY, created_at REAL NOT NULL",
    "sessions": "session_id TEXT PRIMARY KEY, last_query TEXT NOT NULL DEFAULT '', last_results TEXT NOT NULL DEFAULT '', "
                "updated_at REAL NOT NULL",
    "search_history": "id INTEGER PRIMARY KEY AUTOINCREMENT, query TEXT NOT NULL, result_count INTEGER NOT NULL DEFAULT 0, "
                      "latency_ms REAL NOT NULL DEFAULT 0, searched_at REAL NOT NULL",
    "presets": "name TEXT PRIMARY KEY, config_json TEXT NOT NULL, created_at REAL NOT NULL",
}
_COUNTER_NAMES = ("total_searches", "total_crawls", "total_fetches", "latency_sum")
_SESSION_RESULT_CHARS = 2000


class PersistentStore:
    def __init__(self, db_path: Path | str | None = None):
        target = str(db_path) if db_path else ":memory:"
        if target != ":memory:":
            Path(target).parent.mkdir(parents=True, exist_ok=True)
        self._db_path = target
        self._guard = threading.RLock()
        self._db = sqlite3.connect(target, check_same_thread=False)
        self._db.row_factory = sqlite3.Row
        self._db.execute("PRAGMA journal_mode=WAL")
        for table, columns in _SCHEMA.items():
            self._db.execute(f"CREATE TABLE IF NOT EXISTS {table} ({columns})")
        self._write_many("INSERT OR IGNORE INTO analytics (key, value) VALUES (?, 0)", ((name,) for name in _COUNTER_NAMES))

    # ---- statement helpers
    def _write(self, sql: str, params: tuple = ()) -> int:
        with self._guard:
            touched = self._db.execute(sql, params).rowcount
            self._db.commit()
        return touched

    def _write_many(self, sql: str, rows: Iterable[tuple]) -> None:
        with self._guard:
            self._db.executemany(sql, rows)
            self._db.commit()

    def _read(self, sql: str, params: tuple = ()) -> list[dict[str, Any]]:
        with self._guard:
            return [dict(row) for row in self._db.execute(sql, params)]

    def _first(self, sql: str, params: tuple = ()) -> dict[str, Any] | None:
        found = self._read(sql, params)
        return found[0] if found else None

    # ---- analytics counters
    def _count(self, **increments: float) -> None:
        self._write_many("UPDATE analytics SET value = value + ? WHERE key = ?", ((step, name) for name, step in increments.items()))

    def record_search(self, latency_ms: float) -> None:
        self._count(total_searches=1, latency_sum=latency_ms)

    def record_crawl(self) -> None:
        self._count(total_crawls=1)

    def record_fetch(self) -> None:
        self._count(total_fetches=1)

    def get_analytics(self) -> dict[str, object]:
        value = {row["key"]: row["value"] for row in self._read("SELECT key, value FROM analytics")}
        searches = value.get("total_searches", 0)
        report: dict[str, object] = {name: int(value.get(name, 0)) for name in _COUNTER_NAMES[:3]}
        report["avg_latency_ms"] = round(value.get("latency_sum", 0) / searches, 1) if searches else 0.0
        return report

    # ---- webhook registry
    def register_webhook(self, url: str) -> None:
        self._write("INSERT OR REPLACE INTO webhooks (url, created_at) VALUES (?, ?)", (url, time.time()))

    def unregister_webhook(self, url: str) -> bool:
        return self._write("DELETE FROM webhooks WHERE url = ?", (url,)) > 0

    def get_webhooks(self) -> list[str]:
        return [row["url"] for row in self._read("SELECT url FROM webhooks ORDER BY created_at")]

    # ---- conversational sessions
    def save_session(self, session_id: str, last_query: str, last_results: str) -> None:
        self._write("INSERT OR REPLACE INTO sessions (session_id, last_query, last_results, updated_at) VALUES (?, ?, ?, ?)",
                    (session_id, last_query, last_results[:_SESSION_RESULT_CHARS], time.time()))

    def get_session(self, session_id: str) -> dict[str, object] | None:
        return self._first("SELECT * FROM sessions WHERE session_id = ?", (session_id,))

    def expire_sessions(self, ttl_seconds: float = 3600) -> int:
        return self._write("DELETE FROM sessions WHERE updated_at < ?", (time.time() - ttl_seconds,))

    # ---- search history
    def add_history(self, query: str, result_count: int = 0, latency_ms: float = 0) -> None:
        self._write("INSERT INTO search_history (query, result_count, latency_ms, searched_at) VALUES (?, ?, ?, ?)",
                    (query, result_count, latency_ms, time.time()))

    def get_history(self, *, limit: int = 50) -> list[dict[str, object]]:
        return self._read("SELECT query, result_count, latency_ms, searched_at FROM search_history "
                          "ORDER BY searched_at DESC, id DESC LIMIT ?", (limit,))

    def clear_history(self) -> int:
        return self._write("DELETE FROM search_history")

    # ---- filter presets
    def save_preset(self, name: str, config: dict[str, object]) -> None:
        self._write("INSERT OR REPLACE INTO presets (name, config_json, created_at) VALUES (?, ?, ?)", (name, json.dumps(config), time.time()))

    def get_preset(self, name: str) -> dict[str, object] | None:
        hit = self._first("SELECT config_json FROM presets WHERE name = ?", (name,))
        return None if hit is None else json.loads(hit["config_json"])

    def list_presets(self) -> list[str]:
        return [row["name"] for row in self._read("SELECT name FROM presets ORDER BY created_at")]

    def delete_preset(self, name: str) -> bool:
        return self._write("DELETE FROM presets WHERE name = ?", (name,)) > 0

    # ---- lifecycle
    def close(self) -> None:
        self._db.close()

    def __enter__(self) -> "PersistentStore":
        return self

    def __exit__(self, *exc_info: object) -> None:
        self.close()
