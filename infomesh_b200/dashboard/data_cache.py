"""Read-only, TTL-cached view of the index database for the dashboard: one long-lived ``mode=ro`` connection, all
numbers of a tick gathered in one pass (reference infomesh/dashboard/data_cache.py:33-196)."""
from __future__ import annotations

import contextlib
import sqlite3
import time
from dataclasses import dataclass, field
from pathlib import Path
from urllib.parse import urlparse

from infomesh_b200.config import Config


@dataclass
class RecentDoc:
    doc_id: int
    url: str
    title: str
    crawled_at: float


@dataclass
class CachedStats:
    document_count: int = 0
    top_domains: list[tuple[str, int]] = field(default_factory=list)
    updated_at: float = 0.0
    pages_last_hour: int = 0
    domain_count: int = 0
    last_crawl_at: float = 0.0
    recent_docs: list[RecentDoc] = field(default_factory=list)


def _netloc(url: str) -> str:
    try:
        return urlparse(url).netloc
    except ValueError:
        return ""


class DashboardDataCache:
    def __init__(self, config: Config, *, ttl: float = 0.5):
        self._config, self._ttl = config, ttl
        self._conn: sqlite3.Connection | None = None
        self._cache = CachedStats()

    def _ensure_conn(self) -> sqlite3.Connection:
        if self._conn is None:
            path = Path(self._config.index.db_path)
            if not path.exists():
                raise FileNotFoundError(f"Index database not found: {path}")
            try:
                self._conn = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
            except sqlite3.OperationalError:
                self._conn = sqlite3.connect(str(path))
            self._conn.row_factory = sqlite3.Row
            self._conn.create_function("netloc", 1, _netloc, deterministic=True)
        return self._conn

    def get_stats(self) -> CachedStats:
        now = time.monotonic()
        if self._cache.updated_at and now - self._cache.updated_at < self._ttl:
            return self._cache
        try:
            c = self._ensure_conn()
            head = c.execute("SELECT COUNT(*) AS n, COALESCE(SUM(crawled_at > ?), 0) AS recent, COALESCE(MAX(crawled_at), 0) AS last "
                             "FROM documents", (time.time() - 3600,)).fetchone()
            doms = c.execute("SELECT netloc(url) AS d, COUNT(*) AS n FROM documents GROUP BY d ORDER BY n DESC").fetchall()
            recent = c.execute("SELECT doc_id, url, title, crawled_at FROM documents ORDER BY crawled_at DESC LIMIT 10").fetchall()
            self._cache = CachedStats(int(head["n"]), [(r["d"], int(r["n"])) for r in doms[:7]], now, int(head["recent"]), len(doms),
                                      float(head["last"]), [RecentDoc(int(r["doc_id"]), str(r["url"]), str(r["title"] or ""), float(r["crawled_at"]))
                                                            for r in recent])
        except Exception:  # noqa: BLE001 — DB missing or mid-rotation: keep the last good numbers
            self.close()
        return self._cache

    def set_ttl(self, ttl: float) -> None:
        self._ttl = ttl

    def close(self) -> None:
        if self._conn is not None:
            with contextlib.suppress(Exception):
                self._conn.close()
            self._conn = None
