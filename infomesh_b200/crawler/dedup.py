"""Three-layer deduplication: canonical URL, exact SHA-256 of the text, SimHash near-duplicates
(reference infomesh/crawler/dedup.py:19-233; on-disk table ``seen_urls`` is identical, fingerprints stored as
signed 64-bit and reloaded on start so near-dup detection survives restarts)."""
from __future__ import annotations

import contextlib
import sqlite3
import threading
import time
from pathlib import Path
from urllib.parse import parse_qs, urlencode, urlparse, urlunparse

from infomesh_b200.crawler.simhash import SimHashIndex, simhash
from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_SIGN_BIT = 1 << 63
_MASK64 = (1 << 64) - 1
_TRACKING_PARAMS = frozenset({"utm_source", "utm_medium", "utm_campaign", "utm_term", "utm_content", "fbclid",
                              "gclid", "ref", "source", "mc_cid", "mc_eid"})


def _to_signed64(v: int) -> int:
    return v - (1 << 64) if v >= _SIGN_BIT else v


def _to_unsigned64(v: int) -> int:
    return v & _MASK64


def normalize_url(url: str) -> str:
    """Lower-case scheme/host, drop the fragment and tracking parameters, sort the query, trim a trailing slash."""
    p = urlparse(url)
    query = {k: v for k, v in parse_qs(p.query, keep_blank_values=True).items() if k.lower() not in _TRACKING_PARAMS}
    path = p.path
    if path != "/" and path.endswith("/"):
        path = path.rstrip("/")
    return urlunparse((p.scheme.lower(), p.netloc.lower(), path or "/", p.params,
                       urlencode(sorted(query.items()), doseq=True), ""))


def _doc_key(url_hash: str) -> int:
    return int(url_hash[:8], 16) & 0x7FFFFFFF


class DeduplicatorDB:
    def __init__(self, db_path: str | None = None):
        path = db_path or ":memory:"
        if path != ":memory:":
            Path(path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(path, check_same_thread=False)
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA busy_timeout=5000")
        self._lock = threading.RLock()
        self._conn.execute("CREATE TABLE IF NOT EXISTS seen_urls (url_hash TEXT PRIMARY KEY, url TEXT NOT NULL, "
                           "content_hash TEXT, simhash INTEGER, crawled_at REAL NOT NULL)")
        with contextlib.suppress(sqlite3.OperationalError):       # databases created before the simhash column
            self._conn.execute("ALTER TABLE seen_urls ADD COLUMN simhash INTEGER")
        self._conn.execute("CREATE INDEX IF NOT EXISTS idx_content_hash ON seen_urls (content_hash)")
        self._conn.commit()
        self._simhash_index = SimHashIndex()
        rows = self._conn.execute("SELECT url_hash, simhash FROM seen_urls WHERE simhash IS NOT NULL").fetchall()
        for uh, fp in rows:
            self._simhash_index.add(_doc_key(uh), _to_unsigned64(fp))
        if rows:
            logger.info("simhash_index_reloaded", count=len(rows))

    def is_url_seen(self, url: str) -> bool:
        h = content_hash(normalize_url(url))
        return self._conn.execute("SELECT 1 FROM seen_urls WHERE url_hash = ?", (h,)).fetchone() is not None

    def is_content_seen(self, text_hash: str) -> bool:
        return self._conn.execute("SELECT 1 FROM seen_urls WHERE content_hash = ?", (text_hash,)).fetchone() is not None

    def is_near_duplicate(self, text: str, *, threshold: int = 3) -> bool:
        return bool(self._simhash_index.find_near_duplicates(simhash(text), threshold=threshold))

    def mark_seen(self, url: str, text_hash: str, text: str = "", *, commit: bool = True,
                  fingerprint: int | None = None) -> None:
        """``fingerprint`` lets batch ingest pass a GPU-computed SimHash instead of re-hashing on the CPU."""
        norm = normalize_url(url)
        uh = content_hash(norm)
        fp = fingerprint if fingerprint is not None else (simhash(text) if text else None)
        with self._lock:
            self._conn.execute("INSERT OR REPLACE INTO seen_urls (url_hash, url, content_hash, simhash, crawled_at) "
                               "VALUES (?, ?, ?, ?, ?)",
                               (uh, norm, text_hash, _to_signed64(fp) if fp is not None else None, time.time()))
            if commit:
                self._conn.commit()
        if fp is not None:
            self._simhash_index.add(_doc_key(uh), fp)

    @property
    def simhash_index(self) -> SimHashIndex:
        return self._simhash_index

    def count(self) -> int:
        return int(self._conn.execute("SELECT COUNT(*) FROM seen_urls").fetchone()[0])

    def flush(self) -> None:
        self._conn.commit()

    def close(self) -> None:
        with contextlib.suppress(sqlite3.Error):
            self._conn.close()
