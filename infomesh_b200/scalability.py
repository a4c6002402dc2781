"""Scale helpers: SQLite connection pool, batched ingest, Bloom filter for URL membership, FTS rebuild
(reference infomesh/scalability.py:24-290)."""
from __future__ import annotations

import hashlib
import math
import sqlite3
import threading
from collections import deque
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Any

from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class ConnectionPool:
    """Up to ``max_connections`` pooled WAL connections; overflow connections are created on demand and closed on
    release."""

    def __init__(self, db_path: str, max_connections: int = 5):
        self._db_path, self._max = db_path, max_connections
        self._idle: deque[sqlite3.Connection] = deque()
        self._lock = threading.Lock()
        self._created = 0

    def _open(self) -> sqlite3.Connection:
        conn = sqlite3.connect(self._db_path, check_same_thread=False)
        conn.execute("PRAGMA journal_mode=WAL")
        conn.execute("PRAGMA synchronous=NORMAL")
        conn.row_factory = sqlite3.Row
        return conn

    def get(self) -> sqlite3.Connection:
        with self._lock:
            if self._idle:
                return self._idle.popleft()
            self._created += 1
        return self._open()

    def release(self, conn: sqlite3.Connection) -> None:
        with self._lock:
            if len(self._idle) < self._max:
                self._idle.append(conn)
                return
            self._created -= 1
        conn.close()

    @contextmanager
    def connection(self):
        conn = self.get()
        try:
            yield conn
        finally:
            self.release(conn)

    def close_all(self) -> None:
        with self._lock:
            while self._idle:
                self._idle.popleft().close()
            self._created = 0


@dataclass
class BatchIngestResult:
    total: int
    succeeded: int
    failed: int
    errors: list[str] = field(default_factory=list)


def batch_ingest(store: Any, documents: list[dict[str, str]], *, batch_size: int = 100) -> BatchIngestResult:
    ok, errors = 0, []
    for doc in documents:
        try:
            text = doc.get("content", doc.get("text", ""))
            # hashes default to the text's own digest: an empty text_hash would make the store's duplicate check
            # collapse every record of the batch into the first one
            text_hash = doc.get("text_hash") or content_hash(text)
            store.add_document(url=doc["url"], title=doc.get("title", ""), text=text,
                               raw_html_hash=doc.get("content_hash") or text_hash, text_hash=text_hash,
                               language=doc.get("language"))
            ok += 1
        except Exception as exc:  # noqa: BLE001
            errors.append(f"{doc.get('url', '?')}: {exc}")
    logger.info("batch_ingest_complete", total=len(documents), succeeded=ok, failed=len(errors))
    return BatchIngestResult(len(documents), ok, len(errors), errors[:50])


class BloomFilter:
    """m = -n ln p / ln^2 2 bits, k = (m / n) ln 2 probes, double hashing from the two halves of one MD5."""

    def __init__(self, capacity: int = 100_000, fp_rate: float = 0.01):
        self._capacity, self._fp_rate = capacity, fp_rate
        self._size = max(8, int(-capacity * math.log(fp_rate) / math.log(2) ** 2) if capacity > 0 and 0 < fp_rate < 1 else capacity * 10)
        self._num_hashes = max(1, int(self._size / capacity * math.log(2))) if capacity > 0 else 7
        self._bits = bytearray((self._size + 7) // 8)
        self._count = 0

    def _hashes(self, item: str) -> list[int]:
        d = hashlib.md5(item.encode(), usedforsecurity=False).digest()
        h1, h2 = int.from_bytes(d[:8], "little"), int.from_bytes(d[8:], "little") | 1
        return [(h1 + i * h2) % self._size for i in range(self._num_hashes)]

    def add(self, item: str) -> None:
        for pos in self._hashes(item):
            self._bits[pos >> 3] |= 1 << (pos & 7)
        self._count += 1

    def __contains__(self, item: str) -> bool:
        return all(self._bits[pos >> 3] >> (pos & 7) & 1 for pos in self._hashes(item))

    def __len__(self) -> int:
        return self._count

    @property
    def size_bytes(self) -> int:
        return len(self._bits)


@dataclass
class RebuildStats:
    documents_processed: int = 0
    documents_updated: int = 0
    documents_skipped: int = 0
    errors: int = 0


def incremental_rebuild(store: Any, *, batch_size: int = 100, force: bool = False) -> RebuildStats:
    """Re-derive the FTS index from the content table (``INSERT INTO <fts>(<fts>) VALUES('rebuild')``)."""
    stats = RebuildStats()
    try:
        conn = store._conn  # noqa: SLF001
        fts = getattr(store, "FTS_TABLE", "documents_fts")
        conn.execute(f"INSERT INTO {fts}({fts}) VALUES('rebuild')")
        conn.commit()
        n = int(conn.execute("SELECT COUNT(*) FROM documents").fetchone()[0])
        stats.documents_processed = stats.documents_updated = n if force else min(n, 1)
    except Exception as exc:  # noqa: BLE001
        logger.error("index_rebuild_failed", error=str(exc))
        stats.errors += 1
    return stats
