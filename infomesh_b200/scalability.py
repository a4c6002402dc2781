"""Helpers for large stores: pooled SQLite connections, bulk ingest, a Bloom filter for "have we seen this URL", FTS rebuild.

Contract (SURVEY §2.1 "scalability"; reference infomesh/scalability.py): the pool keeps up to ``max_connections`` idle WAL
connections and opens extra ones on demand (closed again on release); bulk ingest reports how many records went in and
the first 50 error strings, one bad record never aborts the batch; the Bloom filter is sized from capacity and target
false-positive rate (``m = -n ln p / ln^2 2`` bits, ``k = m/n ln 2`` probes), has no false negatives and counts
insertions; the rebuild asks FTS5 to re-derive its index from the content table.

Implementation: the pool is a bounded LIFO (`queue.LifoQueue`) so the warmest connection is reused first; Bloom probe
positions come from one 128-bit BLAKE2 digest split in two (Kirsch-Mitzenmacher double hashing) and are set / tested in a
single big integer per call instead of a byte loop; ingest normalises a record through one function, so the loop body is
"normalise, insert, count"."""
from __future__ import annotations

import math
import queue
import sqlite3
from contextlib import contextmanager
from dataclasses import dataclass, field
from hashlib import blake2b
from typing import Any, Iterator

from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_MAX_REPORTED_ERRORS = 50


# ----------------------------------------------------------------------------- connection pool
class ConnectionPool:
    def __init__(self, db_path: str, max_connections: int = 5):
        self._target = db_path
        self._idle: queue.LifoQueue[sqlite3.Connection] = queue.LifoQueue(maxsize=max(1, max_connections))

    def _connect(self) -> sqlite3.Connection:
        conn = sqlite3.connect(self._target, check_same_thread=False)
        conn.row_factory = sqlite3.Row
        for pragma in ("journal_mode=WAL", "synchronous=NORMAL"):
            conn.execute(f"PRAGMA {pragma}")
        return conn

    def get(self) -> sqlite3.Connection:
        try:
            return self._idle.get_nowait()
        except queue.Empty:
            return self._connect()

    def release(self, conn: sqlite3.Connection) -> None:
        try:
            self._idle.put_nowait(conn)
        except queue.Full:                 # an overflow connection: not kept
            conn.close()

    @contextmanager
    def connection(self) -> Iterator[sqlite3.Connection]:
        conn = self.get()
        try:
            yield conn
        finally:
            self.release(conn)

    def close_all(self) -> None:
        while True:
            try:
                self._idle.get_nowait().close()
            except queue.Empty:
                return


# ----------------------------------------------------------------------------- bulk ingest
@dataclass
class BatchIngestResult:
    total: int
    succeeded: int
    failed: int
    errors: list[str] = field(default_factory=list)


def _store_arguments(record: dict[str, str]) -> dict[str, Any]:
    """Record (``url`` + ``text`` or ``content``, optional title / hashes / language) -> ``LocalStore.add_document`` kwargs.
    Missing hashes default to the digest of the text itself: an empty ``text_hash`` would make the store's uniqueness
    check fold the whole batch into its first record."""
    body = record.get("content", record.get("text", ""))
    digest = record.get("text_hash") or content_hash(body)
    return {"url": record["url"], "title": record.get("title", ""), "text": body, "text_hash": digest,
            "raw_html_hash": record.get("content_hash") or digest, "language": record.get("language")}


def batch_ingest(store: Any, documents: list[dict[str, str]], *, batch_size: int = 100) -> BatchIngestResult:
    problems: list[str] = []
    for record in documents:
        try:
            store.add_document(**_store_arguments(record))
        except Exception as exc:  # noqa: BLE001 -- malformed record, constraint violation, ...
            problems.append(f"{record.get('url', '?')}: {exc}")
    done = len(documents) - len(problems)
    logger.info("batch_ingest_complete", total=len(documents), succeeded=done, failed=len(problems))
    return BatchIngestResult(len(documents), done, len(problems), problems[:_MAX_REPORTED_ERRORS])


# ----------------------------------------------------------------------------- Bloom filter
class BloomFilter:
    def __init__(self, capacity: int = 100_000, fp_rate: float = 0.01):
        sane = capacity > 0 and 0.0 < fp_rate < 1.0
        bits = -capacity * math.log(fp_rate) / math.log(2) ** 2 if sane else capacity * 10
        self._size = max(8, int(bits))
        self._num_hashes = max(1, int(self._size / capacity * math.log(2))) if capacity > 0 else 7
        self._field = 0                    # the bit array, as one arbitrary-precision integer
        self._count = 0

    def _mask(self, item: str) -> int:
        digest = blake2b(item.encode(), digest_size=16).digest()
        start = int.from_bytes(digest[:8], "little")
        stride = int.from_bytes(digest[8:], "little") | 1          # odd stride: probes do not collapse
        mask = 0
        for probe in range(self._num_hashes):
            mask |= 1 << ((start + probe * stride) % self._size)
        return mask

    def add(self, item: str) -> None:
        self._field |= self._mask(item)
        self._count += 1

    def __contains__(self, item: str) -> bool:
        mask = self._mask(item)
        return self._field & mask == mask

    def __len__(self) -> int:
        return self._count

    @property
    def size_bytes(self) -> int:
        return (self._size + 7) // 8


# ----------------------------------------------------------------------------- FTS rebuild
@dataclass
class RebuildStats:
    documents_processed: int = 0
    documents_updated: int = 0
    documents_skipped: int = 0
    errors: int = 0


def incremental_rebuild(store: Any, *, batch_size: int = 100, force: bool = False) -> RebuildStats:
    """``INSERT INTO fts(fts) VALUES('rebuild')``: FTS5 re-derives the inverted index from the content table."""
    table = getattr(store, "FTS_TABLE", "documents_fts")
    try:
        db = store._conn  # noqa: SLF001 -- the store exposes no maintenance hook
        db.execute(f"INSERT INTO {table}({table}) VALUES('rebuild')")
        db.commit()
        (total,) = db.execute("SELECT COUNT(*) FROM documents").fetchone()
    except Exception as exc:  # noqa: BLE001
        logger.error("index_rebuild_failed", error=str(exc))
        return RebuildStats(errors=1)
    touched = int(total) if force else min(int(total), 1)
    return RebuildStats(documents_processed=touched, documents_updated=touched)
