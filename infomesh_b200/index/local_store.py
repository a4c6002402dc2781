"""LocalStore — the durable document store + CPU keyword index (SQLite FTS5).

On-disk format and public API are those of reference infomesh/index/local_store.py:62-634 (``documents`` table +
external-content FTS5 table ``documents_fts(title, text)`` kept in sync by triggers, WAL, 5 s busy timeout,
additive schema migration for the recrawl / JS columns).  In this rebuild SQLite is the *metadata / text* store and
the CPU path for config #1; the same documents are mirrored into HBM-resident shards (``engine.gpu_index``) where
BM25 and dense retrieval run on the GPU.  ``add_listener`` lets the GPU index follow inserts / deletes.
"""
from __future__ import annotations

import sqlite3
import threading
import time
from collections.abc import Callable
from dataclasses import dataclass
from pathlib import Path
from typing import Any

from infomesh_b200.compression.zstd import Compressor
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_ALLOWED_TOKENIZERS = frozenset({"unicode61", "ascii", "porter", "trigram"})
MAX_SEARCH_LIMIT = 1000
MAX_SEARCH_OFFSET = 10_000

# host part of a URL, in SQL (used for domain filters and domain statistics)
_HOST_SQL = ("SUBSTR(url, INSTR(url, '://') + 3, CASE WHEN INSTR(SUBSTR(url, INSTR(url, '://') + 3), '/') > 0 "
             "THEN INSTR(SUBSTR(url, INSTR(url, '://') + 3), '/') - 1 ELSE LENGTH(url) END)")

_RECRAWL_COLUMNS: tuple[tuple[str, str], ...] = (
    ("compressed_text", "BLOB"),
    ("raw_html_hash", "TEXT NOT NULL DEFAULT ''"),
    ("etag", "TEXT"),
    ("last_modified", "TEXT"),
    ("recrawl_interval", "INTEGER DEFAULT 604800"),
    ("stale_count", "INTEGER DEFAULT 0"),
    ("last_recrawl_at", "REAL"),
    ("change_frequency", "REAL DEFAULT 0.0"),
    ("js_required", "INTEGER DEFAULT 0"),
)


@dataclass(frozen=True)
class IndexedDocument:
    doc_id: int
    url: str
    title: str
    text: str
    language: str | None
    raw_html_hash: str
    text_hash: str
    crawled_at: float
    etag: str | None = None
    last_modified: str | None = None
    recrawl_interval: int = 604800
    stale_count: int = 0
    last_recrawl_at: float | None = None
    change_frequency: float = 0.0
    js_required: int = 0


@dataclass(frozen=True)
class SearchResult:
    doc_id: int
    url: str
    title: str
    snippet: str
    score: float
    language: str | None
    crawled_at: float


def _serialised(cls):
    """One SQLite connection is shared by every thread that touches the store (crawler, MCP handlers, GPU index workers);
    the ``sqlite3`` module leaves serialising statement + fetch on a shared connection to the caller.  Every public method
    therefore runs under the store's re-entrant lock (generators lock per batch themselves)."""
    import functools
    import inspect

    for name, fn in list(vars(cls).items()):
        if name.startswith("_") or not inspect.isfunction(fn) or inspect.isgeneratorfunction(fn) or inspect.iscoroutinefunction(fn):
            continue

        def locked(self, *a, __fn=fn, **kw):
            with self._lock:
                return __fn(self, *a, **kw)

        setattr(cls, name, functools.wraps(fn)(locked))
    return cls


@_serialised
class LocalStore:
    def __init__(self, db_path: Path | str | None = None, tokenizer: str = "unicode61", *,
                 compression_enabled: bool = False, compression_level: int = 3):
        if tokenizer not in _ALLOWED_TOKENIZERS:  # the name is interpolated into DDL: whitelist only
            raise ValueError(f"Invalid tokenizer '{tokenizer}'; allowed: {sorted(_ALLOWED_TOKENIZERS)}")
        self._tokenizer = tokenizer
        self._db_path = str(db_path) if db_path else ":memory:"
        if self._db_path != ":memory:":
            Path(self._db_path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(self._db_path, check_same_thread=False)
        self._conn.row_factory = sqlite3.Row
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA busy_timeout=5000")
        self._lock = threading.RLock()
        self._compressor = Compressor(level=compression_level) if compression_enabled else None
        self._listeners: list[Callable[[str, dict[str, Any]], None]] = []
        self._create_schema()

    # ------------------------------------------------------------------ schema
    def _create_schema(self) -> None:
        ddl = f"""
        CREATE TABLE IF NOT EXISTS documents (
            doc_id INTEGER PRIMARY KEY AUTOINCREMENT,
            url TEXT UNIQUE NOT NULL,
            title TEXT NOT NULL DEFAULT '',
            text TEXT NOT NULL,
            compressed_text BLOB,
            language TEXT,
            raw_html_hash TEXT NOT NULL,
            text_hash TEXT UNIQUE NOT NULL,
            crawled_at REAL NOT NULL
        );
        CREATE VIRTUAL TABLE IF NOT EXISTS documents_fts USING fts5(
            title, text, content='documents', content_rowid='doc_id', tokenize='{self._tokenizer}'
        );
        CREATE TRIGGER IF NOT EXISTS documents_ai AFTER INSERT ON documents BEGIN
            INSERT INTO documents_fts(rowid, title, text) VALUES (new.doc_id, new.title, new.text);
        END;
        CREATE TRIGGER IF NOT EXISTS documents_ad AFTER DELETE ON documents BEGIN
            INSERT INTO documents_fts(documents_fts, rowid, title, text)
            VALUES ('delete', old.doc_id, old.title, old.text);
        END;
        CREATE TRIGGER IF NOT EXISTS documents_au AFTER UPDATE ON documents BEGIN
            INSERT INTO documents_fts(documents_fts, rowid, title, text)
            VALUES ('delete', old.doc_id, old.title, old.text);
            INSERT INTO documents_fts(rowid, title, text) VALUES (new.doc_id, new.title, new.text);
        END;
        """
        with self._lock:
            self._conn.executescript(ddl)
            have = {row[1] for row in self._conn.execute("PRAGMA table_info(documents)")}
            for col, decl in _RECRAWL_COLUMNS:
                if col not in have:
                    self._conn.execute(f"ALTER TABLE documents ADD COLUMN {col} {decl}")
                    logger.info("schema_migrated", column=col)
            self._conn.commit()

    # ------------------------------------------------------------------ change feed (GPU mirror)
    def add_listener(self, fn: Callable[[str, dict[str, Any]], None]) -> None:
        """``fn(event, payload)`` is called after ``add`` / ``delete`` / ``update`` commits."""
        self._listeners.append(fn)

    def _notify(self, event: str, **payload: Any) -> None:
        for fn in self._listeners:
            try:
                fn(event, payload)
            except Exception as exc:  # noqa: BLE001 — a mirror failure must never lose a document
                logger.warning("store_listener_failed", store_event=event, error=str(exc))

    # ------------------------------------------------------------------ writes
    def add_document(self, url: str, title: str, text: str, raw_html_hash: str, text_hash: str, *,
                     language: str | None = None, js_required: bool = False) -> int | None:
        """Insert; ``None`` when the URL or the text hash already exists."""
        blob = self._compressor.compress_text(text) if self._compressor else None
        try:
            with self._lock:
                cur = self._conn.execute(
                    "INSERT INTO documents (url, title, text, compressed_text, language, raw_html_hash, text_hash, "
                    "crawled_at, js_required) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?)",
                    (url, title, text, blob, language, raw_html_hash, text_hash, time.time(), int(bool(js_required))))
                self._conn.commit()
                doc_id = cur.lastrowid
        except sqlite3.IntegrityError:
            logger.debug("doc_duplicate", url=url)
            return None
        logger.info("doc_indexed", doc_id=doc_id, url=url, text_len=len(text))
        self._notify("add", doc_id=doc_id, url=url, title=title, text=text, language=language)
        return doc_id

    def delete_document(self, doc_id: int) -> bool:
        with self._lock:
            cur = self._conn.execute("DELETE FROM documents WHERE doc_id = ?", (doc_id,))
            self._conn.commit()
        if cur.rowcount > 0:
            self._notify("delete", doc_id=doc_id)
            return True
        return False

    def soft_delete(self, url: str) -> bool:
        """Remove a stale document by URL (the reference's "soft" delete is a hard DELETE as well)."""
        row = self._conn.execute("SELECT doc_id FROM documents WHERE url = ?", (url,)).fetchone()
        if row is None:
            return False
        ok = self.delete_document(int(row["doc_id"]))
        if ok:
            logger.info("doc_soft_deleted", url=url)
        return ok

    def update_document(self, url: str, *, title: str | None = None, text: str | None = None,
                        text_hash: str | None = None, raw_html_hash: str | None = None, etag: str | None = None,
                        last_modified: str | None = None, recrawl_interval: int | None = None,
                        stale_count: int | None = None, last_recrawl_at: float | None = None,
                        change_frequency: float | None = None) -> bool:
        """Write only the provided columns; the UPDATE trigger refreshes the FTS rows."""
        changes = {k: v for k, v in dict(
            title=title, text=text, text_hash=text_hash, raw_html_hash=raw_html_hash, etag=etag,
            last_modified=last_modified, recrawl_interval=recrawl_interval, stale_count=stale_count,
            last_recrawl_at=last_recrawl_at, change_frequency=change_frequency).items() if v is not None}
        if not changes:
            return False
        if text is not None and self._compressor:
            changes["compressed_text"] = self._compressor.compress_text(text)
        assignments = ", ".join(f"{col} = ?" for col in changes)
        with self._lock:
            cur = self._conn.execute(f"UPDATE documents SET {assignments} WHERE url = ?", (*changes.values(), url))
            self._conn.commit()
        if cur.rowcount > 0:
            if text is not None or title is not None:
                doc = self.get_document_by_url(url)
                if doc is not None:
                    self._notify("update", doc_id=doc.doc_id, url=url, title=doc.title, text=doc.text,
                                 language=doc.language)
            return True
        return False

    # ------------------------------------------------------------------ search
    def search(self, query: str, *, limit: int = 10, offset: int = 0, language: str | None = None,
               date_from: float | None = None, date_to: float | None = None,
               include_domains: list[str] | None = None, exclude_domains: list[str] | None = None
               ) -> list[SearchResult]:
        """FTS5 ``MATCH`` ordered by ``bm25()``; scores are returned positive (higher = better)."""
        limit = min(max(int(limit), 1), MAX_SEARCH_LIMIT)
        offset = min(max(int(offset), 0), MAX_SEARCH_OFFSET)
        where = ["documents_fts MATCH ?"]
        args: list[Any] = [query]
        if language:
            where.append("d.language = ?")
            args.append(language)
        if date_from is not None:
            where.append("d.crawled_at >= ?")
            args.append(date_from)
        if date_to is not None:
            where.append("d.crawled_at <= ?")
            args.append(date_to)
        if include_domains:
            where.append(f"{_HOST_SQL} IN ({', '.join('?' * len(include_domains))})")
            args.extend(include_domains)
        if exclude_domains:
            where.append(f"{_HOST_SQL} NOT IN ({', '.join('?' * len(exclude_domains))})")
            args.extend(exclude_domains)
        sql = ("SELECT d.doc_id, d.url, d.title, snippet(documents_fts, 1, '<b>', '</b>', '...', 40) AS snippet, "
               "bm25(documents_fts) AS score, d.language, d.crawled_at FROM documents_fts "
               "JOIN documents d ON d.doc_id = documents_fts.rowid WHERE " + " AND ".join(where) +
               " ORDER BY bm25(documents_fts) LIMIT ? OFFSET ?")
        try:
            with self._lock:
                rows = self._conn.execute(sql, (*args, limit, offset)).fetchall()
        except sqlite3.OperationalError as exc:
            logger.error("search_error", query=query, error=str(exc))
            return []
        logger.debug("local_search", query=query, results=len(rows))
        return [SearchResult(r["doc_id"], r["url"], r["title"], r["snippet"], abs(r["score"]), r["language"],
                             r["crawled_at"]) for r in rows]

    def suggest(self, prefix: str, *, limit: int = 10) -> list[str]:
        limit = min(max(int(limit), 1), 50)
        needle = prefix.replace("%", "").replace("_", "")[:100]
        try:
            rows = self._conn.execute(
                "SELECT DISTINCT title FROM documents WHERE title LIKE ? COLLATE NOCASE "
                "ORDER BY crawled_at DESC LIMIT ?", (f"%{needle}%", limit)).fetchall()
        except sqlite3.OperationalError:
            return []
        return [r["title"] for r in rows]

    # ------------------------------------------------------------------ reads
    def _to_document(self, row: sqlite3.Row) -> IndexedDocument:
        keys = row.keys()
        text = row["text"]
        if not text and "compressed_text" in keys and row["compressed_text"] and self._compressor:
            text = self._compressor.decompress_text(row["compressed_text"])

        def opt(name: str, default: Any) -> Any:
            return row[name] if name in keys and row[name] is not None else default

        return IndexedDocument(
            doc_id=row["doc_id"], url=row["url"], title=row["title"], text=text, language=row["language"],
            raw_html_hash=row["raw_html_hash"], text_hash=row["text_hash"], crawled_at=row["crawled_at"],
            etag=opt("etag", None), last_modified=opt("last_modified", None),
            recrawl_interval=opt("recrawl_interval", 604800), stale_count=opt("stale_count", 0),
            last_recrawl_at=opt("last_recrawl_at", None), change_frequency=opt("change_frequency", 0.0),
            js_required=opt("js_required", 0))

    def get_document(self, doc_id: int) -> IndexedDocument | None:
        row = self._conn.execute("SELECT * FROM documents WHERE doc_id = ?", (doc_id,)).fetchone()
        return self._to_document(row) if row else None

    def get_document_by_url(self, url: str) -> IndexedDocument | None:
        row = self._conn.execute("SELECT * FROM documents WHERE url = ?", (url,)).fetchone()
        return self._to_document(row) if row else None

    def iter_documents(self, batch: int = 1000, after: int = 0):
        """Stream every document with ``doc_id > after`` in doc_id order (GPU shard build / incremental append, snapshots)."""
        last = int(after)
        while True:
            with self._lock:
                rows = self._conn.execute("SELECT * FROM documents WHERE doc_id > ? ORDER BY doc_id LIMIT ?",
                                          (last, batch)).fetchall()
            if not rows:
                return
            for r in rows:
                yield self._to_document(r)
            last = rows[-1]["doc_id"]

    def get_stats(self) -> dict[str, int]:
        row = self._conn.execute("SELECT COUNT(*) AS n FROM documents").fetchone()
        return {"document_count": int(row["n"]) if row else 0}

    def max_doc_id(self) -> int:
        """Highest document id in the store (0 when empty): the GPU index uses it to tell whether saved device segments
        still describe the store."""
        row = self._conn.execute("SELECT MAX(doc_id) AS m FROM documents").fetchone()
        return int(row["m"] or 0) if row else 0

    def get_top_domains(self, limit: int = 7) -> list[tuple[str, int]]:
        rows = self._conn.execute(f"SELECT {_HOST_SQL} AS domain, COUNT(*) AS cnt FROM documents GROUP BY domain "
                                  "ORDER BY cnt DESC LIMIT ?", (limit,)).fetchall()
        return [(r["domain"], r["cnt"]) for r in rows]

    def get_js_required_domains(self, limit: int = 20) -> list[tuple[str, int, int]]:
        rows = self._conn.execute(
            f"SELECT {_HOST_SQL} AS domain, SUM(CASE WHEN js_required = 1 THEN 1 ELSE 0 END) AS js_cnt, "
            "COUNT(*) AS total FROM documents GROUP BY domain HAVING js_cnt > 0 "
            "ORDER BY CAST(js_cnt AS REAL) / total DESC LIMIT ?", (limit,)).fetchall()
        return [(r["domain"], r["js_cnt"], r["total"]) for r in rows]

    def get_domain_count(self) -> int:
        row = self._conn.execute(f"SELECT COUNT(DISTINCT {_HOST_SQL}) AS cnt FROM documents").fetchone()
        return int(row["cnt"]) if row else 0

    def export_documents(self) -> list[dict[str, object]]:
        cols = ("url", "title", "text", "language", "raw_html_hash", "text_hash", "crawled_at")
        rows = self._conn.execute(f"SELECT {', '.join(cols)} FROM documents ORDER BY doc_id").fetchall()
        return [{c: r[c] for c in cols} for r in rows]

    def get_documents_for_publish(self, *, limit: int = 500, offset: int = 0) -> list[dict[str, object]]:
        limit = min(max(int(limit), 1), 10_000)
        rows = self._conn.execute("SELECT doc_id, url, title, text FROM documents ORDER BY doc_id LIMIT ? OFFSET ?",
                                  (limit, max(0, int(offset)))).fetchall()
        return [dict(r) for r in rows]

    def get_recrawl_candidates(self, *, limit: int = 200) -> list[dict[str, object]]:
        rows = self._conn.execute(
            "SELECT doc_id, url, text_hash, etag, last_modified, recrawl_interval, stale_count, change_frequency, "
            "crawled_at, last_recrawl_at FROM documents WHERE stale_count < 3 "
            "ORDER BY last_recrawl_at ASC NULLS FIRST LIMIT ?", (limit,)).fetchall()
        return [dict(r) for r in rows]

    def get_compression_stats(self) -> dict[str, object]:
        row = self._conn.execute("SELECT COALESCE(SUM(LENGTH(text)), 0) AS raw, "
                                 "COALESCE(SUM(LENGTH(compressed_text)), 0) AS comp, "
                                 "SUM(CASE WHEN compressed_text IS NOT NULL THEN 1 ELSE 0 END) AS n "
                                 "FROM documents").fetchone()
        raw, comp = int(row["raw"] or 0), int(row["comp"] or 0)
        return {"enabled": self._compressor is not None, "compressed_docs": int(row["n"] or 0),
                "raw_bytes": raw, "compressed_bytes": comp, "ratio": round(raw / comp, 3) if comp else 0.0}

    # ------------------------------------------------------------------ maintenance
    def optimize(self) -> None:
        """Merge FTS5 segments (run periodically)."""
        try:
            with self._lock:
                self._conn.execute("INSERT INTO documents_fts(documents_fts) VALUES('optimize')")
                self._conn.commit()
        except sqlite3.Error as exc:
            logger.debug("fts_optimize_failed", error=str(exc))

    def close(self) -> None:
        try:
            self._conn.close()
        except sqlite3.Error:
            pass

    def __enter__(self) -> "LocalStore":
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()
