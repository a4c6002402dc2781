"""Bulk bootstrap from Common Crawl: parse WET (WARC extracted-text) files — local or remote, plain or gzip, bounded to
256 MiB compressed and decompressed — dedup (exact hash + SimHash), index; or register a URL list for the crawler
(reference infomesh/index/commoncrawl.py:39-401).  The record parser is a streaming state machine instead of a
regex split over the whole file."""
from __future__ import annotations

import gzip
import io
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Iterator
from urllib.parse import urlparse

from infomesh_b200.crawler.dedup import DeduplicatorDB
from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_MAX_TEXT_SIZE = 102_400
_MAX_WET_FILE_BYTES = 256 * 1024 * 1024
_WET_CHUNK_SIZE = 1024 * 1024
_MIN_TEXT = 50


def _too_large_message(source: str, limit: int | None = None) -> str:
    return f"WET input exceeds {(limit or _MAX_WET_FILE_BYTES) / 2 ** 20:.0f} MiB limit: {source}"


def _read_binary_limited(stream: Any, source: str, *, limit: int | None = None) -> bytes:
    cap = _MAX_WET_FILE_BYTES if limit is None else limit
    parts, total = [], 0
    while chunk := stream.read(_WET_CHUNK_SIZE):
        total += len(chunk)
        if total > cap:
            raise ValueError(_too_large_message(source, cap))
        parts.append(chunk)
    return b"".join(parts)


def _decode_gzip_limited(data: bytes, source: str, *, limit: int | None = None) -> str:
    with gzip.GzipFile(fileobj=io.BytesIO(data)) as gz:
        return _read_binary_limited(gz, source, limit=limit).decode("utf-8", errors="replace")


@dataclass(frozen=True)
class ImportStats:
    total_records: int
    imported: int
    skipped_duplicate: int
    skipped_too_short: int
    skipped_error: int
    elapsed_ms: float


@dataclass
class WETRecord:
    url: str
    text: str
    date: str
    content_length: int


def iter_wet_records(data: str) -> Iterator[WETRecord]:
    """Yield ``conversion`` records with >= 50 characters of text."""
    for raw in data.split("WARC/1.0"):
        raw = raw.lstrip("\r\n")
        if not raw.strip():
            continue
        cut, skip = raw.find("\r\n\r\n"), 4
        if cut < 0:
            cut, skip = raw.find("\n\n"), 2
        if cut < 0:
            continue
        headers: dict[str, str] = {}
        for line in raw[:cut].splitlines():
            k, sep, v = line.partition(":")
            if sep:
                headers[k.strip().lower()] = v.strip()
        if headers.get("warc-type") != "conversion":
            continue
        body = raw[cut + skip:].strip()
        url = headers.get("warc-target-uri", "")
        try:
            length = int(headers.get("content-length", "0"))
        except ValueError:
            length = 0
        if url and len(body) >= _MIN_TEXT:
            yield WETRecord(url, body[:_MAX_TEXT_SIZE], headers.get("warc-date", ""), length)


def parse_wet_content(data: str) -> list[WETRecord]:
    return list(iter_wet_records(data))


class CommonCrawlImporter:
    def __init__(self, store: Any, dedup: DeduplicatorDB | None = None, *, vector_store: Any | None = None):
        self._store, self._dedup, self._vector_store = store, dedup or DeduplicatorDB(), vector_store

    async def import_wet_file(self, path_or_url: str) -> ImportStats:
        t0 = time.monotonic()
        data = await self._download_wet(path_or_url) if path_or_url.startswith(("http://", "https://")) else self._read_local_wet(path_or_url)
        total = imported = dup = short = err = 0
        for rec in iter_wet_records(data):
            total += 1
            try:
                if len(rec.text.strip()) < _MIN_TEXT:
                    short += 1
                    continue
                th = content_hash(rec.text)
                if self._dedup.is_content_seen(th) or self._dedup.is_near_duplicate(rec.text):
                    dup += 1
                    continue
                title = rec.text.split("\n", 1)[0][:200].strip()
                if len(title) < 5:
                    p = urlparse(rec.url)
                    title = p.path.rsplit("/", 1)[-1] or p.netloc
                doc_id = self._store.add_document(url=rec.url, title=title, text=rec.text, raw_html_hash=content_hash(rec.url + rec.date),
                                                  text_hash=th, language=None)
                if doc_id is None:
                    dup += 1
                    continue
                self._dedup.mark_seen(rec.url, th, rec.text)
                if self._vector_store is not None:
                    self._vector_store.add_document(doc_id=doc_id, url=rec.url, title=title, text=rec.text, language=None)
                imported += 1
            except Exception:  # noqa: BLE001
                logger.exception("wet_record_failed")
                err += 1
        logger.info("wet_import_complete", total=total, imported=imported, skipped_dup=dup, skipped_short=short)
        return ImportStats(total, imported, dup, short, err, (time.monotonic() - t0) * 1000)

    async def import_url_list(self, path: str | Path, *, max_urls: int = 10_000) -> ImportStats:
        """Register URLs (one per line, ``#`` comments) as pending; the crawler fetches them later."""
        t0 = time.monotonic()
        urls: list[str] = []
        with open(path, encoding="utf-8") as f:
            for line in f:
                u = line.strip()
                if u and not u.startswith("#"):
                    urls.append(u)
                    if len(urls) >= max_urls:
                        break
        new = 0
        for u in urls:
            if not self._dedup.is_url_seen(u):
                self._dedup.mark_seen(u, "pending")
                new += 1
        return ImportStats(len(urls), new, len(urls) - new, 0, 0, (time.monotonic() - t0) * 1000)

    async def _download_wet(self, url: str) -> str:
        import httpx

        from infomesh_b200.crawler import create_ssl_context

        async with httpx.AsyncClient(timeout=120.0, verify=create_ssl_context()) as client, client.stream("GET", url) as resp:
            resp.raise_for_status()
            declared = resp.headers.get("content-length", "")
            if declared.isdigit() and int(declared) > _MAX_WET_FILE_BYTES:
                raise ValueError(_too_large_message(url))
            parts, total = [], 0
            async for chunk in resp.aiter_bytes(chunk_size=_WET_CHUNK_SIZE):
                total += len(chunk)
                if total > _MAX_WET_FILE_BYTES:
                    raise ValueError(_too_large_message(url))
                parts.append(chunk)
        raw = b"".join(parts)
        return _decode_gzip_limited(raw, url) if url.endswith(".gz") else raw.decode("utf-8", errors="replace")

    def _read_local_wet(self, path: str) -> str:
        if Path(path).stat().st_size > _MAX_WET_FILE_BYTES:
            raise ValueError(_too_large_message(str(path)))
        with open(path, "rb") as f:
            raw = _read_binary_limited(f, str(path))
        return _decode_gzip_limited(raw, str(path)) if str(path).endswith(".gz") else raw.decode("utf-8", errors="replace")
