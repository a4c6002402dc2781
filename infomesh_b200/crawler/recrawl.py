"""Adaptive re-crawling (reference infomesh/crawler/recrawl.py:29-325): interval from the EMA change frequency
(6 h / 24 h / 7 d / 30 d), conditional GET (ETag / If-Modified-Since), 3 consecutive failures => deleted."""
from __future__ import annotations

import time
from collections.abc import Callable
from dataclasses import dataclass
from typing import Any

from infomesh_b200.crawler import MAX_RESPONSE_BYTES, create_ssl_context
from infomesh_b200.hashing import content_hash
from infomesh_b200.security import SSRFError, validate_url
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

INTERVAL_HIGH = 6 * 3600
INTERVAL_MEDIUM = 24 * 3600
INTERVAL_LOW = 7 * 24 * 3600
INTERVAL_STATIC = 30 * 24 * 3600
STALE_THRESHOLD = 3


@dataclass
class RecrawlCandidate:
    doc_id: int
    url: str
    text_hash: str
    etag: str | None
    last_modified: str | None
    recrawl_interval: int
    stale_count: int
    change_frequency: float
    crawled_at: float
    last_recrawl_at: float | None


@dataclass
class RecrawlOutcome:
    url: str
    status: str  # not_modified | updated | deleted | error
    new_text_hash: str | None = None
    new_etag: str | None = None
    new_last_modified: str | None = None
    stale_count: int = 0
    elapsed_ms: float = 0.0
    new_text: str | None = None


def compute_recrawl_interval(change_frequency: float) -> int:
    if change_frequency <= 0.0:
        return INTERVAL_STATIC
    if change_frequency < 0.10:
        return INTERVAL_LOW
    return INTERVAL_MEDIUM if change_frequency <= 0.50 else INTERVAL_HIGH


def update_change_frequency(old_freq: float, changed: bool, *, alpha: float = 0.3) -> float:
    return alpha * (1.0 if changed else 0.0) + (1.0 - alpha) * old_freq


async def recrawl_url(url: str, etag: str | None, last_modified: str | None, old_text_hash: str, stale_count: int, *,
                      client: Any | None = None, user_agent: str = "InfoMesh/0.1",
                      extract_fn: Callable[[str, str], str | None] | None = None) -> RecrawlOutcome:
    t0 = time.monotonic()
    ms = lambda: (time.monotonic() - t0) * 1000  # noqa: E731
    try:
        validate_url(url)
    except SSRFError as exc:
        logger.warning("recrawl_ssrf_blocked", url=url, reason=str(exc))
        return RecrawlOutcome(url, "error", stale_count=stale_count, elapsed_ms=ms())
    own = client is None
    if own:
        import httpx

        client = httpx.AsyncClient(headers={"User-Agent": user_agent}, follow_redirects=True, timeout=30.0,
                                   verify=create_ssl_context())
    try:
        cond: dict[str, str] = {}
        if etag:
            cond["If-None-Match"] = etag
        if last_modified:
            cond["If-Modified-Since"] = last_modified
        try:
            resp = await client.get(url, headers=cond, timeout=30.0)
        except Exception as exc:  # noqa: BLE001
            logger.warning("recrawl_network_error", url=url, error=str(exc))
            return RecrawlOutcome(url, "error", stale_count=stale_count + 1, elapsed_ms=ms())
        code = resp.status_code
        if code == 304:
            return RecrawlOutcome(url, "not_modified", new_etag=etag, new_last_modified=last_modified,
                                  stale_count=0, elapsed_ms=ms())
        if code >= 400:
            n = stale_count + 1
            return RecrawlOutcome(url, "deleted" if n >= STALE_THRESHOLD else "error", stale_count=n, elapsed_ms=ms())
        body = resp.text
        if len(body.encode("utf-8", errors="replace")) > MAX_RESPONSE_BYTES:
            return RecrawlOutcome(url, "error", stale_count=stale_count, elapsed_ms=ms())
        text = extract_fn(body, url) if extract_fn is not None else body
        if text is None:
            return RecrawlOutcome(url, "error", stale_count=stale_count, elapsed_ms=ms())
        new_hash = content_hash(text)
        base = dict(new_text_hash=new_hash, new_etag=resp.headers.get("etag"),
                    new_last_modified=resp.headers.get("last-modified"), stale_count=0, elapsed_ms=ms())
        if new_hash == old_text_hash:
            return RecrawlOutcome(url, "not_modified", **base)
        logger.info("recrawl_updated", url=url, old_hash=old_text_hash[:12], new_hash=new_hash[:12])
        return RecrawlOutcome(url, "updated", new_text=text, **base)
    finally:
        if own:
            await client.aclose()


def select_candidates(docs: list[RecrawlCandidate], *, now: float | None = None, max_batch: int = 50
                      ) -> list[RecrawlCandidate]:
    """Documents whose interval elapsed, most overdue first."""
    now = now or time.time()
    due = []
    for d in docs:
        if d.stale_count >= STALE_THRESHOLD:
            continue
        last = d.last_recrawl_at if d.last_recrawl_at is not None else d.crawled_at
        overdue = now - last - d.recrawl_interval
        if overdue >= 0:
            due.append((overdue, d))
    due.sort(key=lambda x: x[0], reverse=True)
    return [d for _, d in due[:max_batch]]


def apply_outcome(store: Any, cand: RecrawlCandidate, out: RecrawlOutcome, *, now: float | None = None) -> None:
    """Persist a recrawl outcome into ``LocalStore`` (interval / frequency / stale bookkeeping)."""
    now = now or time.time()
    if out.status == "deleted":
        store.soft_delete(cand.url)
        return
    changed = out.status == "updated"
    freq = update_change_frequency(cand.change_frequency, changed) if out.status != "error" else cand.change_frequency
    store.update_document(cand.url, text=out.new_text if changed else None,
                          text_hash=out.new_text_hash if changed else None, etag=out.new_etag,
                          last_modified=out.new_last_modified, recrawl_interval=compute_recrawl_interval(freq),
                          stale_count=out.stale_count, last_recrawl_at=now, change_frequency=freq)
