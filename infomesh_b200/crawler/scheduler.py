"""Crawl scheduler: bounded asyncio queue, per-domain politeness (robots Crawl-delay capped at 60 s), pending-per-domain
cap, hourly budget, stale-domain pruning (reference infomesh/crawler/scheduler.py:27-208)."""
from __future__ import annotations

import asyncio
import time
from collections import defaultdict
from dataclasses import dataclass
from urllib.parse import urlparse

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_MAX_TRACKED_DOMAINS = 50_000
_DOMAIN_PRUNE_THRESHOLD = int(_MAX_TRACKED_DOMAINS * 0.8)
_DOMAIN_STALE_SECONDS = 3600.0
_QUEUE_SIZE = 10_000
MAX_CRAWL_DELAY = 60.0


@dataclass
class DomainState:
    last_request_at: float = 0.0
    pending_count: int = 0
    error_count: int = 0
    crawl_delay: float | None = None


class Scheduler:
    def __init__(self, *, politeness_delay: float = 1.0, urls_per_hour: int = 60, pending_per_domain: int = 10,
                 max_depth: int = 0):
        self._delay = politeness_delay
        self._per_hour = urls_per_hour
        self._per_domain = pending_per_domain
        self._max_depth = max_depth
        self._domains: dict[str, DomainState] = defaultdict(DomainState)
        self._queue: asyncio.Queue[tuple[str, int]] = asyncio.Queue(maxsize=_QUEUE_SIZE)
        self._hour_count = 0
        self._hour_start = time.monotonic()

    async def add_url(self, url: str, depth: int = 0) -> bool:
        if self._max_depth > 0 and depth > self._max_depth:
            return False
        if len(self._domains) > _DOMAIN_PRUNE_THRESHOLD:
            self._prune(_DOMAIN_PRUNE_THRESHOLD)
        st = self._domains[urlparse(url).netloc]
        if st.pending_count >= self._per_domain or self._queue.full():
            return False
        st.pending_count += 1
        await self._queue.put((url, depth))
        return True

    def set_urls_per_hour(self, limit: int) -> None:
        """0 = unlimited."""
        self._per_hour = limit

    def set_crawl_delay(self, domain: str, delay: float) -> None:
        self._domains[domain].crawl_delay = min(float(delay), MAX_CRAWL_DELAY)

    async def get_url(self) -> tuple[str, int]:
        while True:
            url, depth = await self._queue.get()
            st = self._domains[urlparse(url).netloc]
            delay = st.crawl_delay if st.crawl_delay is not None else self._delay
            wait = delay - (time.monotonic() - st.last_request_at)
            if wait > 0:
                await asyncio.sleep(wait)
            if self._per_hour > 0:
                self._roll_hour()
                if self._hour_count >= self._per_hour:
                    remaining = max(3600 - (time.monotonic() - self._hour_start), 1.0)
                    logger.info("scheduler_hourly_limit", count=self._hour_count, wait_secs=round(remaining))
                    await self._queue.put((url, depth))
                    await asyncio.sleep(remaining)
                    continue
                self._hour_count += 1
            st.last_request_at = time.monotonic()
            return url, depth

    def mark_done(self, url: str) -> None:
        st = self._domains.get(urlparse(url).netloc)
        if st is None:
            return
        st.pending_count = max(0, st.pending_count - 1)
        if len(self._domains) > _DOMAIN_PRUNE_THRESHOLD:
            self._prune(_DOMAIN_PRUNE_THRESHOLD)

    def mark_error(self, url: str) -> None:
        st = self._domains.get(urlparse(url).netloc)
        if st is not None:
            st.error_count += 1
            self.mark_done(url)

    def _roll_hour(self) -> None:
        now = time.monotonic()
        if now - self._hour_start >= 3600:
            self._hour_count, self._hour_start = 0, now
            self._prune(_MAX_TRACKED_DOMAINS)

    def _prune(self, threshold: int) -> None:
        if len(self._domains) <= threshold:
            return
        cutoff = time.monotonic() - _DOMAIN_STALE_SECONDS
        for d in [d for d, s in self._domains.items() if s.pending_count == 0 and s.last_request_at < cutoff]:
            del self._domains[d]

    @property
    def pending_count(self) -> int:
        return self._queue.qsize()

    @property
    def tracked_domains(self) -> int:
        return len(self._domains)

    def domain_state(self, domain: str) -> DomainState | None:
        return self._domains.get(domain)
