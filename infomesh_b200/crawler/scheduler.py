"""Crawl frontier: which URL may be fetched next, and when.

Contract (SURVEY §2.1 crawler/ "scheduler"; reference infomesh/crawler/scheduler.py): a bounded FIFO of ``(url, depth)``;
at most ``pending_per_domain`` queued URLs per host; an optional depth limit; between two fetches of one host at least the
politeness delay -- or the host's robots ``Crawl-delay``, capped at 60 s -- must pass; at most ``urls_per_hour`` fetches
per rolling hour (0 = unlimited), a URL that hits the cap goes back to the queue while the caller sleeps out the hour;
bookkeeping for hosts that are idle for an hour is dropped once the table grows large.

Implementation: three small collaborators instead of one class doing everything -- ``_HostBook`` (per-host counters and
pruning), ``_HourlyAllowance`` (the rolling budget) and the queue; ``Scheduler`` only sequences them."""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass
from urllib.parse import urlsplit

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_MAX_TRACKED_DOMAINS = 50_000
_DOMAIN_PRUNE_THRESHOLD = _MAX_TRACKED_DOMAINS * 4 // 5
_DOMAIN_STALE_SECONDS = 3600.0
_QUEUE_SIZE = 10_000
_HOUR = 3600.0
MAX_CRAWL_DELAY = 60.0


@dataclass
class DomainState:
    last_request_at: float = 0.0
    pending_count: int = 0
    error_count: int = 0
    crawl_delay: float | None = None

    def idle_since(self, cutoff: float) -> bool:
        return self.pending_count == 0 and self.last_request_at < cutoff


def _host(url: str) -> str:
    return urlsplit(url).netloc


class _HostBook(dict):
    """host -> :class:`DomainState`, created on first touch."""

    def __missing__(self, host: str) -> DomainState:
        state = self[host] = DomainState()
        return state

    def shrink_to(self, limit: int) -> None:
        """Forget idle hosts when more than ``limit`` are tracked (hosts with queued URLs are never dropped)."""
        if len(self) <= limit:
            return
        cutoff = time.monotonic() - _DOMAIN_STALE_SECONDS
        for host in [h for h, st in self.items() if st.idle_since(cutoff)]:
            del self[host]


class _HourlyAllowance:
    """Fetches granted in the current hour window; ``limit == 0`` disables the cap."""

    def __init__(self, limit: int):
        self.limit = int(limit)
        self.used = 0
        self.window_opened = time.monotonic()

    def roll(self) -> bool:
        """Open a new window when the hour is over.  True if it rolled."""
        now = time.monotonic()
        if now - self.window_opened < _HOUR:
            return False
        self.used, self.window_opened = 0, now
        return True

    def exhausted(self) -> bool:
        return self.limit > 0 and self.used >= self.limit

    def seconds_left(self) -> float:
        return max(_HOUR - (time.monotonic() - self.window_opened), 1.0)


class Scheduler:
    def __init__(self, *, politeness_delay: float = 1.0, urls_per_hour: int = 60, pending_per_domain: int = 10, max_depth: int = 0):
        self._politeness = politeness_delay
        self._host_cap = pending_per_domain
        self._depth_cap = max_depth
        self._domains = _HostBook()
        self._budget = _HourlyAllowance(urls_per_hour)
        self._queue: asyncio.Queue[tuple[str, int]] = asyncio.Queue(maxsize=_QUEUE_SIZE)

    # ---- configuration
    @property
    def _per_hour(self) -> int:
        return self._budget.limit

    def set_urls_per_hour(self, limit: int) -> None:
        """0 = unlimited."""
        self._budget.limit = int(limit)

    def set_crawl_delay(self, domain: str, delay: float) -> None:
        self._domains[domain].crawl_delay = min(float(delay), MAX_CRAWL_DELAY)

    # ---- producers
    async def add_url(self, url: str, depth: int = 0) -> bool:
        if 0 < self._depth_cap < depth:
            return False
        self._domains.shrink_to(_DOMAIN_PRUNE_THRESHOLD)
        state = self._domains[_host(url)]
        if state.pending_count >= self._host_cap or self._queue.full():
            return False
        state.pending_count += 1
        await self._queue.put((url, depth))
        return True

    # ---- consumer
    async def get_url(self) -> tuple[str, int]:
        """Next ``(url, depth)``; returns only once the host's delay has passed and the hourly budget allows it."""
        while True:
            item = await self._queue.get()
            state = self._domains[_host(item[0])]
            gap = self._politeness if state.crawl_delay is None else state.crawl_delay
            pause = state.last_request_at + gap - time.monotonic()
            if pause > 0:
                await asyncio.sleep(pause)
            if self._budget.limit > 0:
                if self._budget.roll():
                    self._domains.shrink_to(_MAX_TRACKED_DOMAINS)
                if self._budget.exhausted():
                    nap = self._budget.seconds_left()
                    logger.info("scheduler_hourly_limit", count=self._budget.used, wait_secs=round(nap))
                    await self._queue.put(item)
                    await asyncio.sleep(nap)
                    continue
                self._budget.used += 1
            state.last_request_at = time.monotonic()
            return item

    # ---- completion
    def mark_done(self, url: str) -> None:
        state = self._domains.get(_host(url))
        if state is None:
            return
        state.pending_count = max(state.pending_count - 1, 0)
        self._domains.shrink_to(_DOMAIN_PRUNE_THRESHOLD)

    def mark_error(self, url: str) -> None:
        state = self._domains.get(_host(url))
        if state is None:
            return
        state.error_count += 1
        self.mark_done(url)

    def _prune(self, threshold: int) -> None:
        self._domains.shrink_to(threshold)

    # ---- introspection
    @property
    def pending_count(self) -> int:
        return self._queue.qsize()

    @property
    def tracked_domains(self) -> int:
        return len(self._domains)

    def domain_state(self, domain: str) -> DomainState | None:
        return self._domains.get(domain)
