"""Freshness tiers, the trigger-priority recrawl queue and conditional-request headers
(reference infomesh/crawler/freshness.py:23-238)."""
from __future__ import annotations

import heapq
import time
from dataclasses import dataclass, field
from enum import StrEnum

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

TIER_HOT_MAX = 3600
TIER_WARM_MAX = 86400
TIER_COLD_MAX = 604800


class FreshnessTier(StrEnum):
    HOT = "hot"
    WARM = "warm"
    COLD = "cold"
    STALE = "stale"


def classify_freshness(crawled_at: float, *, now: float | None = None) -> FreshnessTier:
    age = (now or time.time()) - crawled_at
    if age <= TIER_HOT_MAX:
        return FreshnessTier.HOT
    if age <= TIER_WARM_MAX:
        return FreshnessTier.WARM
    return FreshnessTier.COLD if age <= TIER_COLD_MAX else FreshnessTier.STALE


class RecrawlTrigger(StrEnum):
    RSS_UPDATE = "rss_update"
    USER_REQUEST = "user_request"
    CONTENT_CHANGE = "content_change"
    SCHEDULED = "scheduled"
    PEER_ANNOUNCE = "peer_announce"


TRIGGER_PRIORITY: dict[RecrawlTrigger, int] = {
    RecrawlTrigger.USER_REQUEST: 0, RecrawlTrigger.RSS_UPDATE: 1, RecrawlTrigger.CONTENT_CHANGE: 2,
    RecrawlTrigger.PEER_ANNOUNCE: 3, RecrawlTrigger.SCHEDULED: 4,
}


@dataclass(frozen=True, order=True)
class PriorityRecrawlItem:
    priority: int
    enqueued_at: float
    url: str = field(compare=False)
    trigger: RecrawlTrigger = field(compare=False)
    source_feed: str = field(default="", compare=False)


class PriorityRecrawlQueue:
    """Min-heap on (trigger priority, enqueue time) with URL de-duplication and lazy deletion."""

    def __init__(self, *, max_size: int = 10000):
        self._heap: list[PriorityRecrawlItem] = []
        self._live: set[str] = set()
        self._max = max_size
        self._enq = self._deq = 0

    def enqueue(self, url: str, trigger: RecrawlTrigger, *, source_feed: str = "", now: float | None = None) -> bool:
        if url in self._live:
            return False
        if len(self._heap) >= self._max:
            logger.warning("recrawl_queue_full", max_size=self._max, url=url)
            return False
        heapq.heappush(self._heap, PriorityRecrawlItem(TRIGGER_PRIORITY.get(trigger, 4), now or time.time(), url,
                                                       trigger, source_feed))
        self._live.add(url)
        self._enq += 1
        return True

    def dequeue(self) -> PriorityRecrawlItem | None:
        while self._heap:
            item = heapq.heappop(self._heap)
            if item.url in self._live:
                self._live.discard(item.url)
                self._deq += 1
                return item
        return None

    def discard(self, url: str) -> None:
        self._live.discard(url)

    def peek(self) -> PriorityRecrawlItem | None:
        while self._heap and self._heap[0].url not in self._live:
            heapq.heappop(self._heap)
        return self._heap[0] if self._heap else None

    @property
    def size(self) -> int:
        return len(self._live)

    @property
    def total_enqueued(self) -> int:
        return self._enq

    @property
    def total_dequeued(self) -> int:
        return self._deq

    def clear(self) -> None:
        self._heap.clear()
        self._live.clear()


@dataclass(frozen=True)
class ConditionalHeaders:
    etag: str | None = None
    last_modified: str | None = None

    def to_request_headers(self) -> dict[str, str]:
        h: dict[str, str] = {}
        if self.etag:
            h["If-None-Match"] = self.etag
        if self.last_modified:
            h["If-Modified-Since"] = self.last_modified
        return h

    @staticmethod
    def from_response_headers(headers: dict[str, str]) -> "ConditionalHeaders":
        low = {k.lower(): v for k, v in dict(headers).items()}
        return ConditionalHeaders(low.get("etag"), low.get("last-modified"))
