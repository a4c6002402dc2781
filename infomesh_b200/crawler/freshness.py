"""Document freshness tiers, the priority recrawl queue and conditional-GET headers.

Contract (SURVEY §2.1 crawler/ "Freshness queue", reference infomesh/crawler/freshness.py): a page is hot for an hour,
warm for a day, cold for a week, stale afterwards; recrawl requests are served strictly by trigger class
(user > RSS > content change > peer announce > scheduled) and oldest-first inside a class, one entry per URL, bounded.

Implementation: the age boundaries are a sorted table searched with ``bisect``; the queue is one time-ordered lane per
trigger class plus a URL index -- a dequeue walks the lanes in class order, so there is no global heap to re-balance
and discarding a URL is a dictionary delete (its lane entry becomes a tombstone that the next dequeue drops)."""
from __future__ import annotations

import bisect
import itertools
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


_AGE_LIMITS = (TIER_HOT_MAX, TIER_WARM_MAX, TIER_COLD_MAX)
_TIER_BY_SLOT = (FreshnessTier.HOT, FreshnessTier.WARM, FreshnessTier.COLD, FreshnessTier.STALE)


def classify_freshness(crawled_at: float, *, now: float | None = None) -> FreshnessTier:
    """Tier of a page crawled at ``crawled_at`` (each limit is inclusive: exactly one hour old is still hot)."""
    age = (time.time() if now is None else now) - crawled_at
    return _TIER_BY_SLOT[bisect.bisect_left(_AGE_LIMITS, age)]


class RecrawlTrigger(StrEnum):
    RSS_UPDATE = "rss_update"
    USER_REQUEST = "user_request"
    CONTENT_CHANGE = "content_change"
    SCHEDULED = "scheduled"
    PEER_ANNOUNCE = "peer_announce"


# serving order of the trigger classes (lower = sooner)
TRIGGER_PRIORITY: dict[RecrawlTrigger, int] = {
    RecrawlTrigger.USER_REQUEST: 0, RecrawlTrigger.RSS_UPDATE: 1, RecrawlTrigger.CONTENT_CHANGE: 2,
    RecrawlTrigger.PEER_ANNOUNCE: 3, RecrawlTrigger.SCHEDULED: 4,
}
_LOWEST_CLASS = max(TRIGGER_PRIORITY.values())


@dataclass(frozen=True, order=True)
class PriorityRecrawlItem:
    priority: int
    enqueued_at: float
    url: str = field(compare=False)
    trigger: RecrawlTrigger = field(compare=False)
    source_feed: str = field(default="", compare=False)


class PriorityRecrawlQueue:
    """Bounded, de-duplicated recrawl queue: trigger class first, then age of the request."""

    def __init__(self, *, max_size: int = 10000):
        self._limit = int(max_size)
        self._lanes: list[list[tuple[float, int, PriorityRecrawlItem]]] = [[] for _ in range(_LOWEST_CLASS + 1)]
        self._by_url: dict[str, PriorityRecrawlItem] = {}
        self._ticket = itertools.count()          # tie-break for identical timestamps: arrival order
        self._served = 0
        self._accepted = 0

    # ---- producers
    def enqueue(self, url: str, trigger: RecrawlTrigger, *, source_feed: str = "", now: float | None = None) -> bool:
        if url in self._by_url:
            return False
        if len(self._by_url) >= self._limit:
            logger.warning("recrawl_queue_full", max_size=self._limit, url=url)
            return False
        lane = TRIGGER_PRIORITY.get(trigger, _LOWEST_CLASS)
        item = PriorityRecrawlItem(lane, time.time() if now is None else now, url, trigger, source_feed)
        bisect.insort(self._lanes[lane], (item.enqueued_at, next(self._ticket), item))
        self._by_url[url] = item
        self._accepted += 1
        return True

    def discard(self, url: str) -> None:
        self._by_url.pop(url, None)

    def clear(self) -> None:
        for lane in self._lanes:
            lane.clear()
        self._by_url.clear()

    # ---- consumer
    def _head(self):
        """(lane, entry) of the next live item, dropping tombstones on the way; ``None`` when empty."""
        for lane in self._lanes:
            while lane:
                entry = lane[0]
                if self._by_url.get(entry[2].url) is entry[2]:
                    return lane, entry
                lane.pop(0)
        return None

    def peek(self) -> PriorityRecrawlItem | None:
        head = self._head()
        return head[1][2] if head else None

    def dequeue(self) -> PriorityRecrawlItem | None:
        head = self._head()
        if head is None:
            return None
        lane, entry = head
        lane.pop(0)
        del self._by_url[entry[2].url]
        self._served += 1
        return entry[2]

    # ---- counters
    @property
    def size(self) -> int:
        return len(self._by_url)

    @property
    def total_enqueued(self) -> int:
        return self._accepted

    @property
    def total_dequeued(self) -> int:
        return self._served


@dataclass(frozen=True)
class ConditionalHeaders:
    """Validators remembered from the last fetch, replayed as a conditional request."""
    etag: str | None = None
    last_modified: str | None = None

    _REQUEST_NAME = {"etag": "If-None-Match", "last_modified": "If-Modified-Since"}

    def to_request_headers(self) -> dict[str, str]:
        return {header: value for attr, header in self._REQUEST_NAME.items() if (value := getattr(self, attr))}

    @staticmethod
    def from_response_headers(headers: dict[str, str]) -> "ConditionalHeaders":
        folded = {str(k).lower(): v for k, v in dict(headers).items()}
        return ConditionalHeaders(etag=folded.get("etag"), last_modified=folded.get("last-modified"))
