"""RSS/Atom feed monitoring: OPML import, priority-based poll intervals (1 / 5 / 15 / 60 min), new-URL detection
(reference infomesh/crawler/feed_monitor.py:25-342).  I/O free: the caller fetches the XML."""
from __future__ import annotations

import re
import time
from dataclasses import dataclass, field
from enum import StrEnum

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class FeedPriority(StrEnum):
    CRITICAL = "critical"
    HIGH = "high"
    NORMAL = "normal"
    LOW = "low"


POLL_INTERVALS: dict[FeedPriority, int] = {FeedPriority.CRITICAL: 60, FeedPriority.HIGH: 300,
                                           FeedPriority.NORMAL: 900, FeedPriority.LOW: 3600}
_ORDER = {FeedPriority.CRITICAL: 0, FeedPriority.HIGH: 1, FeedPriority.NORMAL: 2, FeedPriority.LOW: 3}


@dataclass
class MonitoredFeed:
    url: str
    priority: FeedPriority = FeedPriority.NORMAL
    poll_interval: int = 0
    last_poll_at: float = 0.0
    last_item_url: str = ""
    error_count: int = 0
    items_discovered: int = 0
    label: str = ""

    @property
    def effective_interval(self) -> int:
        return self.poll_interval if self.poll_interval > 0 else POLL_INTERVALS[self.priority]


@dataclass
class FeedUpdate:
    feed_url: str
    new_urls: list[str] = field(default_factory=list)
    poll_elapsed_ms: float = 0.0
    error: str | None = None


@dataclass
class FeedMonitorStats:
    total_feeds: int = 0
    total_polls: int = 0
    total_new_urls: int = 0
    total_errors: int = 0
    feeds_by_priority: dict[str, int] = field(default_factory=dict)


_OUTLINE = re.compile(r"<outline\b([^>]*)/?>", re.I)
_ATTR = {k: re.compile(rf'{k}=["\']([^"\']*)["\']', re.I) for k in ("xmlUrl", "text", "title")}


def parse_opml(opml_text: str) -> list[MonitoredFeed]:
    feeds: list[MonitoredFeed] = []
    seen: set[str] = set()
    for m in _OUTLINE.finditer(opml_text):
        attrs = m.group(1)
        u = _ATTR["xmlUrl"].search(attrs)
        if not u:
            continue
        url = u.group(1).strip()
        if not url or url in seen:
            continue
        seen.add(url)
        lab = _ATTR["text"].search(attrs) or _ATTR["title"].search(attrs)
        feeds.append(MonitoredFeed(url=url, label=lab.group(1).strip() if lab else ""))
    logger.info("opml_parsed", feed_count=len(feeds))
    return feeds


class FeedMonitor:
    def __init__(self, max_feeds: int = 0):
        self._feeds: dict[str, MonitoredFeed] = {}
        self._seen: set[str] = set()
        self._stats = FeedMonitorStats()
        self._max = max_feeds

    def add_feed(self, url: str, *, priority: FeedPriority = FeedPriority.NORMAL, poll_interval: int = 0,
                 label: str = "") -> MonitoredFeed:
        cur = self._feeds.get(url)
        if cur is not None:
            cur.priority = priority
            if poll_interval > 0:
                cur.poll_interval = poll_interval
            if label:
                cur.label = label
            return cur
        if self._max and len(self._feeds) >= self._max:
            raise ValueError(f"feed limit reached ({self._max})")
        feed = MonitoredFeed(url=url, priority=priority, poll_interval=poll_interval, label=label)
        self._feeds[url] = feed
        self._stats.total_feeds = len(self._feeds)
        return feed

    def remove_feed(self, url: str) -> bool:
        gone = self._feeds.pop(url, None) is not None
        self._stats.total_feeds = len(self._feeds)
        return gone

    def add_feeds_from_opml(self, opml_text: str) -> int:
        n = 0
        for f in parse_opml(opml_text):
            if f.url not in self._feeds:
                self.add_feed(f.url, priority=f.priority, label=f.label)
                n += 1
        return n

    def get_due_feeds(self, *, now: float | None = None) -> list[MonitoredFeed]:
        """Never-polled feeds are always due; ordering = priority, then most overdue."""
        now = now or time.time()
        due = []
        for f in self._feeds.values():
            if f.last_poll_at == 0.0:
                due.append((_ORDER[f.priority], 0.0, f))
            elif now - f.last_poll_at >= f.effective_interval:
                due.append((_ORDER[f.priority], -(now - f.last_poll_at - f.effective_interval), f))
        due.sort(key=lambda t: (t[0], t[1]))
        return [f for *_, f in due]

    def process_feed_response(self, feed_url: str, xml_text: str, *, now: float | None = None) -> FeedUpdate:
        from infomesh_b200.crawler.rss import parse_feed_xml

        now = now or time.time()
        t0 = time.monotonic()
        feed = self._feeds.get(feed_url)
        if feed is None:
            return FeedUpdate(feed_url, error="feed not registered")
        try:
            parsed = parse_feed_xml(xml_text, feed_url)
        except Exception as exc:  # noqa: BLE001
            feed.error_count += 1
            feed.last_poll_at = now
            self._stats.total_errors += 1
            return FeedUpdate(feed_url, error=str(exc), poll_elapsed_ms=(time.monotonic() - t0) * 1000)
        fresh = []
        for it in parsed.items:
            if it.url and it.url not in self._seen:
                self._seen.add(it.url)
                fresh.append(it.url)
        feed.last_poll_at = now
        feed.error_count = 0
        feed.items_discovered += len(fresh)
        if parsed.items:
            feed.last_item_url = parsed.items[0].url
        self._stats.total_polls += 1
        self._stats.total_new_urls += len(fresh)
        return FeedUpdate(feed_url, fresh, (time.monotonic() - t0) * 1000)

    def mark_url_seen(self, url: str) -> None:
        self._seen.add(url)

    @property
    def stats(self) -> FeedMonitorStats:
        by: dict[str, int] = {}
        for f in self._feeds.values():
            by[f.priority.value] = by.get(f.priority.value, 0) + 1
        self._stats.feeds_by_priority = by
        return self._stats

    @property
    def feeds(self) -> list[MonitoredFeed]:
        return list(self._feeds.values())
