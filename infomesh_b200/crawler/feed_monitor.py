"""Watching RSS / Atom feeds for new URLs.  No I/O here: the caller fetches the XML and hands it in.

Contract (SURVEY §2.1 crawler/ "feed monitor"; reference infomesh/crawler/feed_monitor.py): feeds carry a priority that
decides how often they are polled (critical 1 min, high 5 min, normal 15 min, low 60 min) unless a per-feed interval
overrides it; a feed never polled is due immediately; due feeds come back most-urgent priority first and, inside a
priority, most overdue first; a URL is reported as new once per monitor, across feeds; OPML subscription lists can be
imported (``<outline xmlUrl=... text=...|title=...>``), duplicates skipped; an optional cap on the number of feeds.

Implementation: OPML goes through a tolerant tag parser (real-world exports are rarely well-formed XML); each feed knows
its own lateness (``lateness(now)``), so "due" is a filter + one sort key; counters are kept in a ``Counter`` and projected
into the public stats record on demand."""
from __future__ import annotations

import time
from collections import Counter
from dataclasses import dataclass, field
from enum import StrEnum
from html.parser import HTMLParser

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class FeedPriority(StrEnum):
    CRITICAL = "critical"
    HIGH = "high"
    NORMAL = "normal"
    LOW = "low"


POLL_INTERVALS: dict[FeedPriority, int] = {FeedPriority.CRITICAL: 60, FeedPriority.HIGH: 5 * 60,
                                           FeedPriority.NORMAL: 15 * 60, FeedPriority.LOW: 60 * 60}
_URGENCY = {p: rank for rank, p in enumerate(FeedPriority)}        # declaration order == urgency order


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
        return self.poll_interval or POLL_INTERVALS[self.priority]

    def lateness(self, now: float) -> float | None:
        """Seconds past the poll deadline (>= 0), ``None`` while not due.  A feed never polled is due with lateness 0."""
        if self.last_poll_at == 0.0:
            return 0.0
        late = now - self.last_poll_at - self.effective_interval
        return late if late >= 0 else None


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


# ----------------------------------------------------------------------------- OPML
class _OutlineReader(HTMLParser):
    """Collects ``(xmlUrl, label)`` of every ``<outline>`` element, in document order."""

    def __init__(self):
        super().__init__(convert_charrefs=True)
        self.subscriptions: dict[str, str] = {}

    def handle_starttag(self, tag, attrs):
        if tag != "outline":
            return
        a = {k.lower(): (v or "") for k, v in attrs}       # html.parser lower-cases names; be explicit anyway
        target = a.get("xmlurl", "").strip()
        if target and target not in self.subscriptions:
            self.subscriptions[target] = (a.get("text") or a.get("title") or "").strip()

    handle_startendtag = handle_starttag


def parse_opml(opml_text: str) -> list[MonitoredFeed]:
    reader = _OutlineReader()
    try:
        reader.feed(opml_text)
        reader.close()
    except Exception:  # noqa: BLE001 -- keep whatever was readable
        pass
    feeds = [MonitoredFeed(url=u, label=lab) for u, lab in reader.subscriptions.items()]
    logger.info("opml_parsed", feed_count=len(feeds))
    return feeds


# ----------------------------------------------------------------------------- monitor
class FeedMonitor:
    def __init__(self, max_feeds: int = 0):
        self._by_url: dict[str, MonitoredFeed] = {}
        self._reported: set[str] = set()          # item URLs already handed out (or marked seen by the crawler)
        self._tally: Counter[str] = Counter()
        self._capacity = max_feeds

    # ---- subscriptions
    def add_feed(self, url: str, *, priority: FeedPriority = FeedPriority.NORMAL, poll_interval: int = 0, label: str = "") -> MonitoredFeed:
        known = self._by_url.get(url)
        if known is None:
            if self._capacity and len(self._by_url) >= self._capacity:
                raise ValueError(f"feed limit reached ({self._capacity})")
            known = self._by_url[url] = MonitoredFeed(url=url, priority=priority, poll_interval=poll_interval, label=label)
            return known
        known.priority = priority                 # re-adding updates the subscription in place
        known.poll_interval = poll_interval if poll_interval > 0 else known.poll_interval
        known.label = label or known.label
        return known

    def remove_feed(self, url: str) -> bool:
        return self._by_url.pop(url, None) is not None

    def add_feeds_from_opml(self, opml_text: str) -> int:
        fresh = [f for f in parse_opml(opml_text) if f.url not in self._by_url]
        for f in fresh:
            self.add_feed(f.url, priority=f.priority, label=f.label)
        return len(fresh)

    @property
    def feeds(self) -> list[MonitoredFeed]:
        return list(self._by_url.values())

    # ---- polling
    def get_due_feeds(self, *, now: float | None = None) -> list[MonitoredFeed]:
        clock = now or time.time()
        due = [(f, late) for f in self._by_url.values() if (late := f.lateness(clock)) is not None]
        due.sort(key=lambda pair: (_URGENCY[pair[0].priority], -pair[1]))
        return [f for f, _ in due]

    def process_feed_response(self, feed_url: str, xml_text: str, *, now: float | None = None) -> FeedUpdate:
        from infomesh_b200.crawler.rss import parse_feed_xml

        started = time.monotonic()
        feed = self._by_url.get(feed_url)
        if feed is None:
            return FeedUpdate(feed_url, error="feed not registered")
        feed.last_poll_at = now or time.time()

        def took() -> float:
            return (time.monotonic() - started) * 1000

        try:
            items = parse_feed_xml(xml_text, feed_url).items
        except Exception as exc:  # noqa: BLE001 -- a broken feed is data, not a crash
            feed.error_count += 1
            self._tally["errors"] += 1
            return FeedUpdate(feed_url, error=str(exc), poll_elapsed_ms=took())
        links = [it.url for it in items if it.url]
        novel = list(dict.fromkeys(u for u in links if u not in self._reported))
        self._reported.update(novel)
        feed.error_count = 0
        feed.items_discovered += len(novel)
        if items:
            feed.last_item_url = items[0].url
        self._tally["polls"] += 1
        self._tally["new_urls"] += len(novel)
        return FeedUpdate(feed_url, novel, took())

    def mark_url_seen(self, url: str) -> None:
        self._reported.add(url)

    # ---- counters
    @property
    def stats(self) -> FeedMonitorStats:
        return FeedMonitorStats(total_feeds=len(self._by_url), total_polls=self._tally["polls"], total_new_urls=self._tally["new_urls"],
                                total_errors=self._tally["errors"], feeds_by_priority=dict(Counter(f.priority.value for f in self._by_url.values())))
