"""Crawler side-knowledge: a shareable robots.txt verdict cache, a load-adaptive politeness delay and image alt-text mining.

Contract (SURVEY §2.1 crawler/ "intelligence"; reference infomesh/crawler/intelligence.py): robots verdicts live one day
and can be exported to / imported from the DHT (never overwriting what this node fetched itself); the delay tuner backs off
x1.5 under heavy load (> 90 % CPU or memory), x1.2 under moderate load (> 70 % CPU or > 80 % memory), speeds up x0.8 when
idle (< 30 % CPU and < 50 % memory) and stays inside ``[min_delay, max_delay]``; alt texts of 4-200 characters are kept once
each, generic placeholders ("logo", "icon", ...) dropped.

Implementation: verdicts are immutable records in a TTL map; the tuner is a table of load bands evaluated top-down; alt
texts come out of a real HTML tokenizer (``html.parser``) rather than a pattern over the raw markup, so attribute order,
quoting style and entities do not matter."""
from __future__ import annotations

import time
from dataclasses import asdict, dataclass, field
from html.parser import HTMLParser
from typing import Callable


# ----------------------------------------------------------------------------- robots verdict cache
@dataclass(frozen=True)
class RobotsCacheEntry:
    domain: str
    allowed: bool
    crawl_delay: float
    sitemaps: list[str]
    cached_at: float
    expires_at: float

    def live(self, now: float) -> bool:
        return now < self.expires_at

    def wire(self) -> dict[str, object]:
        """DHT form: everything but the local expiry (the importer applies its own TTL)."""
        rec = asdict(self)
        rec.pop("expires_at")
        return rec


def _as_delay(value: object) -> float:
    if isinstance(value, bool) or not isinstance(value, (int, float, str)):
        return 0.0
    try:
        return float(value)
    except ValueError:
        return 0.0


class RobotsCache:
    """Domain -> robots verdict with a time-to-live; expired verdicts vanish on access."""

    def __init__(self, ttl_seconds: float = 86400):
        self._ttl = float(ttl_seconds)
        self._by_domain: dict[str, RobotsCacheEntry] = {}

    def put(self, domain: str, allowed: bool, crawl_delay: float = 0.0, sitemaps: list[str] | None = None) -> RobotsCacheEntry:
        stamp = time.time()
        entry = RobotsCacheEntry(domain, allowed, crawl_delay, list(sitemaps or ()), stamp, stamp + self._ttl)
        self._by_domain[domain] = entry
        return entry

    def get(self, domain: str) -> RobotsCacheEntry | None:
        entry = self._by_domain.get(domain)
        if entry is not None and not entry.live(time.time()):
            self._by_domain.pop(domain, None)
            return None
        return entry

    def _live_entries(self) -> list[RobotsCacheEntry]:
        now = time.time()
        return [e for e in self._by_domain.values() if e.live(now)]

    def export_for_dht(self) -> list[dict[str, object]]:
        return [e.wire() for e in self._live_entries()]

    def import_from_dht(self, entries: list[dict[str, object]]) -> int:
        """Adopt peers' verdicts for domains this node knows nothing about.  Returns how many were taken."""
        taken = 0
        for rec in entries:
            domain = str(rec.get("domain", ""))
            if not domain or domain in self._by_domain:
                continue
            maps = rec.get("sitemaps", [])
            self.put(domain, bool(rec.get("allowed", True)), _as_delay(rec.get("crawl_delay", 0)),
                     [str(m) for m in maps] if isinstance(maps, list) else [])
            taken += 1
        return taken

    def cleanup(self) -> int:
        before = len(self._by_domain)
        self._by_domain = {e.domain: e for e in self._live_entries()}
        return before - len(self._by_domain)

    @property
    def size(self) -> int:
        return len(self._by_domain)


# ----------------------------------------------------------------------------- politeness delay tuner
@dataclass
class CrawlTuningState:
    base_delay: float = 1.0
    current_delay: float = 1.0
    cpu_usage: float = 0.0
    memory_usage: float = 0.0
    adjustment_reason: str = ""


@dataclass(frozen=True)
class _LoadBand:
    label: str
    factor: float
    applies: Callable[[float, float], bool] = field(compare=False)


# evaluated in order; the first band that applies scales the delay
_BANDS: tuple[_LoadBand, ...] = (
    _LoadBand("high load", 1.5, lambda cpu, mem: cpu > 90 or mem > 90),
    _LoadBand("moderate load", 1.2, lambda cpu, mem: cpu > 70 or mem > 80),
    _LoadBand("low load", 0.8, lambda cpu, mem: cpu < 30 and mem < 50),
)


def _probe_load(cpu: float | None, mem: float | None) -> tuple[float, float]:
    if cpu is not None and mem is not None:
        return cpu, mem
    try:
        import psutil
    except ImportError:
        return cpu or 0.0, mem or 0.0
    return (psutil.cpu_percent(interval=0.1) if cpu is None else cpu,
            psutil.virtual_memory().percent if mem is None else mem)


class CrawlSpeedTuner:
    def __init__(self, base_delay: float = 1.0, min_delay: float = 0.2, max_delay: float = 10.0):
        self._base, self._floor, self._ceiling = base_delay, min_delay, max_delay
        self._delay = base_delay

    @property
    def current_delay(self) -> float:
        return self._delay

    def adjust(self, cpu: float | None = None, mem: float | None = None) -> CrawlTuningState:
        cpu, mem = _probe_load(cpu, mem)
        reason = "stable"
        for band in _BANDS:
            if band.applies(cpu, mem):
                self._delay = min(self._ceiling, max(self._floor, self._delay * band.factor))
                reason = f"{band.label} (CPU={cpu:.0f}%, MEM={mem:.0f}%)"
                break
        return CrawlTuningState(self._base, round(self._delay, 2), cpu, mem, reason)


# ----------------------------------------------------------------------------- image alt texts
_GENERIC_ALTS = frozenset({"image", "photo", "picture", "img", "icon", "logo", "banner", "thumbnail", "avatar"})
_ALT_MIN, _ALT_MAX = 4, 200


class _AltCollector(HTMLParser):
    def __init__(self):
        super().__init__(convert_charrefs=True)
        self.found: dict[str, None] = {}          # insertion-ordered set

    def handle_starttag(self, tag, attrs):
        if tag != "img":
            return
        for name, value in attrs:
            if name == "alt" and value:
                text = value.strip()
                if _ALT_MIN <= len(value) <= _ALT_MAX and text.lower() not in _GENERIC_ALTS:
                    self.found.setdefault(text)
                return

    handle_startendtag = handle_starttag


def extract_image_alt_texts(html: str) -> list[str]:
    """Descriptive ``alt`` attributes of ``<img>`` tags, document order, no repeats."""
    parser = _AltCollector()
    try:
        parser.feed(html)
        parser.close()
    except Exception:  # noqa: BLE001 -- malformed markup: keep what was collected
        pass
    return list(parser.found)
