"""Crawl "intelligence": shareable robots.txt verdict cache, load-adaptive crawl delay, image alt-text mining
(reference infomesh/crawler/intelligence.py:24-218)."""
from __future__ import annotations

import re
import time
from dataclasses import dataclass


@dataclass
class RobotsCacheEntry:
    domain: str
    allowed: bool
    crawl_delay: float
    sitemaps: list[str]
    cached_at: float
    expires_at: float


class RobotsCache:
    def __init__(self, ttl_seconds: float = 86400):
        self._ttl = ttl_seconds
        self._cache: dict[str, RobotsCacheEntry] = {}

    def get(self, domain: str) -> RobotsCacheEntry | None:
        e = self._cache.get(domain)
        if e is None:
            return None
        if time.time() < e.expires_at:
            return e
        del self._cache[domain]
        return None

    def put(self, domain: str, allowed: bool, crawl_delay: float = 0.0, sitemaps: list[str] | None = None
            ) -> RobotsCacheEntry:
        now = time.time()
        e = RobotsCacheEntry(domain, allowed, crawl_delay, sitemaps or [], now, now + self._ttl)
        self._cache[domain] = e
        return e

    def export_for_dht(self) -> list[dict[str, object]]:
        now = time.time()
        return [{"domain": e.domain, "allowed": e.allowed, "crawl_delay": e.crawl_delay, "sitemaps": e.sitemaps,
                 "cached_at": e.cached_at} for e in self._cache.values() if now < e.expires_at]

    def import_from_dht(self, entries: list[dict[str, object]]) -> int:
        n = 0
        for ent in entries:
            dom = str(ent.get("domain", ""))
            if not dom or dom in self._cache:
                continue
            cd, sm = ent.get("crawl_delay", 0), ent.get("sitemaps", [])
            try:
                delay = float(cd) if isinstance(cd, (int, float, str)) else 0.0
            except ValueError:
                delay = 0.0
            self.put(dom, bool(ent.get("allowed", True)), delay, [str(s) for s in sm] if isinstance(sm, list) else [])
            n += 1
        return n

    @property
    def size(self) -> int:
        return len(self._cache)

    def cleanup(self) -> int:
        now = time.time()
        dead = [k for k, v in self._cache.items() if now >= v.expires_at]
        for k in dead:
            del self._cache[k]
        return len(dead)


@dataclass
class CrawlTuningState:
    base_delay: float = 1.0
    current_delay: float = 1.0
    cpu_usage: float = 0.0
    memory_usage: float = 0.0
    adjustment_reason: str = ""


class CrawlSpeedTuner:
    """x1.5 above 90 % CPU/mem, x1.2 above 70 / 80 %, x0.8 when idle (< 30 % CPU and < 50 % mem)."""

    def __init__(self, base_delay: float = 1.0, min_delay: float = 0.2, max_delay: float = 10.0):
        self._base, self._min, self._max = base_delay, min_delay, max_delay
        self._current = base_delay

    def adjust(self, cpu: float | None = None, mem: float | None = None) -> CrawlTuningState:
        if cpu is None or mem is None:
            try:
                import psutil

                cpu = psutil.cpu_percent(interval=0.1) if cpu is None else cpu
                mem = psutil.virtual_memory().percent if mem is None else mem
            except ImportError:
                cpu, mem = cpu or 0.0, mem or 0.0
        if cpu > 90 or mem > 90:
            self._current, why = min(self._current * 1.5, self._max), f"high load (CPU={cpu:.0f}%, MEM={mem:.0f}%)"
        elif cpu > 70 or mem > 80:
            self._current, why = min(self._current * 1.2, self._max), f"moderate load (CPU={cpu:.0f}%, MEM={mem:.0f}%)"
        elif cpu < 30 and mem < 50:
            self._current, why = max(self._current * 0.8, self._min), f"low load (CPU={cpu:.0f}%, MEM={mem:.0f}%)"
        else:
            why = "stable"
        return CrawlTuningState(self._base, round(self._current, 2), cpu, mem, why)

    @property
    def current_delay(self) -> float:
        return self._current


_IMG_ALT = re.compile(r'<img\b[^>]*\balt=["\']([^"\']{3,200})["\']', re.I)
_PLACEHOLDER = frozenset({"image", "photo", "picture", "img", "icon", "logo", "banner", "thumbnail", "avatar"})


def extract_image_alt_texts(html: str) -> list[str]:
    out: list[str] = []
    seen: set[str] = set()
    for m in _IMG_ALT.finditer(html):
        alt = m.group(1).strip()
        if alt.lower() in _PLACEHOLDER or alt in seen:
            continue
        seen.add(alt)
        out.append(alt)
    return out
