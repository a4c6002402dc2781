"""robots.txt compliance with per-domain cache (TTL 1 h, 10 k domains), Crawl-delay and Sitemap extraction, and
per-domain locks so concurrent workers fetch a robots file once (reference infomesh/crawler/robots.py:32-184).
Unreachable / non-200 robots.txt => allowed (RFC 9309 §2.4)."""
from __future__ import annotations

import asyncio
import re
import time
from urllib.parse import urlparse
from urllib.robotparser import RobotFileParser

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_SITEMAP_RE = re.compile(r"^\s*sitemap:\s*(\S+)", re.IGNORECASE | re.MULTILINE)
_CRAWL_DELAY_RE = re.compile(r"^\s*crawl-delay:\s*(\d+(?:\.\d+)?)", re.IGNORECASE | re.MULTILINE)


class RobotsChecker:
    MAX_CACHE_SIZE: int = 10_000

    def __init__(self, user_agent: str, *, cache_ttl: int = 3600):
        self._ua = user_agent
        self._ttl = cache_ttl
        self._cache: dict[str, tuple[RobotFileParser, float]] = {}
        self._locks: dict[str, asyncio.Lock] = {}
        self._sitemaps: dict[str, list[str]] = {}
        self._delays: dict[str, float | None] = {}

    @staticmethod
    def parse_robots(text: str, robots_url: str = "") -> tuple[RobotFileParser, list[str], float | None]:
        """Parse robots.txt text -> (parser, sitemap urls, first Crawl-delay)."""
        parser = RobotFileParser()
        if robots_url:
            parser.set_url(robots_url)
        parser.parse(text.splitlines())
        m = _CRAWL_DELAY_RE.search(text)
        return parser, _SITEMAP_RE.findall(text), (float(m.group(1)) if m else None)

    async def _fetch(self, client, base_url: str) -> tuple[RobotFileParser, list[str], float | None]:
        robots_url = f"{base_url}/robots.txt"
        try:
            resp = await client.get(robots_url, timeout=10.0, follow_redirects=True)
            if resp.status_code == 200:
                return self.parse_robots(resp.text, robots_url)
            logger.debug("robots_not_found", url=robots_url, status=resp.status_code)
        except Exception as exc:  # noqa: BLE001 — network / TLS / decode errors are all "unreachable"
            logger.warning("robots_fetch_error", url=robots_url, error=str(exc))
        return self.parse_robots("", robots_url)

    async def is_allowed(self, client, url: str) -> bool:
        p = urlparse(url)
        domain = p.netloc
        lock = self._locks.setdefault(domain, asyncio.Lock())
        async with lock:
            hit = self._cache.get(domain)
            if hit is not None and time.monotonic() - hit[1] < self._ttl:
                return hit[0].can_fetch(self._ua, url)
            parser, sitemaps, delay = await self._fetch(client, f"{p.scheme}://{p.netloc}")
            if len(self._cache) >= self.MAX_CACHE_SIZE:
                self._evict_oldest()
            self._cache[domain] = (parser, time.monotonic())
            self._sitemaps[domain] = sitemaps
            self._delays[domain] = delay
        return parser.can_fetch(self._ua, url)

    def prime(self, domain: str, robots_text: str) -> None:
        """Install a robots.txt body received out of band (shared robots cache, tests)."""
        parser, sitemaps, delay = self.parse_robots(robots_text)
        self._cache[domain] = (parser, time.monotonic())
        self._sitemaps[domain] = sitemaps
        self._delays[domain] = delay

    def get_sitemaps(self, domain: str) -> list[str]:
        return self._sitemaps.get(domain, [])

    def get_crawl_delay(self, domain: str) -> float | None:
        return self._delays.get(domain)

    def _evict_oldest(self) -> None:
        victims = sorted(self._cache.items(), key=lambda kv: kv[1][1])[:max(1, len(self._cache) // 10)]
        for dom, _ in victims:
            for table in (self._cache, self._locks, self._sitemaps, self._delays):
                table.pop(dom, None)

    def clear_cache(self) -> None:
        self._cache.clear()
        self._sitemaps.clear()
        self._delays.clear()
