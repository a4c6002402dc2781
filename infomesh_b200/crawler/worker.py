"""CrawlWorker — fetch one URL through the full ingest pipeline (reference infomesh/crawler/worker.py:126-617):

    DHT crawl lock -> SSRF validation (with DNS) -> URL dedup -> robots.txt (+ Crawl-delay, sitemaps)
    -> GET with 2 retries (1 s, 2 s) on 5xx / network errors, redirect re-validation
    -> content-type / size (10 MiB) checks -> extraction -> JS detection / optional render
    -> canonical redirect -> exact-hash dedup -> SimHash near-dup -> mark seen
    -> BFS link scheduling within scope -> feed discovery

Every rejection returns a ``CrawlResult(success=False, error=<reason code>)`` with the reference's reason strings.
The HTTP client is injectable (``client_factory``) so the pipeline is testable without a network.
"""
from __future__ import annotations

import asyncio
import re
import time
from collections.abc import Callable
from dataclasses import dataclass, field
from typing import Any
from urllib.parse import urlparse

from infomesh_b200.config import CrawlConfig
from infomesh_b200.crawler import MAX_RESPONSE_BYTES, create_ssl_context
from infomesh_b200.crawler.dedup import DeduplicatorDB
from infomesh_b200.crawler.js_detect import detect_js_requirement
from infomesh_b200.crawler.parser import ParsedPage, extract_canonical, extract_content, extract_links
from infomesh_b200.crawler.robots import RobotsChecker
from infomesh_b200.crawler.scheduler import Scheduler
from infomesh_b200.hashing import content_hash
from infomesh_b200.security import SSRFError, validate_url, validate_url_post_redirect
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_MAX_RETRIES = 2
_RETRY_BACKOFF_BASE = 1.0
_SITEMAP_LOC = re.compile(r"<loc>\s*(https?://[^<]+?)\s*</loc>", re.I)
_MAX_SITEMAP_URLS = 500
_JS_RENDER_TEXT_FLOOR = 200


@dataclass
class CrawlResult:
    url: str
    success: bool
    page: ParsedPage | None = None
    error: str | None = None
    elapsed_ms: float = 0.0
    discovered_links: list[str] = field(default_factory=list)
    discovered_feeds: list[str] = field(default_factory=list)
    js_required: bool = False
    js_rendered: bool = False


def _ms(t0: float) -> float:
    return (time.monotonic() - t0) * 1000


class CrawlWorker:
    def __init__(self, config: CrawlConfig, scheduler: Scheduler, dedup: DeduplicatorDB, robots: RobotsChecker, *,
                 dht: Any | None = None, js_renderer: Any | None = None,
                 client_factory: Callable[[], Any] | None = None, resolve_dns: bool = True):
        self._config = config
        self._scheduler = scheduler
        self._dedup = dedup
        self._robots = robots
        self._dht = dht
        self._js = js_renderer
        self._client_factory = client_factory
        self._resolve_dns = resolve_dns
        self._client: Any | None = None
        self._scope_domain: str | None = None
        self._scope_path: str | None = None

    # ------------------------------------------------------------------ scope
    def set_scope(self, url: str) -> None:
        """Restrict link following to the directory tree of ``url``."""
        p = urlparse(url)
        self._scope_domain = p.netloc
        self._scope_path = p.path.rstrip("/") or "/"

    def clear_scope(self) -> None:
        self._scope_domain = self._scope_path = None

    def _in_scope(self, link: str) -> bool:
        if self._scope_domain is None:
            return True
        p = urlparse(link)
        if p.netloc != self._scope_domain:
            return False
        return True if self._scope_path in (None, "/") else p.path.startswith(self._scope_path)

    # ------------------------------------------------------------------ http client
    async def _get_client(self):
        if self._client is None or getattr(self._client, "is_closed", False):
            if self._client_factory is not None:
                self._client = self._client_factory()
            else:
                import httpx

                n = self._config.max_concurrent
                self._client = httpx.AsyncClient(headers={"User-Agent": self._config.user_agent},
                                                 follow_redirects=True, timeout=30.0, verify=create_ssl_context(),
                                                 limits=httpx.Limits(max_connections=n, max_keepalive_connections=n))
        return self._client

    async def get_http_client(self):
        return await self._get_client()

    async def close(self) -> None:
        if self._client is not None and hasattr(self._client, "aclose"):
            try:
                await self._client.aclose()
            except Exception:  # noqa: BLE001
                pass
        self._client = None

    # ------------------------------------------------------------------ entry point
    async def crawl_url(self, url: str, depth: int = 0, *, force: bool = False) -> CrawlResult:
        t0 = time.monotonic()
        locked = False
        try:
            if self._dht is not None:
                try:
                    locked = await self._dht.acquire_crawl_lock(url)
                except Exception:  # noqa: BLE001
                    logger.debug("crawl_lock_attempt_failed", url=url)
                if not locked:
                    return CrawlResult(url, False, error="locked_by_peer", elapsed_ms=_ms(t0))
            return await self._pipeline(url, depth, t0, force)
        finally:
            if locked:
                try:
                    await self._dht.release_crawl_lock(url)
                except Exception:  # noqa: BLE001
                    logger.debug("crawl_lock_release_failed", url=url)
            self._scheduler.mark_done(url)

    def _fail(self, url: str, t0: float, error: str, **kw: Any) -> CrawlResult:
        return CrawlResult(url, False, error=error, elapsed_ms=_ms(t0), **kw)

    async def _pipeline(self, url: str, depth: int, t0: float, force: bool) -> CrawlResult:
        try:
            validate_url(url, resolve_dns=self._resolve_dns)
        except SSRFError as exc:
            logger.warning("crawl_ssrf_blocked", url=url, reason=str(exc))
            return self._fail(url, t0, f"blocked: {exc}")
        if not force and self._dedup.is_url_seen(url):
            return self._fail(url, t0, "already_seen")
        client = await self._get_client()
        if self._config.respect_robots:
            if not await self._robots.is_allowed(client, url):
                logger.info("crawl_blocked_robots", url=url)
                return self._fail(url, t0, "blocked_by_robots")
            domain = urlparse(url).netloc
            delay = self._robots.get_crawl_delay(domain)
            if delay is not None:
                self._scheduler.set_crawl_delay(domain, delay)
            await self._schedule_sitemap_urls(domain)

        resp = await self._fetch_with_retry(client, url, t0)
        if isinstance(resp, CrawlResult):
            return resp
        ctype = resp.headers.get("content-type", "")
        if "text/html" not in ctype and "text/plain" not in ctype:
            return self._fail(url, t0, f"unsupported_content_type: {ctype}")
        declared = resp.headers.get("content-length")
        if declared is not None and declared.isdigit() and int(declared) > MAX_RESPONSE_BYTES:
            return self._fail(url, t0, "response_too_large")
        html = resp.text
        if len(html.encode("utf-8", errors="replace")) > MAX_RESPONSE_BYTES:
            return self._fail(url, t0, "response_too_large")

        raw_hash = content_hash(html)
        page = extract_content(html, url, raw_hash=raw_hash)
        js_required = detect_js_requirement(html).js_required
        js_rendered = False
        if js_required and (page is None or len(page.text) < _JS_RENDER_TEXT_FLOOR):
            rendered = await self._try_js_render(url)
            if rendered:
                js_rendered = True
                html = rendered
                page = extract_content(html, url, raw_hash=content_hash(html))
        if page is None:
            return self._fail(url, t0, "extraction_failed", js_required=js_required)

        canonical = extract_canonical(html, url)
        if canonical and canonical not in (url, url.rstrip("/")):
            self._dedup.mark_seen(url, page.text_hash, page.text)
            if not self._dedup.is_url_seen(canonical):
                await self._scheduler.add_url(canonical, depth=depth)
            return self._fail(url, t0, f"canonical_redirect:{canonical}")
        if not force:
            if self._dedup.is_content_seen(page.text_hash):
                self._dedup.mark_seen(url, page.text_hash, page.text)
                return self._fail(url, t0, "duplicate_content")
            if self._dedup.is_near_duplicate(page.text):
                self._dedup.mark_seen(url, page.text_hash, page.text)
                return self._fail(url, t0, "near_duplicate")
        self._dedup.mark_seen(url, page.text_hash, page.text)

        links: list[str] = []
        if self._config.max_depth == 0 or depth < self._config.max_depth:
            links = extract_links(html, url)
            scheduled = 0
            for link in links:
                if self._in_scope(link) and not self._dedup.is_url_seen(link):
                    scheduled += 1 if await self._scheduler.add_url(link, depth=depth + 1) else 0
            if scheduled:
                logger.info("links_scheduled", url=url, discovered=len(links), scheduled=scheduled,
                            next_depth=depth + 1)
        feeds: list[str] = []
        try:
            from infomesh_b200.crawler.rss import discover_feeds

            feeds = discover_feeds(html, url)
        except Exception:  # noqa: BLE001 — feed discovery is best effort
            pass
        elapsed = _ms(t0)
        logger.info("crawl_success", url=url, title=(page.title or "")[:60], text_len=len(page.text),
                    elapsed_ms=round(elapsed, 1), js_rendered=js_rendered)
        return CrawlResult(url, True, page=page, elapsed_ms=elapsed, discovered_links=links, discovered_feeds=feeds,
                           js_required=js_required, js_rendered=js_rendered)

    # ------------------------------------------------------------------ helpers
    async def _try_js_render(self, url: str) -> str | None:
        if self._js is None or not self._config.js_rendering:
            return None
        try:
            res = await self._js.render(url)
            return res.html if res.success and res.html else None
        except Exception as exc:  # noqa: BLE001
            logger.debug("js_render_error", url=url, error=str(exc))
            return None

    async def _fetch_with_retry(self, client, url: str, t0: float):
        """GET with exponential backoff on 5xx / transport errors; the final URL is re-validated (DNS included)."""
        last = ""
        for attempt in range(_MAX_RETRIES + 1):
            resp = None
            try:
                resp = await client.get(url, timeout=30.0)
                status = getattr(resp, "status_code", 200)
                if status >= 400:
                    if 500 <= status < 600 and attempt < _MAX_RETRIES:
                        await asyncio.sleep(_RETRY_BACKOFF_BASE * (2 ** attempt))
                        continue
                    logger.warning("crawl_http_error", url=url, status=status)
                    self._scheduler.mark_error(url)
                    last = f"http_{status}"
                    break
                final = str(getattr(resp, "url", url))
                if final != url:
                    if self._resolve_dns:
                        validate_url_post_redirect(final)
                    else:
                        validate_url(final)
                return resp
            except SSRFError as exc:
                logger.warning("crawl_ssrf_redirect", url=url, reason=str(exc))
                return self._fail(url, t0, f"redirect_blocked: {exc}")
            except Exception as exc:  # noqa: BLE001 — httpx.HTTPError, OSError, timeouts …
                if attempt < _MAX_RETRIES:
                    await asyncio.sleep(_RETRY_BACKOFF_BASE * (2 ** attempt))
                    continue
                logger.warning("crawl_network_error", url=url, error=str(exc))
                self._scheduler.mark_error(url)
                last = str(exc) or type(exc).__name__
        return self._fail(url, t0, last)

    async def _schedule_sitemap_urls(self, domain: str) -> None:
        """Once per domain: pull <loc> URLs out of the sitemaps announced in robots.txt (depth 0)."""
        sitemaps = self._robots.get_sitemaps(domain)
        if not sitemaps:
            return
        marker = f"__sitemap_processed__{domain}"
        if self._dedup.is_url_seen(marker):
            return
        self._dedup.mark_seen(marker, "sitemap", "")
        client = await self._get_client()
        total = 0
        for sm in sitemaps:
            try:
                validate_url(sm)
                resp = await client.get(sm, timeout=15.0, follow_redirects=True)
                if getattr(resp, "status_code", 0) != 200:
                    continue
                for loc in _SITEMAP_LOC.findall(resp.text)[:_MAX_SITEMAP_URLS]:
                    if loc.endswith(".xml") or self._dedup.is_url_seen(loc):
                        continue
                    if await self._scheduler.add_url(loc, depth=0):
                        total += 1
            except Exception as exc:  # noqa: BLE001
                logger.debug("sitemap_fetch_failed", sitemap=sm, error=str(exc))
        if total:
            logger.info("sitemap_urls_scheduled", domain=domain, scheduled=total)
