"""Optional headless-Chromium rendering through Playwright (lazy import; absent here -> ``available`` is False).
Limits follow reference infomesh/crawler/js_render.py:67-215: bounded concurrent tabs, per-page timeout, memory cap."""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass
class RenderResult:
    success: bool = False
    html: str = ""
    error: str | None = None
    elapsed_ms: float = 0.0
    url: str = ""                # this build also records what was asked for and where the browser ended up
    final_url: str = ""


def is_playwright_available() -> bool:
    try:
        import playwright.async_api  # type: ignore # noqa: F401

        return True
    except Exception:  # noqa: BLE001
        return False


class JSRenderer:
    def __init__(self, *, max_tabs: int = 3, timeout_ms: int = 30_000, max_memory_mb: int = 512,
                 user_agent: str = "InfoMesh/0.1"):
        self._sem = asyncio.Semaphore(max(1, max_tabs))
        self._timeout_ms = timeout_ms
        self._max_memory_mb = max_memory_mb
        self._ua = user_agent
        self._pw = None
        self._browser = None
        self._lock = asyncio.Lock()

    @property
    def available(self) -> bool:
        return is_playwright_available()

    async def _ensure_browser(self):
        async with self._lock:
            if self._browser is None:
                from playwright.async_api import async_playwright  # type: ignore

                self._pw = await async_playwright().start()
                self._browser = await self._pw.chromium.launch(
                    headless=True, args=["--no-sandbox", "--disable-gpu", "--disable-dev-shm-usage",
                                         f"--js-flags=--max-old-space-size={self._max_memory_mb}"])
            return self._browser

    async def render(self, url: str) -> RenderResult:
        t0 = time.monotonic()
        if not self.available:
            return RenderResult(url=url, error="playwright_not_installed")
        from infomesh_b200.security import SSRFError, validate_url

        try:
            validate_url(url, resolve_dns=True)
        except SSRFError as exc:
            return RenderResult(url=url, error=f"blocked: {exc}")
        async with self._sem:
            page = None
            try:
                browser = await self._ensure_browser()
                page = await browser.new_page(user_agent=self._ua)
                await page.goto(url, timeout=self._timeout_ms, wait_until="networkidle")
                html = await page.content()
                return RenderResult(True, html, None, (time.monotonic() - t0) * 1000, url=url, final_url=page.url)
            except Exception as exc:  # noqa: BLE001
                logger.warning("js_render_failed", url=url, error=str(exc))
                return RenderResult(url=url, error=str(exc), elapsed_ms=(time.monotonic() - t0) * 1000)
            finally:
                if page is not None:
                    try:
                        await page.close()
                    except Exception:  # noqa: BLE001
                        pass

    async def close(self) -> None:
        try:
            if self._browser is not None:
                await self._browser.close()
            if self._pw is not None:
                await self._pw.stop()
        except Exception:  # noqa: BLE001
            pass
        self._browser = self._pw = None
