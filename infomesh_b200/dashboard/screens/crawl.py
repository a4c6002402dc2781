"""Crawl tab: throughput, top domains, live crawl log (reference infomesh/dashboard/screens/crawl.py:24-295)."""
from __future__ import annotations

import time

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import BarChart, LiveLog



def _own_cache(config):
    """A pane constructed the reference's way (``Pane(config)``) reads through its own data cache."""
    from infomesh_b200.dashboard.data_cache import DashboardDataCache

    return DashboardDataCache(config, ttl=max(getattr(config.dashboard, "refresh_interval", 0.5), 0.2))

class CrawlStatsPanel(Static):
    """One line of crawl throughput and the limits currently in force."""

    def __init__(self, config, **kw):
        super().__init__("", **kw)
        self.config = config

    def on_mount(self) -> None:
        self.update("[dim]waiting for crawl statistics…[/]")

    def show(self, stats) -> None:
        self.update_stats(pages_per_hour=stats.pages_last_hour, last_crawl_at=stats.last_crawl_at)

    def update_stats(self, total_pages: int = 0, pages_per_hour: int = 0, domain_count: int = 0, last_crawl_at: float = 0.0,
                     countdown: int = 0) -> None:
        """Push-style update for callers that already hold the numbers (reference screens/crawl.py:60)."""
        self._values = (total_pages, pages_per_hour, domain_count, last_crawl_at)
        self._countdown = countdown
        self._render_line()

    def update_countdown(self, countdown: int) -> None:
        self._countdown = countdown
        self._render_line()

    def _render_line(self) -> None:
        total, per_hour, domains, last = getattr(self, "_values", (0, 0, 0, 0.0))
        ago = U.format_uptime(time.time() - last) + " ago" if last else "never"
        c = self.config.crawl
        extra = (f"{total:,} pages · {domains:,} domains  ·  " if total or domains else "")
        tick = f"  ·  refresh in {self._countdown}s" if getattr(self, "_countdown", 0) > 0 else ""
        self.update(f"{extra}pages last hour [bold]{per_hour:,}[/]  ·  last crawl {ago}  ·  limit {c.urls_per_hour}/h  ·  "
                    f"{c.max_concurrent} connections  ·  delay {c.politeness_delay}s  ·  RSS {'on' if c.rss_enabled else 'off'}{tick}")


class TopDomainsPanel(BarChart):
    """Domains ranked by indexed pages."""

    def __init__(self, *a, cache=None, **kw):
        super().__init__(*a, **kw)
        self.cache = cache

    def show(self, stats) -> None:
        self.set_items([(d, float(n)) for d, n in stats.top_domains])

    def refresh_data(self) -> None:
        if self.cache is not None:
            self.show(self.cache.get_stats())


class CrawlPane(Vertical):
    def __init__(self, config, cache=None, **kw):
        super().__init__(**kw)
        self.config, self.cache = config, cache if cache is not None else _own_cache(config)
        self._seen: set[int] = set()

    def compose(self) -> ComposeResult:
        yield CrawlStatsPanel(self.config, id="cr-head")
        yield Static("[bold]Top domains[/]")
        yield TopDomainsPanel("", cache=self.cache, id="cr-domains")
        yield Static("[bold]Crawl log[/]")
        yield LiveLog(visible=14, id="cr-log")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(max(self.config.dashboard.refresh_interval, 0.2), self.refresh_data)

    def refresh_data(self) -> None:
        st = self.cache.get_stats()
        self.query_one(CrawlStatsPanel).show(st)
        self.query_one(TopDomainsPanel).show(st)
        self._seen, self._last_count = U.push_new_docs_to_log(st.recent_docs, st.document_count, self._seen, getattr(self, "_last_count", -1),
                                                                  self.query_one("#cr-log", LiveLog))
