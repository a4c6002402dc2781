"""Crawl tab: throughput, top domains, live crawl log (reference infomesh/dashboard/screens/crawl.py:24-295)."""
from __future__ import annotations

import time

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import BarChart, LiveLog


class CrawlStatsPanel(Static):
    """One line of crawl throughput and the limits currently in force."""

    def __init__(self, config, **kw):
        super().__init__("", **kw)
        self.config = config

    def show(self, stats) -> None:
        ago = U.format_uptime(time.time() - stats.last_crawl_at) + " ago" if stats.last_crawl_at else "never"
        c = self.config.crawl
        self.update(f"pages last hour [bold]{stats.pages_last_hour:,}[/]  ·  last crawl {ago}  ·  limit {c.urls_per_hour}/h  ·  "
                    f"{c.max_concurrent} connections  ·  delay {c.politeness_delay}s  ·  RSS {'on' if c.rss_enabled else 'off'}")


class TopDomainsPanel(BarChart):
    """Domains ranked by indexed pages."""

    def show(self, stats) -> None:
        self.set_items([(d, float(n)) for d, n in stats.top_domains])


class CrawlPane(Vertical):
    def __init__(self, config, cache, **kw):
        super().__init__(**kw)
        self.config, self.cache = config, cache
        self._seen: set[int] = set()

    def compose(self) -> ComposeResult:
        yield CrawlStatsPanel(self.config, id="cr-head")
        yield Static("[bold]Top domains[/]")
        yield TopDomainsPanel("", id="cr-domains")
        yield Static("[bold]Crawl log[/]")
        yield LiveLog(visible=14, id="cr-log")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(max(self.config.dashboard.refresh_interval, 0.2), self.refresh_data)

    def refresh_data(self) -> None:
        st = self.cache.get_stats()
        self.query_one(CrawlStatsPanel).show(st)
        self.query_one(TopDomainsPanel).show(st)
        U.push_new_docs_to_log(self.query_one("#cr-log", LiveLog), st.recent_docs, self._seen)
