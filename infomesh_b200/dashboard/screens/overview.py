"""Overview tab: node identity, resource bars (disk / CPU / RAM / HBM), activity sparklines, recent documents
(reference infomesh/dashboard/screens/overview.py:30-408).  Three panels, each refreshable on its own."""
from __future__ import annotations

import shutil

from textual.app import ComposeResult
from textual.containers import Horizontal, Vertical
from textual.widgets import Static

from infomesh_b200 import __version__
from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import LiveLog, ResourceBar, SparklineChart



def _own_cache(config):
    """A pane constructed the reference's way (``Pane(config)``) reads through its own data cache."""
    from infomesh_b200.dashboard.data_cache import DashboardDataCache

    return DashboardDataCache(config, ttl=max(getattr(config.dashboard, "refresh_interval", 0.5), 0.2))

class NodeInfoPanel(Static):
    """Version, run state, uptime, identity, role and the index head-line numbers."""

    def __init__(self, config, cache=None, **kw):
        super().__init__("", **kw)
        self.config, self.cache = config, cache if cache is not None else _own_cache(config)

    def on_mount(self) -> None:
        self.refresh_data()

    def refresh_data(self) -> None:
        cfg = self.config
        running, up = U.is_node_running_with_uptime(cfg)
        st = self.cache.get_stats()
        state = "[green]● running[/]" if running else "[red]● stopped[/]"
        self.update(f"[bold]InfoMesh v{__version__}[/]  {state}  up {U.format_uptime(up)}  ·  peer {U.get_peer_id(cfg)[:16]}…  ·  role {cfg.node.role}\n"
                    f"documents [bold]{st.document_count:,}[/]  ·  domains {st.domain_count:,}  ·  last hour {st.pages_last_hour:,}")


class ResourcePanel(Vertical):
    """Disk / CPU / RAM / HBM usage bars; returns the CPU sample so the activity panel can plot it."""

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config

    def compose(self) -> ComposeResult:
        yield ResourceBar("Disk", id="ov-disk")
        yield ResourceBar("CPU", id="ov-cpu")
        yield ResourceBar("RAM", id="ov-ram")
        yield ResourceBar("HBM", id="ov-hbm")

    def refresh_data(self) -> float | None:
        cpu = None
        try:
            d = self.config.node.data_dir
            du = shutil.disk_usage(str(d if d.exists() else "/"))
            self.query_one("#ov-disk", ResourceBar).set_value(du.used / du.total, f"{U.format_bytes(du.used)} / {U.format_bytes(du.total)}")
        except OSError:
            pass
        try:
            import psutil

            cpu = psutil.cpu_percent(interval=None)
            mem = psutil.virtual_memory()
            self.query_one("#ov-cpu", ResourceBar).set_value(cpu / 100)
            self.query_one("#ov-ram", ResourceBar).set_value(mem.percent / 100, f"{U.format_bytes(mem.used)} / {U.format_bytes(mem.total)}")
        except ImportError:
            pass
        try:
            import torch

            if torch.cuda.is_available():
                free, total = torch.cuda.mem_get_info()
                self.query_one("#ov-hbm", ResourceBar).set_value(1 - free / total, f"{U.format_bytes(total - free)} / {U.format_bytes(total)}")
        except Exception:  # noqa: BLE001
            pass
        return cpu


class ActivityPanel(Vertical):
    """Indexing rate and CPU sparklines over a log of recently indexed documents."""

    def __init__(self, config, cache=None, **kw):
        super().__init__(**kw)
        self.config, self.cache = config, cache if cache is not None else _own_cache(config)
        self._seen: set[int] = set()
        self._last_count = -1

    def compose(self) -> ComposeResult:
        yield SparklineChart("docs/min", color="green", id="ov-rate")
        yield SparklineChart("cpu %", color="cyan", id="ov-cpuspark")
        yield Static("[bold]Recently indexed[/]")
        yield LiveLog(id="ov-log")

    # push-style updates for callers that count events themselves (reference screens/overview.py:253-276)
    def update_crawl(self, count: int) -> None:
        self.query_one("#ov-rate", SparklineChart).push(float(count))

    def update_index(self, count: int) -> None:
        self._last_count = int(count)

    def update_search(self, count: int) -> None:
        self.query_one("#ov-log", LiveLog).log_event(f"{count} queries served", style="dim")

    def refresh_data(self, cpu: float | None = None) -> None:
        st = self.cache.get_stats()
        period = max(self.config.dashboard.refresh_interval, 0.2)
        if cpu is not None:
            self.query_one("#ov-cpuspark", SparklineChart).push(cpu)
        if self._last_count >= 0:
            self.query_one("#ov-rate", SparklineChart).push(max(0, st.document_count - self._last_count) * 60 / period)
        self._last_count = st.document_count
        self._seen, self._last_count = U.push_new_docs_to_log(st.recent_docs, st.document_count, self._seen, getattr(self, "_last_count", -1),
                                                                  self.query_one("#ov-log", LiveLog))


class OverviewPane(Vertical):
    def __init__(self, config, cache=None, **kw):
        super().__init__(**kw)
        self.config, self.cache = config, cache if cache is not None else _own_cache(config)

    def compose(self) -> ComposeResult:
        yield NodeInfoPanel(self.config, self.cache, id="ov-node")
        with Horizontal():
            yield ResourcePanel(self.config, id="ov-resources")
            yield ActivityPanel(self.config, self.cache, id="ov-activity")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(max(self.config.dashboard.refresh_interval, 0.2), self.refresh_data)

    def refresh_data(self) -> None:
        self.query_one(NodeInfoPanel).refresh_data()
        cpu = self.query_one(ResourcePanel).refresh_data()
        self.query_one(ActivityPanel).refresh_data(cpu)
