"""Overview tab: node identity, resource bars (disk / CPU / RAM / HBM), activity sparklines, recent documents
(reference infomesh/dashboard/screens/overview.py:30-408)."""
from __future__ import annotations

import shutil

from textual.app import ComposeResult
from textual.containers import Horizontal, Vertical
from textual.widgets import Static

from infomesh_b200 import __version__
from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import LiveLog, ResourceBar, SparklineChart


class OverviewPane(Vertical):
    def __init__(self, config, cache, **kw):
        super().__init__(**kw)
        self.config, self.cache = config, cache
        self._seen: set[int] = set()
        self._last_count = -1

    def compose(self) -> ComposeResult:
        yield Static("", id="ov-node")
        with Horizontal():
            with Vertical():
                yield ResourceBar("Disk", id="ov-disk")
                yield ResourceBar("CPU", id="ov-cpu")
                yield ResourceBar("RAM", id="ov-ram")
                yield ResourceBar("HBM", id="ov-hbm")
            with Vertical():
                yield SparklineChart("docs/min", color="green", id="ov-rate")
                yield SparklineChart("cpu %", color="cyan", id="ov-cpuspark")
        yield Static("[bold]Recently indexed[/]")
        yield LiveLog(id="ov-log")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(max(self.config.dashboard.refresh_interval, 0.2), self.refresh_data)

    def refresh_data(self) -> None:
        cfg = self.config
        running, up = U.is_node_running_with_uptime(cfg)
        st = self.cache.get_stats()
        state = "[green]● running[/]" if running else "[red]● stopped[/]"
        self.query_one("#ov-node", Static).update(
            f"[bold]InfoMesh v{__version__}[/]  {state}  up {U.format_uptime(up)}  ·  peer {U.get_peer_id(cfg)[:16]}…  ·  role {cfg.node.role}\n"
            f"documents [bold]{st.document_count:,}[/]  ·  domains {st.domain_count:,}  ·  last hour {st.pages_last_hour:,}")
        try:
            d = cfg.node.data_dir
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
            self.query_one("#ov-cpuspark", SparklineChart).push(cpu)
        except ImportError:
            pass
        try:
            import torch

            if torch.cuda.is_available():
                free, total = torch.cuda.mem_get_info()
                self.query_one("#ov-hbm", ResourceBar).set_value(1 - free / total, f"{U.format_bytes(total - free)} / {U.format_bytes(total)}")
        except Exception:  # noqa: BLE001
            pass
        if self._last_count >= 0:
            self.query_one("#ov-rate", SparklineChart).push(max(0, st.document_count - self._last_count) * 60 / max(cfg.dashboard.refresh_interval, 0.2))
        self._last_count = st.document_count
        U.push_new_docs_to_log(self.query_one("#ov-log", LiveLog), st.recent_docs, self._seen)
