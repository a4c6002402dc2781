"""The Textual dashboard application: six tabs (overview / crawl / search / network / credits / settings), key bindings
1-6, r refresh, m music, q quit with a "stop the node too?" prompt (reference infomesh/dashboard/app.py:40-518)."""
from __future__ import annotations

from textual.app import App, ComposeResult
from textual.binding import Binding
from textual.containers import Horizontal, Vertical
from textual.screen import ModalScreen
from textual.widgets import Button, Footer, Header, Static, TabbedContent, TabPane

from infomesh_b200 import __version__
from infomesh_b200.config import Config, load_config
from infomesh_b200.dashboard.bgm import BGMPlayer
from infomesh_b200.dashboard.data_cache import DashboardDataCache
from infomesh_b200.dashboard.screens.crawl import CrawlPane
from infomesh_b200.dashboard.screens.credits import CreditsPane
from infomesh_b200.dashboard.screens.network import NetworkPane
from infomesh_b200.dashboard.screens.overview import OverviewPane
from infomesh_b200.dashboard.screens.search import SearchPane
from infomesh_b200.dashboard.screens.settings import SettingsPane

TABS = ("overview", "crawl", "search", "network", "credits", "settings")


class QuitConfirmScreen(ModalScreen[str]):
    """-> "stop_all" | "dashboard_only" | "cancel"."""
    BINDINGS = [Binding("escape", "cancel", "Cancel")]
    DEFAULT_CSS = "QuitConfirmScreen {align: center middle;} #quit-box {width: 58; height: auto; border: round $accent; padding: 1 2; background: $surface;}"

    def compose(self) -> ComposeResult:
        with Vertical(id="quit-box"):
            yield Static("[bold]Quit the dashboard[/]\nThe node keeps crawling and serving peers unless you stop it too.")
            with Horizontal():
                yield Button("Dashboard only", id="dashboard_only", variant="primary")
                yield Button("Stop node too", id="stop_all", variant="error")
                yield Button("Cancel", id="cancel")

    def on_button_pressed(self, event: Button.Pressed) -> None:
        self.dismiss(event.button.id or "cancel")

    def action_cancel(self) -> None:
        self.dismiss("cancel")


class DashboardApp(App[None]):
    TITLE = "InfoMesh Dashboard"
    SUB_TITLE = f"v{__version__}"
    BINDINGS = [*(Binding(str(i + 1), f"tab('{name}')", name.title()) for i, name in enumerate(TABS)), Binding("r", "refresh", "Refresh"),
                Binding("m", "toggle_bgm", "Music"), Binding("question_mark", "help", "Help"), Binding("q", "quit", "Quit")]

    def __init__(self, config: Config | None = None, *, initial_tab: str = "overview", node_pid: int | None = None):
        super().__init__()
        self.config = config or load_config()
        self.initial_tab = initial_tab if initial_tab in TABS else "overview"
        self.node_pid = node_pid
        self.exit_action = "dashboard_only"
        self.cache = DashboardDataCache(self.config, ttl=max(self.config.dashboard.refresh_interval, 0.2))
        self.bgm = BGMPlayer()

    def compose(self) -> ComposeResult:
        yield Header()
        with TabbedContent(initial=self.initial_tab, id="tabs"):
            with TabPane("Overview", id="overview"):
                yield OverviewPane(self.config, self.cache)
            with TabPane("Crawl", id="crawl"):
                yield CrawlPane(self.config, self.cache)
            with TabPane("Search", id="search"):
                yield SearchPane(self.config)
            with TabPane("Network", id="network"):
                yield NetworkPane(self.config)
            with TabPane("Credits", id="credits"):
                yield CreditsPane(self.config)
            with TabPane("Settings", id="settings"):
                yield SettingsPane(self.config)
        yield Footer()

    def on_mount(self) -> None:
        theme = getattr(self.config.dashboard, "theme", "")
        if theme and theme in getattr(self, "available_themes", {}):
            self.theme = theme
        self.set_interval(15.0, self._check_bgm_health)

    def _check_bgm_health(self) -> None:
        self.bgm.reap_sfx()
        self.bgm.check_and_restart()

    def set_data_cache_ttl(self, ttl: float) -> None:
        self.cache.set_ttl(ttl)

    def update_config(self, config: Config) -> None:
        self.config = config

    def on_settings_pane_config_changed(self, event: SettingsPane.ConfigChanged) -> None:
        self.update_config(event.config)
        self.set_data_cache_ttl(max(event.config.dashboard.refresh_interval, 0.2))

    def on_credits_pane_credit_earned(self, event: CreditsPane.CreditEarned) -> None:
        self.notify(f"+{event.amount:.2f} credits", title="Credits earned", timeout=3)

    def action_tab(self, name: str) -> None:
        self.query_one("#tabs", TabbedContent).active = name

    def action_refresh(self) -> None:
        self.cache.set_ttl(0.0)
        for pane in self.query(".refreshable"):
            pane.refresh()
        for cls in (OverviewPane, CrawlPane, NetworkPane, CreditsPane):
            for pane in self.query(cls):
                pane.refresh_data()
        self.cache.set_ttl(max(self.config.dashboard.refresh_interval, 0.2))

    def action_toggle_bgm(self) -> None:
        from infomesh_b200.dashboard.bgm import ensure_bgm_assets

        tracks = sorted(p for p in ensure_bgm_assets().glob("*") if p.suffix.lower() in (".mp3", ".ogg", ".wav", ".flac"))
        if not self.bgm.available or not tracks:
            self.notify("No audio player (mpv/ffplay) or no tracks in ~/.infomesh/bgm", title="Music", severity="warning")
            return
        on = self.bgm.toggle(tracks[0], volume=getattr(self.config.dashboard, "bgm_volume", 50))
        self.notify("playing" if on else "stopped", title="Music")

    def action_help(self) -> None:
        self.notify("1-6 switch tabs · r refresh · m music · q quit · in Settings type section.key = value", title="Keys", timeout=8)

    def action_quit(self) -> None:  # type: ignore[override]
        if self.node_pid is None:
            self._finish("dashboard_only")
        else:
            self.push_screen(QuitConfirmScreen(), self._finish)

    def _finish(self, result: str | None) -> None:
        if result in (None, "cancel"):
            return
        self.exit_action = result
        self.bgm.stop()
        self.cache.close()
        self.exit()


def run_dashboard(config: Config | None = None, *, initial_tab: str = "overview", node_pid: int | None = None) -> str:
    """Blocks until the TUI exits; returns "stop_all" or "dashboard_only"."""
    app = DashboardApp(config, initial_tab=initial_tab, node_pid=node_pid)
    try:
        app.run()
    finally:
        app.bgm.stop()
        app.cache.close()
    return app.exit_action
