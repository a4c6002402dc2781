"""The Textual dashboard application: six tabs (overview / crawl / search / network / credits / settings), key bindings
1-6, r refresh, m music, q quit with a "stop the node too?" prompt (reference infomesh/dashboard/app.py:40-518)."""
from __future__ import annotations

import contextlib

from textual.app import App, ComposeResult
from textual.binding import Binding
from textual.command import DiscoveryHit, Hit, Hits, Provider
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


class DashboardCommandProvider(Provider):
    """Ctrl+P palette entries: the six tabs plus refresh / music / help / quit (reference dashboard/app.py:40-76)."""

    @staticmethod
    def _entries() -> list[tuple[str, str, str]]:
        tabs = [(name.title(), f"Switch to the {name.title()} tab ({i + 1})", f"app.tab_{i + 1}") for i, name in enumerate(TABS)]
        return tabs + [("Refresh", "Re-read every panel now (r)", "app.refresh"), ("Toggle BGM", "Background music on / off (m)", "app.toggle_bgm"),
                       ("Help", "Keyboard shortcuts (?)", "app.help"), ("Exit", "Quit the dashboard (q)", "app.quit")]

    def _runner(self, action: str):
        return lambda: self.app.run_action(action)

    async def discover(self) -> Hits:
        for name, text, action in self._entries():
            yield DiscoveryHit(display=name, command=self._runner(action), help=text)

    async def search(self, query: str) -> Hits:
        matcher = self.matcher(query)
        for name, text, action in self._entries():
            score = matcher.match(name)
            if score > 0:
                yield Hit(score=score, match_display=matcher.highlight(name), command=self._runner(action), help=text)


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

    def on_mount(self) -> None:
        with contextlib.suppress(Exception):           # the safe choice has the focus when the dialog opens
            self.query_one("#cancel", Button).focus()

    def on_button_pressed(self, event: Button.Pressed) -> None:
        self.dismiss(event.button.id or "cancel")

    def action_cancel(self) -> None:
        self.dismiss("cancel")


class DashboardApp(App[None]):
    TITLE = "InfoMesh Dashboard"
    SUB_TITLE = f"v{__version__}"
    COMMANDS = App.COMMANDS | {DashboardCommandProvider}
    BINDINGS = [*(Binding(str(i + 1), f"tab('{name}')", name.title()) for i, name in enumerate(TABS)), Binding("r", "refresh", "Refresh"),
                Binding("m", "toggle_bgm", "Music"), Binding("question_mark", "help", "Help"), Binding("q", "quit", "Quit")]

    def __init__(self, config: Config | None = None, initial_tab: str = "overview", node_pid: int | None = None):
        super().__init__()
        self.config = config or load_config()
        self.initial_tab = initial_tab if initial_tab in TABS else "overview"
        self.node_pid = node_pid
        self.exit_action = "dashboard_only"
        self.cache = DashboardDataCache(self.config, ttl=max(self.config.dashboard.refresh_interval, 0.2))
        self.bgm = BGMPlayer()
        self._theme_ready = False

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
        self._theme_ready = True
        self.set_interval(15.0, self._check_bgm_health)

    def watch_theme(self, new_theme: str) -> None:
        """A theme picked in the command palette is written back to config.toml."""
        if not getattr(self, "_theme_ready", False) or new_theme == getattr(self.config.dashboard, "theme", ""):
            return
        from dataclasses import replace

        from infomesh_b200.config import save_config

        try:
            new = replace(self.config, dashboard=replace(self.config.dashboard, theme=new_theme))
            save_config(new)
            self.update_config(new)
            with contextlib.suppress(Exception):
                self.query_one(SettingsPane).update_config(new)
        except Exception as exc:  # noqa: BLE001 — a read-only config file must not crash the TUI
            self.notify(f"theme not saved: {exc}", severity="warning")

    def on_unmount(self) -> None:
        self._cleanup()

    def _cleanup(self) -> None:
        """Idempotent: every exit path (quit, ctrl+c, crash) ends here."""
        self.bgm.stop()
        self.cache.close()

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

    def on_settings_pane_restart_requested(self, event: SettingsPane.RestartRequested) -> None:
        """Stop the node this dashboard was started with and launch a fresh worker with the saved config."""
        if self.node_pid is None:
            self.notify("Node is not running under this dashboard; restart it with `infomesh start`.", title="Restart", severity="warning")
            return
        from infomesh_b200 import runtime as RT
        from infomesh_b200.cli.serve import _serve_cmd, _spawn

        try:
            if not RT.request_graceful_stop(self.node_pid, timeout_seconds=10.0):
                self.notify(f"node {self.node_pid} did not exit; not restarted", title="Restart", severity="error")
                return
        except ProcessLookupError:
            pass
        RT.clear_pid_file(self.config.node.data_dir, self.node_pid)
        self.node_pid = _spawn(_serve_cmd(None, None)).pid
        RT.write_pid_file(self.config.node.data_dir, self.node_pid)
        self.notify(f"node restarted (PID {self.node_pid}) for: {', '.join(event.keys)}", title="Restart")

    def on_credits_pane_credit_earned(self, event: CreditsPane.CreditEarned) -> None:
        self.notify(f"+{event.amount:.2f} credits", title="Credits earned", timeout=3)

    def action_tab(self, name: str) -> None:
        self.query_one("#tabs", TabbedContent).active = name

    def action_tab_1(self) -> None:
        self.action_tab(TABS[0])

    def action_tab_2(self) -> None:
        self.action_tab(TABS[1])

    def action_tab_3(self) -> None:
        self.action_tab(TABS[2])

    def action_tab_4(self) -> None:
        self.action_tab(TABS[3])

    def action_tab_5(self) -> None:
        self.action_tab(TABS[4])

    def action_tab_6(self) -> None:
        self.action_tab(TABS[5])

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
        self._cleanup()
        self.exit()


def run_dashboard(config: Config | None = None, initial_tab: str = "overview", node_pid: int | None = None) -> str:
    """Blocks until the TUI exits; returns "stop_all" or "dashboard_only"."""
    app = DashboardApp(config, initial_tab=initial_tab, node_pid=node_pid)
    try:
        app.run()
    finally:
        app.bgm.stop()
        app.cache.close()
    return app.exit_action
