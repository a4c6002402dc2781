"""Settings tab: browse every config section, edit values in place with ``section.key = value`` (validated and persisted
through ``set_config_value`` / ``save_config``), stage several edits and save them together, reset to defaults, and ask
before restarting the node when a changed key only takes effect at start-up
(reference infomesh/dashboard/screens/settings.py:40-665)."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Horizontal, Vertical
from textual.message import Message
from textual.screen import ModalScreen
from textual.widgets import Button, Input, Static

from infomesh_b200.config import Config, config_to_dict, save_config, set_config_value

# read once when the node process starts; everything else is picked up by the running dashboard immediately
RESTART_REQUIRED_KEYS = frozenset({
    "network.upload_limit_mbps", "network.download_limit_mbps", "network.replication_factor", "node.listen_port", "node.role",
    "storage.compression_enabled", "storage.compression_level", "storage.max_cache_size_mb", "storage.max_index_size_gb",
    "storage.cache_ttl_days", "gpu.enabled", "gpu.devices", "gpu.tp", "gpu.backend", "gpu.shard_dtype"})


def restart_keys_changed(old: Config, new: Config) -> list[str]:
    a, b = config_to_dict(old), config_to_dict(new)
    return sorted(k for k in RESTART_REQUIRED_KEYS
                  if a.get(k.split(".")[0], {}).get(k.split(".")[1]) != b.get(k.split(".")[0], {}).get(k.split(".")[1]))


class RestartConfirmScreen(ModalScreen[bool]):
    """-> True when the user wants the node restarted now."""
    DEFAULT_CSS = "RestartConfirmScreen {align: center middle;} #restart-box {width: 60; height: auto; border: round $warning; padding: 1 2; background: $surface;}"

    def __init__(self, changed_keys: list[str]):
        super().__init__()
        self.keys = list(changed_keys)

    def compose(self) -> ComposeResult:
        with Vertical(id="restart-box"):
            yield Static("[bold]Restart needed[/]\nThese settings are read when the node starts:\n  " + "\n  ".join(self.keys))
            with Horizontal():
                yield Button("Restart now", id="btn-restart-yes", variant="warning")
                yield Button("Later", id="btn-restart-no", variant="primary")

    def on_button_pressed(self, event: Button.Pressed) -> None:
        self.dismiss(event.button.id == "btn-restart-yes")


class SettingsPane(Vertical):
    class ConfigChanged(Message):
        def __init__(self, config):
            super().__init__()
            self.config = config

    class RestartRequested(Message):
        def __init__(self, keys: list[str]):
            super().__init__()
            self.keys = keys

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config
        self._saved = config               # what is on disk; `config` may carry a staged reset

    def compose(self) -> ComposeResult:
        yield Static("[bold]Edit a setting[/] — type [cyan]section.key = value[/] and press Enter (e.g. crawl.politeness_delay = 2.0)")
        yield Input(placeholder="section.key = value", id="st-input")
        with Horizontal(id="st-buttons"):
            yield Button("Save", id="btn-save", variant="success")
            yield Button("Reset to defaults", id="btn-reset")
        yield Static("", id="st-msg")
        yield Static("", id="st-table")

    def on_mount(self) -> None:
        self.render_table()

    # ------------------------------------------------------------------ public hooks used by the app
    def update_config(self, config: Config) -> None:
        """The app changed the config behind our back (theme picked in the palette): show and keep the new one."""
        self.config = self._saved = config
        if self.is_mounted:
            self.render_table()

    def refresh_data(self) -> None:
        """Nothing to poll: this tab only changes when the user edits something."""

    # ------------------------------------------------------------------ rendering / editing
    def render_table(self) -> None:
        rows = []
        for section, values in config_to_dict(self.config, redact=True).items():
            rows.append(f"[bold cyan]\\[{section}][/]")
            rows += [f"  {k:<24} {v}" + ("  [dim](restart)[/]" if f"{section}.{k}" in RESTART_REQUIRED_KEYS else "") for k, v in values.items()]
        self.query_one("#st-table", Static).update("\n".join(rows))

    def _commit(self, new: Config) -> list[str]:
        save_config(new)
        keys = restart_keys_changed(self._saved, new)
        self.config = self._saved = new
        self.post_message(self.ConfigChanged(new))
        self.render_table()
        return keys

    def apply_edit(self, text: str) -> str:
        key, sep, value = text.partition("=")
        if not sep or "." not in key:
            return "[red]expected: section.key = value[/]"
        try:
            keys = self._commit(set_config_value(self.config, key.strip(), value.strip()))
        except (KeyError, ValueError, TypeError, OSError) as exc:
            return f"[red]{exc}[/]"
        self._offer_restart(keys)
        return f"[green]✔ saved {key.strip()}[/]" + (" [yellow](takes effect after a node restart)[/]" if keys else "")

    def save(self) -> str:
        try:
            keys = self._commit(self.config)
        except OSError as exc:
            return f"[red]{exc}[/]"
        self._offer_restart(keys)
        return "[green]✔ saved[/]" + (" [yellow]— some changes need a restart[/]" if keys else "")

    def reset_to_defaults(self) -> str:
        """Stage the defaults (identity-related and path settings are kept); nothing is written until Save."""
        from dataclasses import replace

        d = Config()
        self.config = replace(d, node=replace(d.node, data_dir=self.config.node.data_dir, github_email=self.config.node.github_email),
                              index=replace(d.index, db_path=self.config.index.db_path))
        self.render_table()
        return "[yellow]↺ defaults staged — press Save to write them[/]"

    def _offer_restart(self, keys: list[str]) -> None:
        if not keys:
            return

        def answered(restart: bool | None) -> None:
            if restart:
                self.post_message(self.RestartRequested(keys))

        try:
            self.app.push_screen(RestartConfirmScreen(keys), answered)
        except Exception:  # noqa: BLE001 — no running app (unit tests drive the pane directly)
            pass

    # ------------------------------------------------------------------ events
    def on_button_pressed(self, event: Button.Pressed) -> None:
        msg = self.save() if event.button.id == "btn-save" else self.reset_to_defaults() if event.button.id == "btn-reset" else ""
        if msg:
            self.query_one("#st-msg", Static).update(msg)

    def on_input_submitted(self, event: Input.Submitted) -> None:
        self.query_one("#st-msg", Static).update(self.apply_edit(event.value))
        event.input.value = ""
