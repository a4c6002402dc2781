"""Settings tab: browse every config section and edit values in place (validated + persisted through
``set_config_value`` / ``save_config``) (reference infomesh/dashboard/screens/settings.py:40-665)."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.message import Message
from textual.widgets import Input, Static

from infomesh_b200.config import config_to_dict, save_config, set_config_value


class SettingsPane(Vertical):
    class ConfigChanged(Message):
        def __init__(self, config):
            super().__init__()
            self.config = config

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config

    def compose(self) -> ComposeResult:
        yield Static("[bold]Edit a setting[/] — type [cyan]section.key = value[/] and press Enter (e.g. crawl.politeness_delay = 2.0)")
        yield Input(placeholder="section.key = value", id="st-input")
        yield Static("", id="st-msg")
        yield Static("", id="st-table")

    def on_mount(self) -> None:
        self.render_table()

    def render_table(self) -> None:
        rows = []
        for section, values in config_to_dict(self.config, redact=True).items():
            rows.append(f"[bold cyan]\\[{section}][/]")
            rows += [f"  {k:<24} {v}" for k, v in values.items()]
        self.query_one("#st-table", Static).update("\n".join(rows))

    def apply_edit(self, text: str) -> str:
        key, sep, value = text.partition("=")
        if not sep or "." not in key:
            return "[red]expected: section.key = value[/]"
        try:
            new = set_config_value(self.config, key.strip(), value.strip())
            save_config(new)
        except (KeyError, ValueError, TypeError, OSError) as exc:
            return f"[red]{exc}[/]"
        self.config = new
        self.post_message(self.ConfigChanged(new))
        self.render_table()
        return f"[green]✔ saved {key.strip()} (restart the node for crawler/network settings to apply)[/]"

    def on_input_submitted(self, event: Input.Submitted) -> None:
        self.query_one("#st-msg", Static).update(self.apply_edit(event.value))
        event.input.value = ""
