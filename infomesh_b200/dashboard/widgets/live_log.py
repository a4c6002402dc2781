"""Bounded scrolling log (reference infomesh/dashboard/widgets/live_log.py:14-99)."""
from __future__ import annotations

import time
from collections import deque

from textual.widgets import Static


class LiveLog(Static):
    def __init__(self, *, max_lines: int = 200, visible: int = 12, **kw):
        super().__init__("[dim]waiting for activity…[/]", **kw)
        self._lines: deque[str] = deque(maxlen=max_lines)
        self._visible = visible

    def write_line(self, line: str) -> None:
        self._lines.append(line.replace("[", "\\["))
        self.update("\n".join(list(self._lines)[-self._visible:]))

    @property
    def lines(self) -> list[str]:
        return list(self._lines)

    def clear_log(self) -> None:
        self._lines.clear()
        self.update("")

    # ---- typed events: one timestamped line each (reference widgets/live_log.py:52-99) ----
    def log_event(self, message: str, *, style: str = "") -> None:
        body = message.replace("[", "\\[")
        self._append_markup(f"[dim]{time.strftime('%H:%M:%S')}[/] " + (f"[{style}]{body}[/]" if style else body))

    def log_crawl(self, url: str, *, success: bool = True, credits: float = 0) -> None:
        mark = "[green]✓[/]" if success else "[red]✗[/]"
        shown = url.replace("[", "\\[")
        earned = f"  [cyan]+{credits:.1f} cr[/]" if credits > 0 else ""
        self._append_markup(f"[dim]{time.strftime('%H:%M:%S')}[/] {mark} " + (f"[bold]{shown}[/]" if success else f"[dim strike]{shown}[/]") + earned)

    def log_search(self, query: str, count: int, elapsed_ms: float) -> None:
        q = query.replace("[", "\\[")
        self._append_markup(f'[dim]{time.strftime("%H:%M:%S")}[/] [yellow]🔍[/] [bold]"{q}"[/] [dim]({count} results, {elapsed_ms:.0f}ms)[/]')

    def log_peer(self, peer_id: str, *, connected: bool = True) -> None:
        color, what = ("green", "connected") if connected else ("red", "disconnected")
        self._append_markup(f"[dim]{time.strftime('%H:%M:%S')}[/] [{color}]Peer {peer_id[:12]}... {what}[/]")

    def _append_markup(self, line: str) -> None:
        self._lines.append(line)
        self.update("\n".join(list(self._lines)[-self._visible:]))
