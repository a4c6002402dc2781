"""Bounded scrolling log (reference infomesh/dashboard/widgets/live_log.py:14-99)."""
from __future__ import annotations

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
