"""Labelled usage bar that turns yellow at 70 % and red at 90 % (reference infomesh/dashboard/widgets/resource_bar.py:12-77)."""
from __future__ import annotations

from textual.widgets import Static


def render_resource(label: str, ratio: float, detail: str = "", *, width: int = 24) -> str:
    ratio = max(0.0, min(1.0, ratio))
    n = int(ratio * width)
    color = "red" if ratio >= 0.9 else "yellow" if ratio >= 0.7 else "green"
    return f"[bold]{label:<6}[/] [{color}]{'█' * n}[/][dim]{'░' * (width - n)}[/] [{color}]{ratio * 100:3.0f}%[/] [dim]{detail}[/]"


class ResourceBar(Static):
    def __init__(self, label: str, **kw):
        super().__init__(render_resource(label, 0.0), **kw)
        self._label = label
        self._max_value = 100.0

    def set_value(self, ratio: float, detail: str = "") -> None:
        self.update(render_resource(self._label, ratio, detail))

    def update_value(self, value: float, max_value: float | None = None) -> None:
        """Absolute value against a maximum (default 100), as the reference widget is driven."""
        if max_value is not None:
            self._max_value = float(max_value)
        self.set_value(value / self._max_value if self._max_value > 0 else 0.0)
