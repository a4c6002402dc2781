"""Labelled usage bar that turns yellow at 70 % and red at 90 % (reference infomesh/dashboard/widgets/resource_bar.py:12-77)."""
from __future__ import annotations

from textual.widgets import Static


def render_resource(label: str, ratio: float, detail: str = "", *, width: int = 24) -> str:
    ratio = max(0.0, min(1.0, ratio))
    n = int(ratio * width)
    color = "red" if ratio >= 0.9 else "yellow" if ratio >= 0.7 else "green"
    return f"[bold]{label:<6}[/] [{color}]{'█' * n}[/][dim]{'░' * (width - n)}[/] [{color}]{ratio * 100:3.0f}%[/] [dim]{detail}[/]"


class ResourceBar(Static):
    """``ResourceBar("CPU")`` then ``set_value(ratio, detail)``; the reference's constructor form
    ``ResourceBar(label="CPU", value=38, max_value=100, unit="%", color="cyan", bar_width=12)`` is accepted as well."""

    def __init__(self, label: str = "", value: float = 0.0, max_value: float = 100.0, *, unit: str = "%", color: str = "green", bar_width: int = 24, **kw):
        self._label, self._max_value, self._unit, self._color, self._bar_width = label, float(max_value), unit, color, int(bar_width)
        super().__init__(self._bar_markup(float(value) / self._max_value if self._max_value > 0 else 0.0), **kw)

    def _bar_markup(self, ratio: float, detail: str = "") -> str:
        if not detail and self._unit != "%":
            detail = f"{ratio * self._max_value:.1f}/{self._max_value:.1f} {self._unit}"
        return render_resource(self._label, ratio, detail, width=self._bar_width)

    def set_value(self, ratio: float, detail: str = "") -> None:
        self.update(self._bar_markup(ratio, detail))

    def update_value(self, value: float, max_value: float | None = None) -> None:
        """Absolute value against a maximum (default 100), as the reference widget is driven."""
        if max_value is not None:
            self._max_value = float(max_value)
        self.set_value(value / self._max_value if self._max_value > 0 else 0.0)
