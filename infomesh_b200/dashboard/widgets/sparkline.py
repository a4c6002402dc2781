"""Unicode sparkline over a rolling window (reference infomesh/dashboard/widgets/sparkline.py:13-75)."""
from __future__ import annotations

from collections import deque

from textual.widgets import Static

_TICKS = "▁▂▃▄▅▆▇█"


def render_sparkline(values, width: int = 30) -> str:
    vals = list(values)[-width:]
    if not vals:
        return " " * width
    lo, hi = min(vals), max(vals)
    span = (hi - lo) or 1.0
    return "".join(_TICKS[min(len(_TICKS) - 1, int((v - lo) / span * (len(_TICKS) - 1) + 0.5))] for v in vals).rjust(width)


class SparklineChart(Static):
    """``SparklineChart("docs/min")`` (label first, as the dashboard panes use it) or, like the reference widget,
    ``SparklineChart([1, 4, 2], color="green")`` with the initial series first."""

    def __init__(self, data: "list[float] | str | None" = None, *, label: str = "", width: int = 30, color: str = "cyan", **kw):
        super().__init__("", **kw)
        if isinstance(data, str):
            label, data = data, None
        self._label, self._width, self._color = label, width, color
        self._values: deque[float] = deque((float(v) for v in data or ()), maxlen=width)
        if self._values:
            self.update(f"[bold]{self._label:<10}[/] [{self._color}]{render_sparkline(self._values, self._width)}[/] {self._values[-1]:,.1f}")

    def push(self, value: float) -> None:
        self._values.append(float(value))
        last = self._values[-1]
        self.update(f"[bold]{self._label:<10}[/] [{self._color}]{render_sparkline(self._values, self._width)}[/] {last:,.1f}")

    def push_value(self, value: float, *, max_points: int = 30) -> None:
        if max_points != self._values.maxlen:
            self._values = deque(self._values, maxlen=max(1, max_points))
        self.push(value)

    @property
    def data(self) -> list[float]:
        return list(self._values)
