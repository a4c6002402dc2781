"""Horizontal bar chart of (label, count) pairs (reference infomesh/dashboard/widgets/bar_chart.py:14-90)."""
from __future__ import annotations

from dataclasses import dataclass

from textual.widgets import Static


@dataclass
class BarItem:
    label: str
    value: float
    color: str = "green"
    suffix: str = ""


def render_bars(items: "list[tuple[str, float] | BarItem]", *, width: int = 28, label_width: int = 26) -> str:
    """``items``: (label, value) pairs or :class:`BarItem` objects (per-bar colour and suffix)."""
    if not items:
        return "[dim]no data yet[/]"
    bars = [i if isinstance(i, BarItem) else BarItem(str(i[0]), float(i[1])) for i in items]
    peak = max(b.value for b in bars) or 1.0
    rows = []
    for b in bars:
        n = max(1, int(b.value / peak * width)) if b.value > 0 else 0
        name = b.label if len(b.label) <= label_width else b.label[:label_width - 1] + "…"
        rows.append(f"{name:<{label_width}} [{b.color}]{'█' * n}[/][dim]{'░' * (width - n)}[/] {b.value:,.0f}{b.suffix}")
    return "\n".join(rows)


class BarChart(Static):
    """``BarChart(items=[BarItem("Crawling", 702, color="cyan"), ...], bar_width=20)`` as in the reference, or empty and fed by ``set_items``."""

    def __init__(self, items: "list[tuple[str, float] | BarItem] | None" = None, *, bar_width: int = 28, **kw):
        self._bar_width = int(bar_width)
        super().__init__(render_bars(items or [], width=self._bar_width), **kw)

    def set_items(self, items: "list[tuple[str, float] | BarItem]") -> None:
        self.update(render_bars(items, width=self._bar_width))
