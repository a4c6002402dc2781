"""Horizontal bar chart of (label, count) pairs (reference infomesh/dashboard/widgets/bar_chart.py:14-90)."""
from __future__ import annotations

from textual.widgets import Static


def render_bars(items: list[tuple[str, float]], *, width: int = 28, label_width: int = 26) -> str:
    if not items:
        return "[dim]no data yet[/]"
    peak = max(v for _, v in items) or 1.0
    rows = []
    for label, v in items:
        n = max(1, int(v / peak * width)) if v > 0 else 0
        name = label if len(label) <= label_width else label[:label_width - 1] + "…"
        rows.append(f"{name:<{label_width}} [green]{'█' * n}[/][dim]{'░' * (width - n)}[/] {v:,.0f}")
    return "\n".join(rows)


class BarChart(Static):
    def set_items(self, items: list[tuple[str, float]]) -> None:
        self.update(render_bars(items))
