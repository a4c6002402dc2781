"""Reusable Textual widgets."""
from infomesh_b200.dashboard.widgets.bar_chart import BarChart
from infomesh_b200.dashboard.widgets.live_log import LiveLog
from infomesh_b200.dashboard.widgets.resource_bar import ResourceBar
from infomesh_b200.dashboard.widgets.sparkline import SparklineChart, render_sparkline

__all__ = ["BarChart", "LiveLog", "ResourceBar", "SparklineChart", "render_sparkline"]
