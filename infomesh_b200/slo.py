"""Service-level objectives: p99 search latency <= 1000 ms, search availability 99 %, crawl success 90 %, uptime
99.5 %, peer connectivity 95 %; sliding one-hour windows, error-budget accounting
(reference infomesh/slo.py:17-173).  Adds a GPU batch-latency objective for the device pipeline."""
from __future__ import annotations

import time
from collections import deque
from dataclasses import dataclass

_MAX_MEASUREMENTS_PER_SLO = 10_000


@dataclass
class SLODefinition:
    name: str
    description: str
    target: float
    unit: str                    # ratio | ms | seconds
    window_seconds: float = 3600.0


DEFAULT_SLOS: list[SLODefinition] = [
    SLODefinition("search_latency_p99", "99th percentile search latency", 1000.0, "ms"),
    SLODefinition("search_availability", "Search success rate", 0.99, "ratio"),
    SLODefinition("crawl_success_rate", "Crawl page success rate", 0.90, "ratio"),
    SLODefinition("node_uptime", "Node availability", 0.995, "ratio"),
    SLODefinition("p2p_connectivity", "Time connected to ≥1 peer", 0.95, "ratio"),
    SLODefinition("gpu_batch_latency_p99", "99th percentile device time of one hybrid query batch", 100.0, "ms"),
]


@dataclass
class SLOStatus:
    slo: SLODefinition
    current_value: float
    target: float
    met: bool
    error_budget_remaining: float
    window_start: float = 0.0


class SLOTracker:
    def __init__(self, slos: list[SLODefinition] | None = None):
        self._slos = slos or DEFAULT_SLOS
        self._samples: dict[str, deque[tuple[float, float]]] = {}
        self._ok: dict[str, int] = {}
        self._n: dict[str, int] = {}

    def record(self, slo_name: str, value: float) -> None:
        self._samples.setdefault(slo_name, deque(maxlen=_MAX_MEASUREMENTS_PER_SLO)).append((time.time(), value))

    def record_success(self, slo_name: str, success: bool) -> None:
        self._n[slo_name] = self._n.get(slo_name, 0) + 1
        self._ok[slo_name] = self._ok.get(slo_name, 0) + bool(success)

    def _one(self, slo: SLODefinition, now: float) -> SLOStatus:
        if slo.unit == "ratio":
            n = self._n.get(slo.name, 0)
            cur = self._ok.get(slo.name, 0) / n if n else 1.0
            met = cur >= slo.target
            budget = max(0.0, (cur - slo.target) / (1.0 - slo.target)) if slo.target < 1.0 else float(met)
        elif slo.unit == "ms":
            recent = sorted(v for t, v in self._samples.get(slo.name, ()) if t > now - slo.window_seconds)
            cur = recent[min(int(len(recent) * 0.99), len(recent) - 1)] if recent else 0.0
            met = cur <= slo.target
            budget = max(0.0, 1.0 - cur / slo.target) if slo.target > 0 else 1.0
        else:
            cur, met, budget = 0.0, True, 1.0
        return SLOStatus(slo, round(cur, 4), slo.target, met, round(budget, 4), now - slo.window_seconds)

    def get_status(self) -> list[SLOStatus]:
        now = time.time()
        return [self._one(s, now) for s in self._slos]

    def summary(self) -> dict[str, object]:
        st = self.get_status()
        return {"total_slos": len(st), "slos_met": sum(s.met for s in st), "slos_violated": sum(not s.met for s in st),
                "details": [{"name": s.slo.name, "target": s.target, "current": s.current_value, "met": s.met,
                             "budget_remaining": s.error_budget_remaining} for s in st]}
