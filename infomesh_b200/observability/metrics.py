"""Observability: optional OpenTelemetry setup, an in-process Prometheus-format collector, distributed query traces,
Grafana dashboard + alert-rule generators (reference infomesh/observability/metrics.py:24-447), plus CUDA-event
kernel spans (:class:`GpuTimer`) so device stages show up in the same collector and traces."""
from __future__ import annotations

import re
import time
from collections import defaultdict, deque
from contextlib import contextmanager
from dataclasses import dataclass, field
from threading import Lock
from typing import Any

_UNSAFE = re.compile(r"[^a-zA-Z0-9_]")


def setup_otel(service_name: str = "infomesh", *, endpoint: str | None = None) -> bool:
    try:
        from opentelemetry import trace
        from opentelemetry.sdk.resources import Resource
        from opentelemetry.sdk.trace import TracerProvider
    except ImportError:
        return False
    provider = TracerProvider(resource=Resource.create({"service.name": service_name}))
    trace.set_tracer_provider(provider)
    if endpoint:
        try:
            from opentelemetry.exporter.otlp.proto.grpc.trace_exporter import OTLPSpanExporter
            from opentelemetry.sdk.trace.export import BatchSpanProcessor

            provider.add_span_processor(BatchSpanProcessor(OTLPSpanExporter(endpoint=endpoint)))
        except ImportError:
            return False
    return True


def get_tracer(name: str = "infomesh") -> Any:
    try:
        from opentelemetry import trace

        return trace.get_tracer(name)
    except ImportError:
        return None


def _sanitize_metric_name(name: str) -> str:
    return _UNSAFE.sub("_", name)


class MetricsCollector:
    """Counters, gauges and sliding-window (1000 samples) summaries; thread-safe."""

    def __init__(self):
        self._lock = Lock()
        self._counters: dict[str, float] = defaultdict(float)
        self._gauges: dict[str, float] = {}
        self._hist: dict[str, deque[float]] = {}
        self._start = time.time()

    def inc(self, name: str, value: float = 1.0) -> None:
        with self._lock:
            self._counters[name] += value

    def set_gauge(self, name: str, value: float) -> None:
        with self._lock:
            self._gauges[name] = value

    def observe(self, name: str, value: float) -> None:
        with self._lock:
            self._hist.setdefault(name, deque(maxlen=1000)).append(value)

    @contextmanager
    def timer(self, name: str):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.observe(name, (time.perf_counter() - t0) * 1000)

    def format_prometheus(self) -> str:
        with self._lock:
            out: list[str] = []
            for name, v in sorted(self._counters.items()):
                n = _sanitize_metric_name(name)
                out += [f"# TYPE {n} counter", f"{n} {v}"]
            for name, v in sorted(self._gauges.items()):
                n = _sanitize_metric_name(name)
                out += [f"# TYPE {n} gauge", f"{n} {v}"]
            for name, vals in sorted(self._hist.items()):
                if not vals:
                    continue
                n, s = _sanitize_metric_name(name), sorted(vals)
                out += [f"# TYPE {n} summary", f'{n}{{quantile="0.5"}} {s[len(s) // 2]:.3f}',
                        f'{n}{{quantile="0.99"}} {s[min(int(len(s) * 0.99), len(s) - 1)]:.3f}',
                        f"{n}_count {len(s)}", f"{n}_sum {sum(s):.3f}", f"{n}_avg {sum(s) / len(s):.3f}"]
            out += ["# TYPE infomesh_uptime_seconds gauge", f"infomesh_uptime_seconds {time.time() - self._start:.0f}"]
            return "\n".join(out) + "\n"

    def to_dict(self) -> dict[str, object]:
        with self._lock:
            return {"counters": dict(self._counters), "gauges": dict(self._gauges),
                    "histograms": {k: {"count": len(v), "sum": sum(v), "avg": sum(v) / len(v)} for k, v in self._hist.items() if v},
                    "uptime_seconds": time.time() - self._start}


_global: MetricsCollector | None = None


def get_collector() -> MetricsCollector:
    global _global
    if _global is None:
        _global = MetricsCollector()
    return _global


class GpuTimer:
    """CUDA-event span on the current stream.  ``with GpuTimer(collector, "gpu_rerank_ms"):`` records device time
    without a host sync on entry; the sample is resolved lazily by :meth:`flush` (or at the next span)."""

    def __init__(self, collector: MetricsCollector | None = None, name: str = "gpu_ms"):
        self.collector, self.name = collector or get_collector(), name
        self._pending: list[tuple[Any, Any]] = []

    def __enter__(self) -> "GpuTimer":
        import torch

        self.flush(block=False)
        self._a, self._b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._a.record()
        return self

    def __exit__(self, *exc: object) -> None:
        self._b.record()
        self._pending.append((self._a, self._b))

    def flush(self, block: bool = True) -> None:
        keep = []
        for a, b in self._pending:
            if block:
                b.synchronize()
            if b.query():
                self.collector.observe(self.name, a.elapsed_time(b))
            else:
                keep.append((a, b))
        self._pending = keep


def configure_log_forwarding(*, format: str = "json", output: str = "stdout") -> dict[str, str]:
    procs = ("add_log_level,TimeStamper(fmt='iso'),JSONRenderer()" if format == "json" else "ConsoleRenderer()")
    return {"format": format, "output": output, "level": "INFO", "processors": procs}


@dataclass
class QuerySpan:
    span_id: str
    peer_id: str
    operation: str
    start_time: float = 0.0
    end_time: float = 0.0
    latency_ms: float = 0.0
    metadata: dict[str, str] = field(default_factory=dict)


@dataclass
class QueryTrace:
    trace_id: str
    query: str
    spans: list[QuerySpan] = field(default_factory=list)
    total_latency_ms: float = 0.0

    def add_span(self, span: QuerySpan) -> None:
        self.spans.append(span)
        self.total_latency_ms = sum(s.latency_ms for s in self.spans)

    def to_dict(self) -> dict[str, object]:
        return {"trace_id": self.trace_id, "query": self.query, "total_latency_ms": self.total_latency_ms,
                "spans": [{"span_id": s.span_id, "peer_id": s.peer_id, "operation": s.operation, "latency_ms": s.latency_ms,
                           "metadata": s.metadata} for s in self.spans]}


@dataclass
class BenchmarkResult:
    name: str
    iterations: int
    total_ms: float
    avg_ms: float
    min_ms: float
    max_ms: float
    p50_ms: float
    p95_ms: float
    p99_ms: float


def run_benchmark(name: str, fn: Any, *, iterations: int = 100) -> BenchmarkResult:
    t = []
    for _ in range(max(1, iterations)):
        t0 = time.monotonic()
        if callable(fn):
            fn()
        t.append((time.monotonic() - t0) * 1000)
    t.sort()
    n = len(t)
    at = lambda p: round(t[min(int(n * p), n - 1)], 2)  # noqa: E731
    return BenchmarkResult(name, n, round(sum(t), 2), round(sum(t) / n, 2), round(t[0], 2), round(t[-1], 2), at(0.5), at(0.95), at(0.99))


def generate_grafana_dashboard() -> dict[str, object]:
    def panel(title, kind, expr, x, y, w, h):
        return {"title": title, "type": kind, "targets": [{"expr": expr}], "gridPos": {"h": h, "w": w, "x": x, "y": y}}

    panels = [
        panel("Search Queries / sec", "graph", "rate(infomesh_search_total[5m])", 0, 0, 12, 8),
        panel("Search Latency (avg)", "graph", "infomesh_search_latency_ms_avg", 12, 0, 12, 8),
        panel("Documents Indexed", "stat", "infomesh_documents_indexed", 0, 8, 6, 4),
        panel("P2P Peers Connected", "stat", "infomesh_p2p_peers", 6, 8, 6, 4),
        panel("Credit Balance", "stat", "infomesh_credit_balance", 12, 8, 6, 4),
        panel("Crawl Rate / min", "graph", "rate(infomesh_crawl_total[5m]) * 60", 0, 12, 12, 8),
        panel("GPU batch latency p99 (ms)", "graph", 'infomesh_gpu_batch_ms{quantile="0.99"}', 12, 12, 12, 8),
        panel("GPU index residency (GB)", "stat", "infomesh_gpu_index_bytes / 1e9", 18, 8, 6, 4),
    ]
    return {"dashboard": {"title": "InfoMesh Monitoring", "tags": ["infomesh", "search", "p2p", "gpu"], "timezone": "browser",
                          "panels": panels, "refresh": "30s", "time": {"from": "now-1h", "to": "now"}}}


def generate_alert_rules() -> list[dict[str, object]]:
    def rule(alert, expr, dur, sev, summary):
        return {"alert": alert, "expr": expr, "for": dur, "labels": {"severity": sev}, "annotations": {"summary": summary}}

    return [
        rule("HighSearchLatency", "infomesh_search_latency_ms_avg > 2000", "5m", "warning", "Search latency exceeds 2s"),
        rule("CrawlRateDropped", "rate(infomesh_crawl_total[10m]) == 0", "10m", "warning", "No crawls in 10 minutes"),
        rule("LowDiskSpace", "infomesh_disk_free_mb < 500", "5m", "critical", "Disk space below 500MB"),
        rule("NoPeersConnected", "infomesh_p2p_peers == 0", "15m", "warning", "No P2P peers connected"),
        rule("CreditsDepleted", "infomesh_credit_balance < 0", "1h", "info", "Credit balance is negative"),
        rule("GpuBatchLatencyHigh", 'infomesh_gpu_batch_ms{quantile="0.99"} > 100', "5m", "warning", "GPU query batch p99 above 100 ms"),
    ]
