"""Micro-benchmark helpers: wall-clock timing for host code (reference infomesh/benchmarks.py:14-99) and a CUDA-event
variant for device work (warm-up, synchronise on both sides — the timing rules of the GPU plane)."""
from __future__ import annotations

import statistics
import time
from dataclasses import dataclass, field
from typing import Any, Callable


@dataclass
class BenchmarkResult:
    name: str
    iterations: int
    total_ms: float
    avg_ms: float
    median_ms: float
    p95_ms: float
    p99_ms: float
    min_ms: float
    max_ms: float
    ops_per_sec: float

    def __str__(self) -> str:
        return (f"{self.name}: avg={self.avg_ms:.1f}ms p95={self.p95_ms:.1f}ms "
                f"({self.ops_per_sec:.0f} ops/s, {self.iterations} iters)")


@dataclass
class BenchmarkSuite:
    results: list[BenchmarkResult] = field(default_factory=list)
    started_at: float = field(default_factory=time.time)
    finished_at: float = 0.0

    def add(self, result: BenchmarkResult) -> None:
        self.results.append(result)
        self.finished_at = time.time()

    def report(self) -> str:
        return "\n".join(["InfoMesh Performance Benchmark", "=" * 40, *map(str, self.results)])


def _summarise(name: str, timings: list[float]) -> BenchmarkResult:
    if not timings:
        return BenchmarkResult(name, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    t = sorted(timings)
    n, total = len(t), sum(t)

    def pct(p: float) -> float:
        return round(t[min(int(n * p), n - 1)], 3)

    return BenchmarkResult(name, n, round(total, 3), round(statistics.mean(t), 3), round(statistics.median(t), 3), pct(0.95),
                           pct(0.99), round(t[0], 3), round(t[-1], 3), round(1000 * n / total, 1) if total > 0 else 0.0)


def benchmark(func: Callable[..., Any], *args: Any, iterations: int = 100, name: str = "", **kwargs: Any) -> BenchmarkResult:
    timings = []
    for _ in range(iterations):
        t0 = time.perf_counter()
        func(*args, **kwargs)
        timings.append((time.perf_counter() - t0) * 1000)
    return _summarise(name or getattr(func, "__name__", str(func)), timings)


def benchmark_cuda(func: Callable[[], Any], *, iterations: int = 50, warmup: int = 5, name: str = "",
                   flush_l2: bool = True) -> BenchmarkResult:
    """Per-iteration CUDA-event timing on the current stream; optionally rewrites a > L2-sized buffer between
    iterations so the measured kernel starts from cold caches."""
    import torch

    for _ in range(warmup):
        func()
    torch.cuda.synchronize()
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush_l2 else None
    timings = []
    for _ in range(iterations):
        if scrub is not None:
            scrub.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        func()
        b.record()
        torch.cuda.synchronize()
        timings.append(a.elapsed_time(b))
    return _summarise(name or getattr(func, "__name__", "cuda_op"), timings)
