"""Device-side measurement helpers used by bench.py, the scripts and the observability layer:
NVTX ranges, CUDA-event timers with max-over-ranks reduction, roofline fractions against ``MEASURED_PEAKS.json``."""
from __future__ import annotations

import contextlib
import json
from dataclasses import dataclass
from pathlib import Path

import torch


@contextlib.contextmanager
def nvtx(name: str):
    """NVTX range (visible in nsys / ncu timelines); free when no profiler is attached."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class CudaTimer:
    """``with CudaTimer() as t: ...`` -> ``t.ms`` (device time on the current stream, synchronised on exit)."""

    def __enter__(self) -> "CudaTimer":
        self._a, self._b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        self._a.record()
        return self

    def __exit__(self, *exc: object) -> None:
        self._b.record()
        torch.cuda.synchronize()
        self.ms = self._a.elapsed_time(self._b)

    def max_over_ranks(self) -> float:
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return self.ms
        t = torch.tensor([self.ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


@dataclass(frozen=True)
class Peaks:
    hbm_gbps: float = 6573.8          # fallbacks = the values measured on the round-1 B200
    bf16_tflops: float = 1456.6

    @classmethod
    def load(cls, path: str | Path = "MEASURED_PEAKS.json") -> "Peaks":
        try:
            raw = json.loads(Path(path).read_text())
        except (OSError, ValueError):
            return cls()

        def find(obj, *needles):
            if isinstance(obj, dict):
                for k, v in obj.items():
                    if isinstance(v, (int, float)) and all(n in k.lower() for n in needles):
                        return float(v)
                    got = find(v, *needles)
                    if got:
                        return got
            return 0.0

        return cls(find(raw, "hbm") or find(raw, "copy") or cls.hbm_gbps, find(raw, "bf16", "sustain") or find(raw, "bf16") or cls.bf16_tflops)


def roofline(*, seconds: float, bytes_moved: float = 0.0, flops: float = 0.0, peaks: Peaks | None = None) -> dict[str, float]:
    """Achieved GB/s and TFLOP/s and their fraction of the measured peaks; ``bound`` names the tighter ceiling."""
    p = peaks or Peaks.load()
    gbps, tfs = bytes_moved / seconds / 1e9 if seconds else 0.0, flops / seconds / 1e12 if seconds else 0.0
    fm, fc = gbps / p.hbm_gbps, tfs / p.bf16_tflops
    floor = max(bytes_moved / (p.hbm_gbps * 1e9), flops / (p.bf16_tflops * 1e12))
    return {"gbps": gbps, "tflops": tfs, "frac_hbm": fm, "frac_bf16": fc, "speed_of_light_s": floor,
            "frac_of_light": floor / seconds if seconds else 0.0, "bound": "memory" if fm >= fc else "compute"}
