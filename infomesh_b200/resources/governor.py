"""Load governor: samples CPU / memory / process RSS (and GPU memory when a device is present) and derives a
degrade level 0..4 plus a crawl throttle factor (reference infomesh/resources/governor.py:40-334).

Level semantics: WARNING -> LLM off + crawl paused, OVERLOADED -> no remote search, SEVERE -> read-only,
DEFENSIVE -> minimal operation.  Thresholds are the reference's (60/80/90/95 % CPU, 70/85/90/95 % memory,
0.75/0.9/1.0/1.2 x process limit); GPU memory pressure maps through the same ladder as system memory."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import time
from dataclasses import dataclass
from enum import IntEnum

from infomesh_b200.resources.profiles import ResourceProfile
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

try:
    import psutil
except ImportError:  # pragma: no cover
    psutil = None

CPU_HIGH_PCT = 80
CPU_LOW_PCT = 30
MEMORY_HIGH_PCT = 85
NETWORK_HIGH_FACTOR = 0.9


class DegradeLevel(IntEnum):
    NORMAL = 0
    WARNING = 1
    OVERLOADED = 2
    SEVERE = 3
    DEFENSIVE = 4


@dataclass
class GovernorState:
    degrade_level: DegradeLevel = DegradeLevel.NORMAL
    cpu_percent: float = 0.0
    memory_percent: float = 0.0
    process_memory_mb: float = 0.0
    process_memory_limit_mb: int = 0
    process_memory_ratio: float = 0.0
    throttle_factor: float = 1.0
    last_check: float = 0.0
    checks_performed: int = 0
    gpu_memory_percent: float = 0.0      # new in this build; after the reference's fields so positional construction matches


_LADDER = (  # (level, cpu >, mem >, process ratio >=)
    (DegradeLevel.DEFENSIVE, 95, 95, 1.2),
    (DegradeLevel.SEVERE, 90, 90, 1.0),
    (DegradeLevel.OVERLOADED, CPU_HIGH_PCT, MEMORY_HIGH_PCT, 0.9),
    (DegradeLevel.WARNING, 60, 70, 0.75),
)


def classify(cpu: float, mem: float, process_ratio: float) -> DegradeLevel:
    for level, c, m, r in _LADDER:
        if cpu > c or mem > m or process_ratio >= r:
            return level
    return DegradeLevel.NORMAL


def throttle_for(level: DegradeLevel, cpu: float) -> float:
    if level >= DegradeLevel.SEVERE:
        return 0.0
    if level == DegradeLevel.OVERLOADED:
        return 0.25
    if level == DegradeLevel.WARNING:
        return 0.5
    if cpu > CPU_LOW_PCT:
        return max(0.3, 1.0 - (cpu - CPU_LOW_PCT) / (CPU_HIGH_PCT - CPU_LOW_PCT) * 0.7)
    return 1.0


class ResourceGovernor:
    def __init__(self, profile: ResourceProfile):
        self._profile = profile
        self._state = GovernorState()
        self._net_prev: tuple[int, int] | None = None
        self._net_stamp = 0.0
        self._process = psutil.Process(os.getpid()) if psutil else None

    profile = property(lambda self: self._profile)
    state = property(lambda self: self._state)
    degrade_level = property(lambda self: self._state.degrade_level)
    throttle_factor = property(lambda self: self._state.throttle_factor)
    cpu_percent = property(lambda self: self._state.cpu_percent)
    memory_percent = property(lambda self: self._state.memory_percent)
    process_memory_mb = property(lambda self: self._state.process_memory_mb)
    should_throttle_crawl = property(lambda self: self._state.throttle_factor < 1.0)
    should_pause_crawl = property(lambda self: self._state.degrade_level >= DegradeLevel.WARNING)
    should_disable_llm = property(lambda self: self._state.degrade_level >= DegradeLevel.WARNING)
    should_disable_remote_search = property(lambda self: self._state.degrade_level >= DegradeLevel.OVERLOADED)
    is_read_only = property(lambda self: self._state.degrade_level >= DegradeLevel.SEVERE)

    @property
    def effective_max_concurrent(self) -> int:
        return max(1, int(self._profile.max_concurrent_crawl * self._state.throttle_factor))

    # ------------------------------------------------------------------ OS priority
    def apply_os_priority(self) -> None:
        if sys.platform == "win32":
            return
        try:
            cur = os.nice(0)
            if cur < self._profile.cpu_nice:
                os.nice(self._profile.cpu_nice - cur)
        except OSError as exc:
            logger.debug("nice_failed", error=str(exc))
        if sys.platform == "linux" and shutil.which("ionice"):
            cls = {"low": ["-c", "3"], "high": ["-c", "2", "-n", "0"]}.get(self._profile.disk_io_priority, ["-c", "2", "-n", "4"])
            try:
                subprocess.run(["ionice", *cls, "-p", str(os.getpid())], check=False, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            except OSError as exc:
                logger.debug("ionice_failed", error=str(exc))

    # ------------------------------------------------------------------ sampling
    @staticmethod
    def _sample_cpu() -> float:
        try:
            return float(psutil.cpu_percent(interval=0)) if psutil else 0.0
        except Exception:  # noqa: BLE001
            return 0.0

    @staticmethod
    def _sample_memory() -> float:
        try:
            return float(psutil.virtual_memory().percent) if psutil else 0.0
        except Exception:  # noqa: BLE001
            return 0.0

    def _sample_process_memory_mb(self) -> float:
        try:
            return float(self._process.memory_info().rss / 2 ** 20) if self._process else 0.0
        except Exception:  # noqa: BLE001
            return 0.0

    @staticmethod
    def _sample_gpu_memory() -> float:
        try:
            import torch

            if not torch.cuda.is_available():
                return 0.0
            free, total = torch.cuda.mem_get_info()
            return 100.0 * (1.0 - free / total)
        except Exception:  # noqa: BLE001
            return 0.0

    def check_and_adjust(self) -> GovernorState:
        st = self._state
        st.last_check = time.monotonic()
        st.checks_performed += 1
        st.cpu_percent, st.memory_percent = self._sample_cpu(), self._sample_memory()
        st.process_memory_mb = self._sample_process_memory_mb()
        st.gpu_memory_percent = self._sample_gpu_memory()
        limit = self._profile.memory_limit_mb
        st.process_memory_limit_mb = limit
        st.process_memory_ratio = st.process_memory_mb / limit if limit > 0 else 0.0
        # GPU pressure only counts beyond the share the profile grants the resident index
        gpu_excess = st.gpu_memory_percent if st.gpu_memory_percent > 100.0 * self._profile.gpu_memory_fraction else 0.0
        level = classify(st.cpu_percent, max(st.memory_percent, gpu_excess), st.process_memory_ratio)
        if level != st.degrade_level:
            logger.info("governor_level_change", old=int(st.degrade_level), new=int(level), cpu=round(st.cpu_percent, 1),
                        mem=round(st.memory_percent, 1))
        st.degrade_level = level
        st.throttle_factor = throttle_for(level, st.cpu_percent)
        return st

    def sample_network_mbps(self) -> tuple[float, float]:
        if not psutil:
            return 0.0, 0.0
        try:
            c = psutil.net_io_counters()
        except Exception:  # noqa: BLE001
            return 0.0, 0.0
        if c is None:
            return 0.0, 0.0
        now = time.monotonic()
        prev, self._net_prev = self._net_prev, (c.bytes_sent, c.bytes_recv)
        dt, self._net_stamp = now - self._net_stamp, now
        if prev is None or dt < 0.1:
            return 0.0, 0.0
        return (round((c.bytes_sent - prev[0]) * 8 / (dt * 1e6), 3), round((c.bytes_recv - prev[1]) * 8 / (dt * 1e6), 3))
