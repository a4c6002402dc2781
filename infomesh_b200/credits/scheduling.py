"""Energy-aware LLM task placement: prefer nodes in their (verified) off-peak window — they earn x1.5
(reference infomesh/credits/scheduling.py:30-311)."""
from __future__ import annotations

import datetime
from dataclasses import dataclass
from zoneinfo import ZoneInfo

from infomesh_b200.credits.timezone_verify import TimezoneConsistencyTracker, verify_timezone
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_OFF_PEAK_START: int = 23
DEFAULT_OFF_PEAK_END: int = 7
OFF_PEAK_MULTIPLIER: float = 1.5
MIN_OFF_PEAK_NODES: int = 1


@dataclass(frozen=True)
class NodeScheduleInfo:
    peer_id: str
    off_peak_start: int
    off_peak_end: int
    timezone: str
    has_llm: bool
    trust_score: float
    ip_address: str = ""


@dataclass(frozen=True)
class ScheduleDecision:
    target_peer_id: str
    is_off_peak: bool
    credit_multiplier: float
    reason: str


def is_off_peak_at(*, hour: int, start: int = DEFAULT_OFF_PEAK_START, end: int = DEFAULT_OFF_PEAK_END) -> bool:
    """``[start, end)`` with midnight wrap-around (23 -> 7)."""
    return (hour >= start or hour < end) if start > end else (start <= hour < end)


def current_hour_in_timezone(timezone: str) -> int:
    try:
        return datetime.datetime.now(tz=ZoneInfo(timezone)).hour
    except (KeyError, ValueError, ModuleNotFoundError):
        return datetime.datetime.now(tz=datetime.UTC).hour


def node_is_off_peak(node: NodeScheduleInfo, *, now_override_hour: int | None = None) -> bool:
    hour = now_override_hour if now_override_hour is not None else current_hour_in_timezone(node.timezone)
    return is_off_peak_at(hour=hour, start=node.off_peak_start, end=node.off_peak_end)


def parse_hhmm(value: str, default: int) -> int:
    """'23:00' -> 23 (config llm.off_peak_start / off_peak_end)."""
    try:
        h = int(str(value).split(":")[0])
        return h if 0 <= h <= 23 else default
    except ValueError:
        return default


class EnergyAwareScheduler:
    def __init__(self):
        self._tz = TimezoneConsistencyTracker()

    def _verified_off_peak(self, node: NodeScheduleInfo, hour: int | None) -> bool:
        if not node_is_off_peak(node, now_override_hour=hour):
            return False
        if node.ip_address:
            check = verify_timezone(node.peer_id, node.timezone, node.ip_address)
            self._tz.record_claim(node.peer_id, node.timezone)
            if not check.plausible or self._tz.is_suspicious(node.peer_id):
                logger.warning("off_peak_denied", peer_id=node.peer_id, claimed=node.timezone)
                return False
        return True

    def _split(self, nodes: list[NodeScheduleInfo], hour: int | None):
        llm = [n for n in nodes if n.has_llm]
        flags = {n.peer_id: self._verified_off_peak(n, hour) for n in llm}
        key = lambda n: n.trust_score  # noqa: E731
        return (sorted([n for n in llm if flags[n.peer_id]], key=key, reverse=True),
                sorted([n for n in llm if not flags[n.peer_id]], key=key, reverse=True))

    def schedule_llm_task(self, nodes: list[NodeScheduleInfo], *, now_override_hour: int | None = None
                          ) -> ScheduleDecision | None:
        off, on = self._split(nodes, now_override_hour)
        if off:
            b = off[0]
            return ScheduleDecision(b.peer_id, True, OFF_PEAK_MULTIPLIER,
                                    f"off-peak in {b.timezone} (trust={b.trust_score:.3f})")
        if on:
            b = on[0]
            return ScheduleDecision(b.peer_id, False, 1.0,
                                    f"on-peak, no off-peak available (trust={b.trust_score:.3f})")
        return None

    def schedule_batch(self, nodes: list[NodeScheduleInfo], task_count: int, *, now_override_hour: int | None = None
                       ) -> list[ScheduleDecision]:
        """Round-robin over off-peak nodes first; overflow to on-peak nodes only when none are off-peak."""
        off, on = self._split(nodes, now_override_hour)
        pool, off_flag = (off, True) if off else (on, False)
        if not pool or task_count <= 0:
            return []
        out = []
        for i in range(task_count):
            n = pool[i % len(pool)]
            out.append(ScheduleDecision(n.peer_id, off_flag, OFF_PEAK_MULTIPLIER if off_flag else 1.0,
                                        f"{'off' if off_flag else 'on'}-peak round-robin slot {i}"))
        return out
