"""Is a peer's claimed timezone believable?  Gate for the off-peak x1.5 credit multiplier.

Contract (reference infomesh/credits/timezone_verify.py): the first octet of the peer's IPv4 address gives a coarse
regional UTC offset; a claim within +-2 h of it (measured around the date line) is plausible, an unknown region cannot be
verified and passes; a peer that switches timezone three or more times within 24 h is suspicious.

Implementation: the regional allocation is stored as inclusive octet spans grouped by UTC offset and expanded once into a
256-slot lookup table; claim histories are bounded deques whose switch count is recomputed over the 24 h window only."""
from __future__ import annotations

import datetime
import time
from collections import deque
from dataclasses import dataclass
from zoneinfo import ZoneInfo

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MAX_OFFSET_DIFF_HOURS = 2
MAX_TZ_CHANGES_PER_DAY = 3
_DAY = 24 * 3600.0
_HISTORY_HORIZON = 2 * _DAY
_HISTORY_CAP = 256

# UTC offset (hours) -> inclusive first-octet spans of address space predominantly allocated in that band
_SPANS_BY_OFFSET: dict[float, tuple[tuple[int, int], ...]] = {
    -5.0: ((3, 4), (6, 8), (12, 13), (15, 19), (23, 26), (32, 35), (38, 44), (47, 48), (50, 54), (63, 76), (96, 100), (128, 144),
           (198, 209)),                                                                            # North America
    1.0: ((2, 2), (5, 5), (31, 31), (37, 37), (62, 62), (77, 81), (83, 95), (145, 149), (176, 177), (185, 195)),   # Europe
    3.0: ((46, 46),),                                                                              # Russia
    5.5: ((27, 27), (61, 61)),                                                                     # South Asia
    8.0: ((1, 1), (58, 58), (101, 125)),                                                           # East Asia
    9.0: ((14, 14), (36, 36), (49, 49), (175, 175), (210, 211), (218, 221)),                       # Japan / Korea
    10.0: ((150, 153),),                                                                           # Oceania
}


def _expand() -> list[float | None]:
    table: list[float | None] = [None] * 256
    for offset, spans in _SPANS_BY_OFFSET.items():
        for lo, hi in spans:
            for octet in range(lo, hi + 1):
                table[octet] = offset
    return table


_OFFSET_OF_OCTET = _expand()


@dataclass(frozen=True)
class TimezoneCheck:
    peer_id: str
    claimed_tz: str
    claimed_offset_hours: float
    estimated_offset_hours: float | None
    offset_diff_hours: float | None
    plausible: bool
    reason: str


@dataclass(frozen=True)
class ConsistencyRecord:
    peer_id: str
    claim_count: int
    unique_timezones: int
    changes_in_24h: int
    suspicious: bool


def get_timezone_offset(tz_name: str) -> float:
    """Current UTC offset of an IANA zone in hours; unknown names count as UTC."""
    try:
        delta = datetime.datetime.now(tz=ZoneInfo(tz_name)).utcoffset()
    except (KeyError, ValueError, ModuleNotFoundError):
        return 0.0
    return delta.total_seconds() / 3600.0 if delta is not None else 0.0


def estimate_offset_from_ip(ip_address: str) -> float | None:
    head = ip_address.partition(".")[0]
    if not head.isdigit() or not 0 <= int(head) <= 255:
        return None
    return _OFFSET_OF_OCTET[int(head)]


def _circular_gap(a: float, b: float) -> float:
    """Distance between two UTC offsets measured the short way around the 24 h circle."""
    gap = abs(a - b) % 24.0
    return min(gap, 24.0 - gap)


def verify_timezone(peer_id: str, claimed_tz: str, ip_address: str) -> TimezoneCheck:
    claimed = get_timezone_offset(claimed_tz)
    regional = estimate_offset_from_ip(ip_address)
    if regional is None:
        return TimezoneCheck(peer_id, claimed_tz, claimed, None, None, True, "IP region unknown, cannot verify")
    gap = _circular_gap(claimed, regional)
    plausible = gap <= MAX_OFFSET_DIFF_HOURS
    if plausible:
        why = f"offset diff {gap:.1f}h within ±{MAX_OFFSET_DIFF_HOURS}h tolerance"
    else:
        why = (f"offset diff {gap:.1f}h exceeds ±{MAX_OFFSET_DIFF_HOURS}h (claimed {claimed_tz}={claimed:+.1f}, "
               f"IP suggests {regional:+.1f})")
        logger.warning("timezone_mismatch", peer_id=peer_id[:16], claimed=claimed_tz, diff_hours=gap)
    return TimezoneCheck(peer_id, claimed_tz, claimed, regional, gap, plausible, why)


def _switches(zones: list[str]) -> int:
    return sum(1 for before, after in zip(zones, zones[1:]) if before != after)


class TimezoneConsistencyTracker:
    """Remembers each peer's recent claims (two days, bounded) and counts zone switches inside the last day."""

    def __init__(self):
        self._history: dict[str, deque[tuple[float, str]]] = {}

    def _recent_zones(self, peer_id: str, now: float) -> list[str]:
        return [zone for stamp, zone in self._history.get(peer_id, ()) if stamp >= now - _DAY]

    def record_claim(self, peer_id: str, timezone: str, *, now: float | None = None) -> ConsistencyRecord:
        now = time.time() if now is None else now
        log = self._history.setdefault(peer_id, deque(maxlen=_HISTORY_CAP))
        while log and log[0][0] < now - _HISTORY_HORIZON:
            log.popleft()
        log.append((now, timezone))
        changes = _switches(self._recent_zones(peer_id, now))
        flagged = changes >= MAX_TZ_CHANGES_PER_DAY
        if flagged:
            logger.warning("timezone_suspicious_changes", peer_id=peer_id[:16], changes=changes)
        return ConsistencyRecord(peer_id, len(log), len({zone for _, zone in log}), changes, flagged)

    def is_suspicious(self, peer_id: str) -> bool:
        return _switches(self._recent_zones(peer_id, time.time())) >= MAX_TZ_CHANGES_PER_DAY
