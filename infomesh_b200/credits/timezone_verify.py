"""Plausibility checks on a peer's claimed timezone (gate for the off-peak x1.5 multiplier):
coarse first-octet -> UTC-offset table, +-2 h tolerance, and a >= 3 changes / 24 h suspicion rule
(reference infomesh/credits/timezone_verify.py:30-300)."""
from __future__ import annotations

import datetime
import time
from dataclasses import dataclass
from zoneinfo import ZoneInfo

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MAX_OFFSET_DIFF_HOURS = 2
MAX_TZ_CHANGES_PER_DAY = 3

# (first-octet range, UTC offset, region) — rough allocation map, first match wins
_IP_REGION_OFFSETS: list[tuple[range, float, str]] = [
    (range(1, 2), 8.0, "Asia (CN/JP/KR)"), (range(14, 15), 9.0, "Asia-Pacific"), (range(27, 28), 5.5, "South Asia"),
    (range(36, 37), 9.0, "Japan"), (range(49, 50), 9.0, "Japan"), (range(58, 59), 8.0, "East Asia"),
    (range(61, 62), 5.5, "India"), (range(101, 126), 8.0, "East Asia"), (range(175, 176), 9.0, "Asia-Pacific"),
    (range(210, 212), 9.0, "Korea"), (range(218, 222), 9.0, "Korea"), (range(2, 3), 1.0, "Europe"),
    (range(5, 6), 1.0, "Europe"), (range(31, 32), 1.0, "Europe"), (range(37, 38), 1.0, "France"),
    (range(46, 47), 3.0, "Russia"), (range(62, 63), 1.0, "Europe"), (range(77, 80), 1.0, "Europe"),
    (range(80, 82), 1.0, "Europe"), (range(83, 88), 1.0, "Europe"), (range(88, 96), 1.0, "Europe"),
    (range(145, 150), 1.0, "Europe"), (range(176, 178), 1.0, "Europe"), (range(185, 195), 1.0, "Europe"),
    (range(193, 196), 1.0, "Europe"), (range(3, 5), -5.0, "North America"), (range(6, 9), -5.0, "North America"),
    (range(12, 14), -5.0, "North America"), (range(15, 20), -5.0, "North America"),
    (range(23, 27), -5.0, "North America"), (range(32, 36), -5.0, "North America"),
    (range(38, 45), -5.0, "North America"), (range(47, 49), -5.0, "North America"),
    (range(50, 55), -5.0, "North America"), (range(63, 77), -5.0, "North America"),
    (range(96, 101), -5.0, "North America"), (range(128, 145), -5.0, "North America"),
    (range(198, 210), -5.0, "North America"), (range(150, 154), 10.0, "Oceania"), (range(202, 204), 10.0, "Oceania"),
]


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
    try:
        off = datetime.datetime.now(tz=ZoneInfo(tz_name)).utcoffset()
        return off.total_seconds() / 3600 if off is not None else 0.0
    except (KeyError, ValueError, ModuleNotFoundError):
        return 0.0


def estimate_offset_from_ip(ip_address: str) -> float | None:
    try:
        first = int(ip_address.split(".")[0])
    except (ValueError, IndexError):
        return None
    for rng, off, _ in _IP_REGION_OFFSETS:
        if first in rng:
            return off
    return None


def verify_timezone(peer_id: str, claimed_tz: str, ip_address: str) -> TimezoneCheck:
    claimed = get_timezone_offset(claimed_tz)
    est = estimate_offset_from_ip(ip_address)
    if est is None:
        return TimezoneCheck(peer_id, claimed_tz, claimed, None, None, True, "IP region unknown, cannot verify")
    diff = abs(claimed - est)
    if diff > 12:
        diff = 24 - diff
    ok = diff <= MAX_OFFSET_DIFF_HOURS
    reason = (f"offset diff {diff:.1f}h within ±{MAX_OFFSET_DIFF_HOURS}h tolerance" if ok else
              f"offset diff {diff:.1f}h exceeds ±{MAX_OFFSET_DIFF_HOURS}h (claimed {claimed_tz}={claimed:+.1f}, "
              f"IP suggests {est:+.1f})")
    if not ok:
        logger.warning("timezone_mismatch", peer_id=peer_id[:16], claimed=claimed_tz, diff_hours=diff)
    return TimezoneCheck(peer_id, claimed_tz, claimed, est, diff, ok, reason)


class TimezoneConsistencyTracker:
    def __init__(self):
        self._claims: dict[str, list[tuple[float, str]]] = {}

    def record_claim(self, peer_id: str, timezone: str, *, now: float | None = None) -> ConsistencyRecord:
        now = time.time() if now is None else now
        kept = [(ts, tz) for ts, tz in self._claims.get(peer_id, []) if ts >= now - 48 * 3600]
        kept.append((now, timezone))
        self._claims[peer_id] = kept
        recent = [tz for ts, tz in kept if ts >= now - 24 * 3600]
        changes = sum(1 for a, b in zip(recent, recent[1:]) if a != b)
        suspicious = changes >= MAX_TZ_CHANGES_PER_DAY
        if suspicious:
            logger.warning("timezone_suspicious_changes", peer_id=peer_id[:16], changes=changes)
        return ConsistencyRecord(peer_id, len(kept), len({tz for _, tz in kept}), changes, suspicious)

    def is_suspicious(self, peer_id: str) -> bool:
        claims = self._claims.get(peer_id)
        if not claims:
            return False
        now = time.time()
        recent = [tz for ts, tz in claims if ts >= now - 24 * 3600]
        return sum(1 for a, b in zip(recent, recent[1:]) if a != b) >= MAX_TZ_CHANGES_PER_DAY
