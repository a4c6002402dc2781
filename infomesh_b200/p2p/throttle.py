"""Bandwidth limits for P2P traffic (``network.upload_limit_mbps`` / ``download_limit_mbps``; 0 = unlimited).

Contract (reference infomesh/p2p/throttle.py): a long-run byte rate per direction with one second of burst; callers
``await acquire_*(nbytes)`` before moving data and get back how long they were held.

Implementation: a virtual-clock meter (GCRA) instead of a refilled token counter.  Each direction keeps one number, the
time up to which its bandwidth is already spoken for.  A transfer of ``n`` bytes books ``n / rate`` seconds starting at
that point (but no earlier than one burst-second in the past -- idle time beyond the burst is not banked), and the caller
sleeps until its booking ends.  Booking is a handful of arithmetic steps with no ``await`` in between, so concurrent
tasks of one event loop serialise naturally without a lock, and an oversized transfer simply books a long slot."""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass

BYTES_PER_MEGABIT = 125_000
BURST_SECONDS = 1.0


@dataclass
class BandwidthStats:
    upload_bytes: int = 0
    download_bytes: int = 0
    upload_waits: int = 0
    download_waits: int = 0


class BandwidthBucket:
    """One direction's meter (the name is kept from the token-bucket formulation it replaces)."""

    def __init__(self, rate_mbps: float, burst_seconds: float = BURST_SECONDS):
        self._bytes_per_sec = float(rate_mbps) * BYTES_PER_MEGABIT
        self._burst = float(burst_seconds)
        self._booked_until = float("-inf")          # virtual clock: bandwidth is reserved up to here

    @property
    def rate_bytes_per_sec(self) -> float:
        return self._bytes_per_sec

    def reserve(self, nbytes: int, now: float | None = None) -> float:
        """Book ``nbytes`` and return the delay (seconds from ``now``) after which the transfer conforms to the rate."""
        if nbytes <= 0 or self._bytes_per_sec <= 0:
            return 0.0
        now = time.monotonic() if now is None else now
        start = max(self._booked_until, now - self._burst)
        self._booked_until = start + nbytes / self._bytes_per_sec
        return max(0.0, self._booked_until - now)

    async def acquire(self, nbytes: int) -> float:
        """Returns the seconds spent waiting."""
        delay = self.reserve(nbytes)
        if delay > 0:
            await asyncio.sleep(delay)
        return delay


class BandwidthThrottle:
    """Upload and download meters plus byte / wait counters."""

    def __init__(self, upload_mbps: float = 5.0, download_mbps: float = 10.0):
        self._meters = {"upload": BandwidthBucket(upload_mbps) if upload_mbps > 0 else None,
                        "download": BandwidthBucket(download_mbps) if download_mbps > 0 else None}
        self._stats = BandwidthStats()

    @property
    def stats(self) -> BandwidthStats:
        return self._stats

    async def _pass(self, direction: str, nbytes: int) -> float:
        setattr(self._stats, f"{direction}_bytes", getattr(self._stats, f"{direction}_bytes") + nbytes)
        meter = self._meters[direction]
        held = await meter.acquire(nbytes) if meter is not None else 0.0
        if held > 0:
            setattr(self._stats, f"{direction}_waits", getattr(self._stats, f"{direction}_waits") + 1)
        return held

    async def acquire_upload(self, nbytes: int) -> float:
        return await self._pass("upload", nbytes)

    async def acquire_download(self, nbytes: int) -> float:
        return await self._pass("download", nbytes)
