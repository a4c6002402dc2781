"""Token-bucket bandwidth limits for P2P traffic (config ``network.upload_limit_mbps`` / ``download_limit_mbps``);
one second of burst, oversized transfers are paid for in bucket-sized chunks
(reference infomesh/p2p/throttle.py:43-169)."""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass

_BYTES_PER_MBIT = 1_000_000 / 8


@dataclass
class BandwidthStats:
    upload_bytes: int = 0
    download_bytes: int = 0
    upload_waits: int = 0
    download_waits: int = 0


class BandwidthBucket:
    def __init__(self, rate_mbps: float):
        self._rate = rate_mbps * _BYTES_PER_MBIT      # bytes / second
        self._tokens = self._rate
        self._stamp = time.monotonic()
        self._lock = asyncio.Lock()

    @property
    def rate_bytes_per_sec(self) -> float:
        return self._rate

    def _refill(self) -> None:
        now = time.monotonic()
        self._tokens = min(self._rate, self._tokens + (now - self._stamp) * self._rate)
        self._stamp = now

    async def acquire(self, nbytes: int) -> float:
        """Returns the seconds spent waiting."""
        if nbytes <= 0 or self._rate <= 0:
            return 0.0
        waited, left = 0.0, float(nbytes)
        while left > 0:
            chunk = min(left, self._rate)
            async with self._lock:
                self._refill()
                while self._tokens < chunk:
                    nap = (chunk - self._tokens) / self._rate
                    await asyncio.sleep(nap)
                    waited += nap
                    self._refill()
                self._tokens -= chunk
            left -= chunk
        return waited


class BandwidthThrottle:
    """0 Mbps disables the limit for that direction."""

    def __init__(self, upload_mbps: float = 5.0, download_mbps: float = 10.0):
        self._up = BandwidthBucket(upload_mbps) if upload_mbps > 0 else None
        self._down = BandwidthBucket(download_mbps) if download_mbps > 0 else None
        self._stats = BandwidthStats()

    @property
    def stats(self) -> BandwidthStats:
        return self._stats

    async def acquire_upload(self, nbytes: int) -> float:
        self._stats.upload_bytes += nbytes
        w = await self._up.acquire(nbytes) if self._up else 0.0
        self._stats.upload_waits += 1 if w > 0 else 0
        return w

    async def acquire_download(self, nbytes: int) -> float:
        self._stats.download_bytes += nbytes
        w = await self._down.acquire(nbytes) if self._down else 0.0
        self._stats.download_waits += 1 if w > 0 else 0
        return w
