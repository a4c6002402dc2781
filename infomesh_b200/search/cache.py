"""Thread-safe LRU + TTL cache of rendered search responses (reference infomesh/search/cache.py:38-183):
1000 entries, 300 s TTL, key = sha256(lower(query):limit)[:16]."""
from __future__ import annotations

import threading
import time
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any

from infomesh_b200.hashing import short_hash


@dataclass
class CacheStats:
    hits: int = 0
    misses: int = 0
    evictions: int = 0
    size: int = 0
    max_size: int = 0

    @property
    def hit_rate(self) -> float:
        total = self.hits + self.misses
        return self.hits / total if total else 0.0


class QueryCache:
    def __init__(self, max_size: int = 1000, ttl_seconds: float = 300.0):
        self._max = max(1, int(max_size))
        self._ttl = float(ttl_seconds)
        self._data: OrderedDict[str, tuple[float, Any]] = OrderedDict()
        self._lock = threading.Lock()
        self._hits = self._misses = self._evictions = 0

    @staticmethod
    def make_key(query: str, limit: int = 10, **extra: Any) -> str:
        base = f"{query.strip().lower()}:{limit}"
        if extra:
            base += ":" + ":".join(f"{k}={extra[k]}" for k in sorted(extra))
        return short_hash(base, 16)

    def get(self, key: str, *, now: float | None = None) -> Any | None:
        now = time.monotonic() if now is None else now
        with self._lock:
            item = self._data.get(key)
            if item is None:
                self._misses += 1
                return None
            stamp, value = item
            if now - stamp > self._ttl:
                del self._data[key]
                self._misses += 1
                return None
            self._data.move_to_end(key)
            self._hits += 1
            return value

    def put(self, key: str, value: Any, *, now: float | None = None) -> None:
        now = time.monotonic() if now is None else now
        with self._lock:
            self._data[key] = (now, value)
            self._data.move_to_end(key)
            while len(self._data) > self._max:
                self._data.popitem(last=False)
                self._evictions += 1

    def invalidate(self, key: str | None = None) -> None:
        with self._lock:
            if key is None:
                self._data.clear()
            else:
                self._data.pop(key, None)

    clear = invalidate

    def __len__(self) -> int:
        return len(self._data)

    @property
    def stats(self) -> CacheStats:
        with self._lock:
            return CacheStats(self._hits, self._misses, self._evictions, len(self._data), self._max)
