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
    def total(self) -> int:
        return self.hits + self.misses

    @property
    def hit_rate(self) -> float:
        return self.hits / self.total if self.total else 0.0


@dataclass
class CacheEntry:
    results: Any
    timestamp: float
    hit_count: int = 0


_UNSET: Any = object()


class QueryCache:
    """Two calling conventions on one store:

    * reference style — ``get(query, limit)`` / ``put(query, limit, results)`` / ``invalidate(query, limit)``; the key is
      derived with :meth:`make_key`;
    * key style — ``get(key)`` / ``put(key, value)`` / ``invalidate(key)``, used by the MCP runtime whose keys also cover
      filters and output format (``make_key(query, limit, **extra)``).
    """

    def __init__(self, max_size: int = 1000, ttl_seconds: float = 300.0):
        self._max = max(1, int(max_size))
        self._ttl = float(ttl_seconds)
        self._data: OrderedDict[str, CacheEntry] = OrderedDict()
        self._lock = threading.Lock()
        self._hits = self._misses = self._evictions = 0

    @staticmethod
    def make_key(query: str, limit: int = 10, **extra: Any) -> str:
        base = f"{query.strip().lower()}:{limit}"
        if extra:
            base += ":" + ":".join(f"{k}={extra[k]}" for k in sorted(extra))
        return short_hash(base, 16)

    _make_key = make_key

    def _key(self, query_or_key: str, limit: int | None) -> str:
        return query_or_key if limit is None else self.make_key(query_or_key, limit)

    def get(self, query: str, limit: int | None = None, *, now: float | None = None) -> Any | None:
        """``get(query, limit)`` as in the reference, or ``get(key)`` with a key made by :meth:`make_key`."""
        key = self._key(query, limit)
        now = time.monotonic() if now is None else now
        with self._lock:
            entry = self._data.get(key)
            if entry is None:
                self._misses += 1
                return None
            if now - entry.timestamp > self._ttl:
                del self._data[key]
                self._misses += 1
                self._evictions += 1
                return None
            self._data.move_to_end(key)
            entry.hit_count += 1
            self._hits += 1
            return entry.results

    def put(self, query: str, limit: Any, results: Any = _UNSET, *, now: float | None = None) -> None:
        """``put(query, limit, results)`` as in the reference, or ``put(key, value)`` with a ready-made key."""
        if results is _UNSET:
            key, value = query, limit
        else:
            key, value = self.make_key(query, int(limit)), results
        now = time.monotonic() if now is None else now
        with self._lock:
            self._data[key] = CacheEntry(value, now)
            self._data.move_to_end(key)
            while len(self._data) > self._max:
                self._data.popitem(last=False)
                self._evictions += 1

    def invalidate(self, query: str | None = None, limit: int | None = None) -> bool:
        """Drop one entry (True when it existed); with no arguments, drop everything."""
        with self._lock:
            if query is None:
                had = bool(self._data)
                self._data.clear()
                return had
            return self._data.pop(self._key(query, limit), None) is not None

    def clear(self) -> None:
        with self._lock:
            self._data.clear()

    def evict_expired(self, *, now: float | None = None) -> int:
        now = time.monotonic() if now is None else now
        with self._lock:
            dead = [k for k, e in self._data.items() if now - e.timestamp > self._ttl]
            for k in dead:
                del self._data[k]
            self._evictions += len(dead)
        return len(dead)

    @property
    def size(self) -> int:
        with self._lock:
            return len(self._data)

    def __len__(self) -> int:
        return self.size

    @property
    def stats(self) -> CacheStats:
        with self._lock:
            return CacheStats(self._hits, self._misses, self._evictions, len(self._data), self._max)
