"""InfoMesh DHT facade: keyword -> pointer lists (read-merge-write, <= 100 pointers, 10 publishes / keyword / hour),
crawl locks with a 300 s TTL, attestations, raw put/get.  Works over any backend exposing async
``put_value(key, bytes)`` / ``get_value(key) -> bytes | None`` (the built-in ``p2p.kademlia.KadDHT`` or py-libp2p's).
Key namespaces and value encodings follow reference infomesh/p2p/dht.py:31-400."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any

import msgpack

from infomesh_b200.hashing import content_hash
from infomesh_b200.p2p.protocol import keyword_to_dht_key, safe_unpackb
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_PREFIX_CRAWL_LOCK = "/infomesh/lock/"
_PREFIX_ATTESTATION = "/infomesh/att/"
_LOCK_TTL_SECONDS = 300
MAX_POINTERS_PER_KEYWORD = 100
MAX_PUBLISHES_PER_KEYWORD_HR = 10


@dataclass
class DHTStats:
    keys_stored: int = 0
    keys_published: int = 0
    gets_performed: int = 0
    puts_performed: int = 0
    locks_acquired: int = 0
    locks_released: int = 0


def _merge_pointers(existing: list[dict[str, Any]], new: list[dict[str, Any]], *, limit: int) -> list[dict[str, Any]]:
    """New pointers first, one entry per (peer_id, doc_id), malformed entries dropped."""
    merged: dict[tuple[str, int], dict[str, Any]] = {}
    for p in [*new, *existing]:
        if not isinstance(p, dict):
            continue
        pid, did = p.get("peer_id", ""), p.get("doc_id", 0)
        if not isinstance(pid, str) or not isinstance(did, int) or isinstance(did, bool):
            continue
        merged.setdefault((pid, did), p)
        if len(merged) >= limit:
            break
    return list(merged.values())


class InfoMeshDHT:
    def __init__(self, kad_dht: Any, local_peer_id: str):
        self._dht = kad_dht
        self._peer_id = local_peer_id
        self._stats = DHTStats()
        self._publishes: dict[str, list[float]] = {}

    @property
    def stats(self) -> DHTStats:
        return self._stats

    # ------------------------------------------------------------------ raw
    # The counters mirror what the reference counts: puts / gets of the public operations (raw put / get, keyword publish and
    # query, attestations).  Internal reads (the merge before a publish, lock checks) and lock writes are not counted.
    async def _store(self, key: str, value: bytes, *, what: str) -> bool:
        try:
            await self._dht.put_value(key, value)
            return True
        except Exception:  # noqa: BLE001
            logger.exception(what, key=key)
            return False

    async def _fetch(self, key: str, *, what: str) -> tuple[bool, bytes | None]:
        try:
            return True, await self._dht.get_value(key)
        except Exception:  # noqa: BLE001
            logger.exception(what, key=key)
            return False, None

    async def put(self, key: str, value: bytes) -> bool:
        ok = await self._store(key, value, what="dht_put_failed")
        self._stats.puts_performed += ok
        return ok

    async def get(self, key: str) -> bytes | None:
        ok, raw = await self._fetch(key, what="dht_get_failed")
        self._stats.gets_performed += ok
        return raw

    async def _get_map(self, key: str, *, count: bool = False) -> dict[str, Any] | None:
        ok, raw = await self._fetch(key, what="dht_get_failed")
        if count:
            self._stats.gets_performed += ok
        if raw is None:
            return None
        try:
            obj = safe_unpackb(raw)
        except Exception:  # noqa: BLE001
            return None
        return obj if isinstance(obj, dict) else None

    # ------------------------------------------------------------------ keyword index
    def _publish_allowed(self, keyword: str) -> bool:
        now = time.time()
        recent = [t for t in self._publishes.get(keyword, []) if now - t < 3600]
        if recent:
            self._publishes[keyword] = recent
        else:
            self._publishes.pop(keyword, None)
        return len(recent) < MAX_PUBLISHES_PER_KEYWORD_HR

    async def publish_keyword(self, keyword: str, pointers: list[dict[str, Any]], *, signature: bytes = b"") -> bool:
        if not self._publish_allowed(keyword):
            logger.warning("dht_publish_rate_limited", keyword=keyword)
            return False
        merged = _merge_pointers(await self._pointers(keyword, count=False), pointers, limit=MAX_POINTERS_PER_KEYWORD)
        value = msgpack.packb({"keyword": keyword, "pointers": merged, "peer_id": self._peer_id,
                               "timestamp": time.time(), "signature": signature}, use_bin_type=True)
        if not await self._store(keyword_to_dht_key(keyword), value, what="dht_publish_failed"):
            return False
        self._stats.puts_performed += 1
        self._stats.keys_published += 1
        self._publishes.setdefault(keyword, []).append(time.time())
        return True

    async def _pointers(self, keyword: str, *, count: bool) -> list[dict[str, Any]]:
        entry = await self._get_map(keyword_to_dht_key(keyword), count=count)
        ptrs = entry.get("pointers", []) if entry else []
        return [p for p in ptrs if isinstance(p, dict)] if isinstance(ptrs, list) else []

    async def query_keyword(self, keyword: str) -> list[dict[str, Any]]:
        return await self._pointers(keyword, count=True)

    # ------------------------------------------------------------------ crawl locks
    @staticmethod
    def _lock_key(url: str) -> str:
        return f"{_PREFIX_CRAWL_LOCK}{content_hash(url)}"

    async def acquire_crawl_lock(self, url: str, ttl_seconds: int = _LOCK_TTL_SECONDS) -> bool:
        """Exclusive while the TTL runs -- also against this node's own second attempt: two local workers that pick the same
        URL must not both crawl it.  An expired (or released: timestamp 0) lock is simply overwritten."""
        held = await self._get_map(self._lock_key(url))
        if held is not None:
            ts = held.get("timestamp", 0)
            if isinstance(ts, (int, float)) and time.time() - ts < ttl_seconds:
                logger.debug("crawl_lock_held", url=url, holder=held.get("peer_id"))
                return False
        ok = await self._store(self._lock_key(url), msgpack.packb(
            {"peer_id": self._peer_id, "url": url, "timestamp": time.time(), "ttl": ttl_seconds}, use_bin_type=True), what="crawl_lock_acquire_failed")
        self._stats.locks_acquired += ok
        return ok

    async def release_crawl_lock(self, url: str) -> bool:
        ok = await self._store(self._lock_key(url), msgpack.packb(
            {"peer_id": self._peer_id, "url": url, "timestamp": 0, "ttl": 0}, use_bin_type=True), what="crawl_lock_release_failed")
        self._stats.locks_released += ok
        return ok

    # ------------------------------------------------------------------ attestations
    async def publish_attestation(self, url: str, raw_hash: str, text_hash: str, signature: bytes = b"") -> bool:
        return await self.put(f"{_PREFIX_ATTESTATION}{content_hash(url)}", msgpack.packb(  # counted as a put, like the reference
            {"url": url, "raw_hash": raw_hash, "text_hash": text_hash, "peer_id": self._peer_id,
             "timestamp": time.time(), "signature": signature}, use_bin_type=True))

    async def get_attestation(self, url: str) -> dict[str, Any] | None:
        return await self._get_map(f"{_PREFIX_ATTESTATION}{content_hash(url)}", count=True)
