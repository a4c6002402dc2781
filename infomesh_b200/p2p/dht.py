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
    async def put(self, key: str, value: bytes) -> bool:
        try:
            await self._dht.put_value(key, value)
            self._stats.puts_performed += 1
            return True
        except Exception:  # noqa: BLE001
            logger.exception("dht_put_failed", key=key)
            return False

    async def get(self, key: str) -> bytes | None:
        try:
            raw = await self._dht.get_value(key)
            self._stats.gets_performed += 1
            return raw
        except Exception:  # noqa: BLE001
            logger.exception("dht_get_failed", key=key)
            return None

    async def _get_map(self, key: str) -> dict[str, Any] | None:
        raw = await self.get(key)
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
        merged = _merge_pointers(await self.query_keyword(keyword), pointers, limit=MAX_POINTERS_PER_KEYWORD)
        value = msgpack.packb({"keyword": keyword, "pointers": merged, "peer_id": self._peer_id,
                               "timestamp": time.time(), "signature": signature}, use_bin_type=True)
        if not await self.put(keyword_to_dht_key(keyword), value):
            return False
        self._stats.keys_published += 1
        self._publishes.setdefault(keyword, []).append(time.time())
        return True

    async def query_keyword(self, keyword: str) -> list[dict[str, Any]]:
        entry = await self._get_map(keyword_to_dht_key(keyword))
        ptrs = entry.get("pointers", []) if entry else []
        return [p for p in ptrs if isinstance(p, dict)] if isinstance(ptrs, list) else []

    # ------------------------------------------------------------------ crawl locks
    @staticmethod
    def _lock_key(url: str) -> str:
        return f"{_PREFIX_CRAWL_LOCK}{content_hash(url)}"

    async def acquire_crawl_lock(self, url: str, ttl_seconds: int = _LOCK_TTL_SECONDS) -> bool:
        held = await self._get_map(self._lock_key(url))
        if held is not None:
            ts = held.get("timestamp", 0)
            if isinstance(ts, (int, float)) and time.time() - ts < ttl_seconds and held.get("peer_id") != self._peer_id:
                return False
        ok = await self.put(self._lock_key(url), msgpack.packb(
            {"peer_id": self._peer_id, "url": url, "timestamp": time.time(), "ttl": ttl_seconds}, use_bin_type=True))
        self._stats.locks_acquired += 1 if ok else 0
        return ok

    async def release_crawl_lock(self, url: str) -> bool:
        ok = await self.put(self._lock_key(url), msgpack.packb(
            {"peer_id": self._peer_id, "url": url, "timestamp": 0, "ttl": 0}, use_bin_type=True))
        self._stats.locks_released += 1 if ok else 0
        return ok

    # ------------------------------------------------------------------ attestations
    async def publish_attestation(self, url: str, raw_hash: str, text_hash: str, signature: bytes = b"") -> bool:
        return await self.put(f"{_PREFIX_ATTESTATION}{content_hash(url)}", msgpack.packb(
            {"url": url, "raw_hash": raw_hash, "text_hash": text_hash, "peer_id": self._peer_id,
             "timestamp": time.time(), "signature": signature}, use_bin_type=True))

    async def get_attestation(self, url: str) -> dict[str, Any] | None:
        return await self._get_map(f"{_PREFIX_ATTESTATION}{content_hash(url)}")
