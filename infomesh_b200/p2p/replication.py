"""N = 3 document replication: replicas go to the peers whose sha256(peer_id) is XOR-closest to
sha256(url DHT key); each REPLICATE_REQUEST carries the full text and is acknowledged with REPLICATE_RESPONSE;
10 s per replica (reference infomesh/p2p/replication.py:37-306)."""
from __future__ import annotations

import asyncio
import hashlib
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Any, Awaitable, Callable

from infomesh_b200.p2p.protocol import MessageType, ReplicateRequest, dataclass_to_payload, url_to_dht_key
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_REPLICATION_FACTOR = 3
REPLICATE_TIMEOUT_SECONDS = 10


@dataclass
class ReplicationStats:
    documents_replicated: int = 0
    replicas_sent: int = 0
    replicas_received: int = 0
    replicas_failed: int = 0
    avg_replicate_ms: float = 0.0
    _times: deque = field(default_factory=lambda: deque(maxlen=10_000), repr=False)

    def record_time(self, ms: float) -> None:
        self._times.append(ms)
        self.avg_replicate_ms = sum(self._times) / len(self._times)


def _h(s: str) -> int:
    return int.from_bytes(hashlib.sha256(s.encode()).digest(), "big")


def replica_peers(url: str, candidates: list[str], n: int = DEFAULT_REPLICATION_FACTOR) -> list[str]:
    target = _h(url_to_dht_key(url))
    return sorted(candidates, key=lambda pid: target ^ _h(pid))[:n]


class Replicator:
    def __init__(self, send, connected_peers: Callable[[], list[str]], local_peer_id: str, *,
                 replication_factor: int = DEFAULT_REPLICATION_FACTOR):
        self._send, self._connected, self._peer_id = send, connected_peers, local_peer_id
        self._n = replication_factor
        self._stats = ReplicationStats()

    @property
    def stats(self) -> ReplicationStats:
        return self._stats

    async def replicate_document(self, doc_id: int, url: str, title: str, text: str, text_hash: str, language: str = "") -> int:
        targets = replica_peers(url, [p for p in self._connected() if p != self._peer_id], self._n)
        if not targets:
            return 0

        async def one(pid: str, idx: int) -> bool:
            req = ReplicateRequest(doc_id=doc_id, url=url, title=title, text=text, text_hash=text_hash, language=language,
                                   source_peer_id=self._peer_id, replica_index=idx)
            t0 = time.monotonic()
            try:
                reply = await asyncio.wait_for(self._send(pid, MessageType.REPLICATE_REQUEST, dataclass_to_payload(req),
                                                          REPLICATE_TIMEOUT_SECONDS), timeout=REPLICATE_TIMEOUT_SECONDS)
                ok = bool(reply and reply[0] == MessageType.REPLICATE_RESPONSE and reply[1].get("success", True))
            except Exception:  # noqa: BLE001
                ok = False
            if ok:
                self._stats.replicas_sent += 1
                self._stats.record_time((time.monotonic() - t0) * 1000)
            else:
                self._stats.replicas_failed += 1
            return ok

        done = sum(await asyncio.gather(*(one(p, i) for i, p in enumerate(targets))))
        if done:
            self._stats.documents_replicated += 1
        return done

    async def handle_replicate_request(self, payload: dict[str, Any],
                                       store_fn: Callable[..., Awaitable[bool]]) -> tuple[MessageType, dict[str, Any]]:
        """``store_fn(url=, title=, text=, text_hash=, language=) -> bool`` persists the replica."""
        url, text = str(payload.get("url", "")), str(payload.get("text", ""))
        claimed = str(payload.get("text_hash", ""))
        if not url or not text:
            return MessageType.REPLICATE_RESPONSE, {"success": False, "error": "empty", "peer_id": self._peer_id}
        if claimed and hashlib.sha256(text.encode()).hexdigest() != claimed:
            return MessageType.REPLICATE_RESPONSE, {"success": False, "error": "hash_mismatch", "peer_id": self._peer_id}
        try:
            ok = bool(await store_fn(url=url, title=str(payload.get("title", "")), text=text, text_hash=claimed,
                                     language=str(payload.get("language", ""))))
        except Exception as exc:  # noqa: BLE001
            logger.warning("replica_store_failed", url=url, error=str(exc))
            ok = False
        if ok:
            self._stats.replicas_received += 1
        return MessageType.REPLICATE_RESPONSE, {"success": ok, "url": url, "peer_id": self._peer_id}
