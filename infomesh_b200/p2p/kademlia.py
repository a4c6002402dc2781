"""A compact Kademlia (160-bit ids, k = 20 buckets, alpha = 3 iterative lookups, STORE / FIND_NODE / FIND_VALUE)
over :class:`infomesh_b200.p2p.transport.Transport`.

The reference delegates the DHT to py-libp2p's KadDHT (infomesh/p2p/node.py:641-655, infomesh/p2p/dht.py:80-140);
that library is not available here, so the overlay is implemented directly.  ``KadDHT`` exposes the
``put_value`` / ``get_value`` pair the :class:`InfoMeshDHT` facade expects, applies the per-bucket subnet quota from
``p2p.sybil`` when inserting contacts, and expires stored values after 24 h.
"""
from __future__ import annotations

import asyncio
import hashlib
import time
from dataclasses import dataclass, field
from typing import Any

from infomesh_b200.p2p.protocol import MessageType
from infomesh_b200.p2p.sybil import SubnetLimiter
from infomesh_b200.p2p.transport import PeerInfo, Transport
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

K_BUCKET = 20
ALPHA = 3
ID_BITS = 160
VALUE_TTL = 24 * 3600
MAX_VALUE_BYTES = 1 << 20
MAX_STORED_VALUES = 50_000


def key_id(key: str | bytes) -> int:
    raw = key.encode() if isinstance(key, str) else key
    return int.from_bytes(hashlib.sha256(raw).digest()[:20], "big")


def node_int(peer_id: str) -> int:
    try:
        if len(peer_id) == 40:
            return int(peer_id, 16)
    except ValueError:
        pass
    return key_id(peer_id)


@dataclass
class Contact:
    peer_id: str
    host: str
    port: int
    last_seen: float = field(default_factory=time.time)

    def to_wire(self) -> dict[str, Any]:
        return {"peer_id": self.peer_id, "host": self.host, "port": self.port}


class RoutingTable:
    def __init__(self, self_id: str, *, k: int = K_BUCKET, subnet_limiter: SubnetLimiter | None = None):
        self.self_id, self._self_int, self.k = self_id, node_int(self_id), k
        self.buckets: list[dict[str, Contact]] = [dict() for _ in range(ID_BITS)]
        self.limiter = subnet_limiter

    def bucket_index(self, peer_id: str) -> int:
        d = self._self_int ^ node_int(peer_id)
        return max(d.bit_length() - 1, 0)

    def add(self, c: Contact) -> bool:
        if c.peer_id == self.self_id or not c.peer_id:
            return False
        i = self.bucket_index(c.peer_id)
        b = self.buckets[i]
        if c.peer_id in b:
            b.pop(c.peer_id)
            b[c.peer_id] = c               # move to tail = most recently seen
            return True
        if len(b) >= self.k:
            return False                   # (classic Kademlia pings the head; stale heads are evicted by remove())
        if self.limiter is not None:
            try:
                if not self.limiter.add(c.host, c.peer_id, i):
                    return False
            except ValueError:
                pass                        # host is a DNS name: quota not applicable
        b[c.peer_id] = c
        return True

    def remove(self, peer_id: str) -> None:
        i = self.bucket_index(peer_id)
        c = self.buckets[i].pop(peer_id, None)
        if c is not None and self.limiter is not None:
            try:
                self.limiter.remove(c.host, peer_id, i)
            except ValueError:
                pass

    def get(self, peer_id: str) -> Contact | None:
        return self.buckets[self.bucket_index(peer_id)].get(peer_id)

    def all(self) -> list[Contact]:
        return [c for b in self.buckets for c in b.values()]

    def closest(self, target: int, n: int = K_BUCKET) -> list[Contact]:
        return sorted(self.all(), key=lambda c: node_int(c.peer_id) ^ target)[:n]

    def __len__(self) -> int:
        return sum(len(b) for b in self.buckets)


class KadDHT:
    def __init__(self, transport: Transport, *, subnet_limiter: SubnetLimiter | None = None, rpc_timeout: float = 3.0):
        self.t = transport
        self.table = RoutingTable(transport.peer_id, subnet_limiter=subnet_limiter)
        self._store: dict[str, tuple[bytes, float]] = {}
        self._timeout = rpc_timeout
        for kind, fn in ((MessageType.PING, self._on_ping), (MessageType.DHT_FIND_NODE, self._on_find_node),
                         (MessageType.DHT_FIND_VALUE, self._on_find_value), (MessageType.DHT_STORE, self._on_store)):
            transport.register(kind, fn)

    # ------------------------------------------------------------------ inbound RPCs
    def _me(self) -> dict[str, Any]:
        return {"peer_id": self.t.peer_id, "host": self.t.host, "port": self.t.port}

    def _note(self, payload: dict[str, Any], peer: PeerInfo) -> None:
        frm = payload.get("from")
        if isinstance(frm, dict) and frm.get("peer_id") and isinstance(frm.get("port"), int):
            pid = str(frm["peer_id"])
            if peer.peer_id and peer.peer_id != pid:
                return                      # signed sender does not match the claimed contact
            self.table.add(Contact(pid, peer.host, int(frm["port"])))

    async def _on_ping(self, payload, peer):
        self._note(payload, peer)
        return MessageType.PONG, {"from": self._me(), "ts": time.time()}

    async def _on_find_node(self, payload, peer):
        self._note(payload, peer)
        try:
            target = int(str(payload.get("target", "0")), 16)     # 160-bit ids do not fit msgpack integers
        except ValueError:
            target = 0
        return MessageType.DHT_NODES, {"from": self._me(), "nodes": [c.to_wire() for c in self.table.closest(target)]}

    async def _on_find_value(self, payload, peer):
        self._note(payload, peer)
        key = str(payload.get("key", ""))
        hit = self._local_get(key)
        if hit is not None:
            return MessageType.DHT_VALUE, {"from": self._me(), "key": key, "value": hit}
        return MessageType.DHT_NODES, {"from": self._me(),
                                       "nodes": [c.to_wire() for c in self.table.closest(key_id(key))]}

    async def _on_store(self, payload, peer):
        self._note(payload, peer)
        key, value = str(payload.get("key", "")), payload.get("value")
        ok = isinstance(value, bytes | bytearray) and len(value) <= MAX_VALUE_BYTES and bool(key)
        if ok:
            self._local_put(key, bytes(value))
        return MessageType.DHT_STORE_ACK, {"from": self._me(), "ok": ok}

    # ------------------------------------------------------------------ local store
    def _local_put(self, key: str, value: bytes) -> None:
        if len(self._store) >= MAX_STORED_VALUES:
            now = time.time()
            for k in [k for k, (_, exp) in self._store.items() if exp < now]:
                del self._store[k]
            if len(self._store) >= MAX_STORED_VALUES:
                self._store.pop(next(iter(self._store)))
        self._store[key] = (value, time.time() + VALUE_TTL)

    def _local_get(self, key: str) -> bytes | None:
        hit = self._store.get(key)
        if hit is None:
            return None
        if hit[1] < time.time():
            del self._store[key]
            return None
        return hit[0]

    # ------------------------------------------------------------------ outbound
    async def _rpc(self, c: Contact, kind: MessageType, payload: dict[str, Any]):
        try:
            reply = await self.t.request((c.host, c.port), kind, {**payload, "from": self._me()}, timeout=self._timeout)
        except Exception:  # noqa: BLE001 — any failure marks the contact dead
            self.table.remove(c.peer_id)
            return None
        c.last_seen = time.time()
        self.table.add(c)
        return reply

    async def ping(self, host: str, port: int) -> Contact | None:
        try:
            reply = await self.t.request((host, port), MessageType.PING, {"from": self._me()}, timeout=self._timeout)
        except Exception:  # noqa: BLE001
            return None
        if not reply or reply[0] != MessageType.PONG:
            return None
        frm = reply[1].get("from") or {}
        if not frm.get("peer_id"):
            return None
        c = Contact(str(frm["peer_id"]), host, int(frm.get("port", port)))
        self.table.add(c)
        return c

    async def bootstrap(self, addrs: list[tuple[str, int]]) -> int:
        found = [c for c in await asyncio.gather(*(self.ping(h, p) for h, p in addrs)) if c]
        if found:
            await self.lookup_nodes(node_int(self.t.peer_id))
        return len(found)

    async def _iterate(self, target: int, *, key: str | None = None) -> tuple[list[Contact], bytes | None]:
        shortlist = {c.peer_id: c for c in self.table.closest(target, K_BUCKET)}
        asked: set[str] = set()
        while True:
            todo = [c for c in sorted(shortlist.values(), key=lambda c: node_int(c.peer_id) ^ target)[:K_BUCKET]
                    if c.peer_id not in asked][:ALPHA]
            if not todo:
                break
            asked.update(c.peer_id for c in todo)
            if key is None:
                replies = await asyncio.gather(*(self._rpc(c, MessageType.DHT_FIND_NODE, {"target": f"{target:040x}"}) for c in todo))
            else:
                replies = await asyncio.gather(*(self._rpc(c, MessageType.DHT_FIND_VALUE, {"key": key}) for c in todo))
            for c, rep in zip(todo, replies):
                if rep is None:
                    shortlist.pop(c.peer_id, None)
                    continue
                kind, body = rep
                if kind == MessageType.DHT_VALUE and isinstance(body.get("value"), bytes | bytearray):
                    return list(shortlist.values()), bytes(body["value"])
                for n in body.get("nodes", []) or []:
                    try:
                        nc = Contact(str(n["peer_id"]), str(n["host"]), int(n["port"]))
                    except (KeyError, TypeError, ValueError):
                        continue
                    if nc.peer_id != self.t.peer_id and nc.peer_id not in shortlist:
                        shortlist[nc.peer_id] = nc
        live = sorted(shortlist.values(), key=lambda c: node_int(c.peer_id) ^ target)[:K_BUCKET]
        return live, None

    async def lookup_nodes(self, target: int) -> list[Contact]:
        return (await self._iterate(target))[0]

    async def put_value(self, key: str, value: bytes) -> int:
        """Store on the k closest nodes (and locally).  Returns the number of remote replicas written."""
        if len(value) > MAX_VALUE_BYTES:
            raise ValueError("DHT value too large")
        self._local_put(key, value)
        nodes = await self.lookup_nodes(key_id(key))
        acks = await asyncio.gather(*(self._rpc(c, MessageType.DHT_STORE, {"key": key, "value": value}) for c in nodes))
        return sum(1 for a in acks if a and a[1].get("ok"))

    async def get_value(self, key: str) -> bytes | None:
        hit = self._local_get(key)
        if hit is not None:
            return hit
        return (await self._iterate(key_id(key), key=key))[1]

    def connected_contacts(self) -> list[Contact]:
        return self.table.all()
