"""The peer node: identity (Ed25519 + cached proof-of-work), transport, Kademlia, InfoMesh DHT facade, distributed
index, query router, replicator, PEX, mDNS, peer store, credit-sync rounds, status file
(reference infomesh/p2p/node.py:65-1662).

The reference runs py-libp2p inside a trio loop on a background thread and bridges asyncio callers through
``trio.from_thread``.  Here the whole stack is asyncio: the node owns one event loop on a daemon thread and public
coroutines hop onto it with ``run_coroutine_threadsafe`` — callers may live in any loop (or none, via the ``*_sync``
helpers).
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import struct
import threading
import time
from dataclasses import dataclass, field
from enum import StrEnum
from pathlib import Path
from typing import Any, Awaitable, Callable

from infomesh_b200 import __version__
from infomesh_b200.p2p import bootstrap as BS
from infomesh_b200.p2p.dht import InfoMeshDHT
from infomesh_b200.p2p.kademlia import KadDHT
from infomesh_b200.p2p.keys import ensure_keys
from infomesh_b200.p2p.mdns import MDNSDiscovery
from infomesh_b200.p2p.peer_store import PeerStore
from infomesh_b200.p2p.pex import PEX_MAX_PEERS, PEX_MAX_PEERS_PER_ROUND, PEX_ROUND_INTERVAL, PeerExchange
from infomesh_b200.p2p.protocol import ALL_PROTOCOLS, MessageType
from infomesh_b200.p2p.replication import Replicator
from infomesh_b200.p2p.routing import QueryRouter
from infomesh_b200.p2p.sybil import (DEFAULT_DIFFICULTY_BITS, SubnetLimiter, compute_pow_hash, generate_pow,
                                     leading_zero_bits)
from infomesh_b200.p2p.throttle import BandwidthThrottle
from infomesh_b200.p2p.transport import PeerInfo, Transport, format_multiaddr, parse_multiaddr
from infomesh_b200.utils.log import get_logger
from infomesh_b200.version_check import PeerVersionTracker

logger = get_logger(__name__)

_ROUTING_REFRESH_INTERVAL = 300
_STATUS_WRITE_INTERVAL = 10
_CREDIT_SYNC_INTERVAL = 300
_VERSION_CHECK_INTERVAL = 3600
_SEARCH_NETWORK_TIMEOUT = 30
_PUBLISH_TIMEOUT = 60


class NodeState(StrEnum):
    STOPPED = "stopped"
    STARTING = "starting"
    RUNNING = "running"
    STOPPING = "stopping"
    ERROR = "error"


@dataclass
class NodeInfo:
    peer_id: str = ""
    listen_addrs: list[str] = field(default_factory=list)
    connected_peers: int = 0
    state: str = NodeState.STOPPED
    uptime_seconds: float = 0.0
    dht_keys_stored: int = 0


def load_cached_pow(path: Path, pub: bytes) -> int | None:
    """Cache layout: sha256(pubkey) (32) + nonce LE64 (8) + difficulty (1); the legacy 40-byte form means 20 bits."""
    try:
        data = Path(path).read_bytes()
    except OSError:
        return None
    if len(data) not in (40, 41) or data[:32] != hashlib.sha256(pub).digest():
        return None
    nonce = struct.unpack("<Q", data[32:40])[0]
    bits = data[40] if len(data) == 41 else 20
    return nonce if leading_zero_bits(compute_pow_hash(pub, nonce)) >= bits else None


def save_cached_pow(path: Path, pub: bytes, nonce: int, difficulty: int = DEFAULT_DIFFICULTY_BITS) -> None:
    try:
        Path(path).write_bytes(hashlib.sha256(pub).digest() + struct.pack("<Q", nonce) + bytes([difficulty]))
    except OSError:
        logger.debug("pow_cache_write_failed")


class InfoMeshNode:
    def __init__(self, config, *, local_search_fn: Callable[[str, int], Awaitable[list[dict]]] | None = None,
                 store_fn: Callable[..., Awaitable[bool]] | None = None, index_submit_receiver: Any | None = None,
                 credit_sync_manager: Any | None = None, llm_handler: Any | None = None,
                 is_isolated_fn: Callable[[str], bool] | None = None, pow_difficulty: int = DEFAULT_DIFFICULTY_BITS,
                 enable_mdns: bool = True):
        self._config = config
        self._local_search_fn, self._store_fn = local_search_fn, store_fn
        self._index_submit_receiver, self._credit_sync_manager = index_submit_receiver, credit_sync_manager
        self._llm_handler = llm_handler
        self._is_isolated = is_isolated_fn
        self._pow_difficulty, self._enable_mdns = pow_difficulty, enable_mdns
        self._state = NodeState.STOPPED
        self._peer_id = ""
        self._start_time = 0.0
        self._key_pair = None
        self._transport: Transport | None = None
        self._kad: KadDHT | None = None
        self._dht: InfoMeshDHT | None = None
        self._distributed_index = None
        self._router: QueryRouter | None = None
        self._replicator: Replicator | None = None
        self._subnet_limiter = SubnetLimiter(max_per_subnet=getattr(config.network, "subnet_max_per_bucket", 3))
        self._mdns: MDNSDiscovery | None = None
        self._peer_store: PeerStore | None = None
        self._pex: PeerExchange | None = None
        self._throttle = BandwidthThrottle(config.network.upload_limit_mbps, config.network.download_limit_mbps)
        self._pow_nonce: int | None = None
        self._url_assigner = None
        self._bootstrap_results: dict[str, object] = {}
        self._peer_version_tracker = PeerVersionTracker()
        self._loop: asyncio.AbstractEventLoop | None = None
        self._thread: threading.Thread | None = None
        self._started = threading.Event()
        self._stop_async: asyncio.Event | None = None
        self._error = ""

    # ------------------------------------------------------------------ accessors
    state = property(lambda self: self._state)
    version_tracker = property(lambda self: self._peer_version_tracker)
    peer_id = property(lambda self: self._peer_id)
    dht = property(lambda self: self._dht)
    router = property(lambda self: self._router)
    distributed_index = property(lambda self: self._distributed_index)
    replicator = property(lambda self: self._replicator)
    throttle = property(lambda self: self._throttle)
    subnet_limiter = property(lambda self: self._subnet_limiter)
    mdns = property(lambda self: self._mdns)
    pow_nonce = property(lambda self: self._pow_nonce)
    url_assigner = property(lambda self: self._url_assigner)
    transport = property(lambda self: self._transport)
    key_pair = property(lambda self: self._key_pair)

    @property
    def listen_addrs(self) -> list[str]:
        t = self._transport
        return [format_multiaddr(t.host, t.port, self._peer_id)] if t is not None and t.port else []

    def get_connected_peers(self) -> list[str]:
        return [c.peer_id for c in self._kad.connected_contacts()] if self._kad else []

    connected_peers = property(get_connected_peers)

    def check_subnet(self, ip: str, peer_id: str, bucket_id: int = 0) -> bool:
        return self._subnet_limiter.add(ip, peer_id, bucket_id)

    def get_info(self) -> NodeInfo:
        return NodeInfo(self._peer_id, self.listen_addrs, len(self.get_connected_peers()), str(self._state),
                        time.time() - self._start_time if self._start_time else 0.0,
                        self._dht.stats.keys_stored if self._dht else 0)

    # ------------------------------------------------------------------ lifecycle
    def start(self, *, blocking: bool = False, timeout: float = 60.0) -> None:
        if self._state in (NodeState.STARTING, NodeState.RUNNING):
            return
        self._state = NodeState.STARTING
        self._started.clear()
        if blocking:
            self._thread_main()
            return
        self._thread = threading.Thread(target=self._thread_main, daemon=True, name="infomesh-p2p")
        self._thread.start()
        if not self._started.wait(timeout):
            self._state, self._error = NodeState.ERROR, "startup timed out"
            raise RuntimeError("P2P node failed to start within timeout")
        if self._state == NodeState.ERROR:
            raise RuntimeError(f"P2P node failed to start: {self._error}")

    def stop(self, timeout: float = 10.0) -> None:
        if self._state in (NodeState.STOPPED, NodeState.STOPPING):
            return
        self._state = NodeState.STOPPING
        if self._loop is not None and self._stop_async is not None:
            self._loop.call_soon_threadsafe(self._stop_async.set)
        if self._thread is not None and self._thread is not threading.current_thread():
            self._thread.join(timeout)
        self._thread = None
        self._state = NodeState.STOPPED

    def _thread_main(self) -> None:
        loop = asyncio.new_event_loop()
        self._loop = loop
        try:
            loop.run_until_complete(self._main())
        except Exception as exc:  # noqa: BLE001
            self._state, self._error = NodeState.ERROR, str(exc)
            logger.exception("p2p_node_crashed")
            self._write_status_file(state="error", error=str(exc))
            self._started.set()
        finally:
            try:
                loop.run_until_complete(loop.shutdown_asyncgens())
            finally:
                loop.close()
                self._loop = None

    def _prepare_identity(self) -> None:
        data_dir = Path(self._config.node.data_dir)
        self._key_pair = ensure_keys(data_dir)
        self._peer_id = self._key_pair.peer_id
        pub = self._key_pair.public_key_bytes()
        cache = data_dir / "keys" / "pow.bin"
        nonce = load_cached_pow(cache, pub)
        if nonce is None and self._pow_difficulty > 0:
            nonce = generate_pow(pub, self._pow_difficulty).nonce
            save_cached_pow(cache, pub, nonce, self._pow_difficulty)
        self._pow_nonce = nonce

    async def _main(self) -> None:
        cfg = self._config
        self._stop_async = asyncio.Event()
        self._prepare_identity()
        self._transport = Transport(self._key_pair, throttle=self._throttle, is_isolated_fn=self._is_isolated,
                                    encrypt=getattr(cfg.network, "encrypt", True),
                                    require_encrypted=getattr(cfg.network, "require_encrypted", False))
        if self._index_submit_receiver is not None and hasattr(self._index_submit_receiver, "bind_key_registry"):
            self._index_submit_receiver.bind_key_registry(self._transport.keys)
        await self._transport.listen(cfg.node.listen_address, cfg.node.listen_port)
        self._kad = KadDHT(self._transport, subnet_limiter=self._subnet_limiter)
        self._dht = InfoMeshDHT(self._kad, self._peer_id)
        from infomesh_b200.index.distributed import DistributedIndex

        self._distributed_index = DistributedIndex(self._dht, self._peer_id)
        self._router = QueryRouter(self._send_to_peer, self._dht, self._peer_id, connected_peers=self.get_connected_peers)
        self._replicator = Replicator(self._send_to_peer, self.get_connected_peers, self._peer_id,
                                      replication_factor=cfg.network.replication_factor)
        self._pex = PeerExchange(self._peer_id)
        try:
            self._peer_store = PeerStore(cfg.node.data_dir)
        except Exception:  # noqa: BLE001
            self._peer_store = None
        try:
            from infomesh_b200.crawler.url_assigner import UrlAssigner

            self._url_assigner = UrlAssigner(self._peer_id)
        except Exception:  # noqa: BLE001
            self._url_assigner = None
        self._register_handlers()
        await self._bootstrap()
        if self._enable_mdns:
            self._mdns = MDNSDiscovery(self._peer_id, self._transport.port)
            if not self._mdns.start():
                self._mdns = None
        self._start_time = time.time()
        self._state = NodeState.RUNNING
        self._write_status_file()
        self._started.set()
        await self._announce_credit_sync()
        try:
            await self._run_main_loop()
        finally:
            await self._save_connected_peers()
            if self._peer_store is not None:
                self._peer_store.close()
                self._peer_store = None
            if self._mdns is not None:
                self._mdns.stop()
                self._mdns = None
            await self._transport.close()
            self._write_status_file(state="stopped")

    # ------------------------------------------------------------------ address book + send
    def _addr_of(self, peer_id: str) -> tuple[str, int] | None:
        c = self._kad.table.get(peer_id) if self._kad else None
        return (c.host, c.port) if c else None

    async def _send_to_peer(self, peer_id: str, kind: MessageType, payload: dict[str, Any], timeout: float):
        addr = self._addr_of(peer_id)
        if addr is None:
            raise ConnectionError(f"no address for peer {peer_id[:16]}")
        return await self._transport.request(addr, kind, payload, timeout=timeout)

    # ------------------------------------------------------------------ handlers
    def _register_handlers(self) -> None:
        t = self._transport

        async def on_ping(payload, peer: PeerInfo):
            ver = payload.get("version")
            if peer.peer_id and isinstance(ver, str):
                self._peer_version_tracker.record(peer.peer_id, ver)
            reply = await self._kad._on_ping(payload, peer)
            reply[1].update(version=__version__, protocols=list(ALL_PROTOCOLS))
            return reply

        async def on_search(payload, peer: PeerInfo):
            if self._local_search_fn is None:
                return MessageType.SEARCH_RESPONSE, {"request_id": payload.get("request_id", ""), "results": [],
                                                      "peer_id": self._peer_id, "elapsed_ms": 0.0}
            return await self._router.handle_search_request(payload, self._local_search_fn, requester=peer.peer_id)

        async def on_replicate(payload, peer: PeerInfo):
            if self._store_fn is None:
                return MessageType.REPLICATE_RESPONSE, {"success": False, "error": "no_store", "peer_id": self._peer_id}
            return await self._replicator.handle_replicate_request(payload, self._store_fn)

        async def on_index_submit(payload, peer: PeerInfo):
            from dataclasses import asdict

            if self._index_submit_receiver is None:
                return MessageType.INDEX_SUBMIT_ACK, {"url": payload.get("url", ""), "success": False,
                                                       "error": "not_an_indexer", "doc_id": 0, "peer_id": self._peer_id}
            if peer.peer_id and payload.get("peer_id") and payload["peer_id"] != peer.peer_id:
                return MessageType.INDEX_SUBMIT_ACK, {"url": payload.get("url", ""), "success": False,
                                                       "error": "peer_id_mismatch", "doc_id": 0, "peer_id": self._peer_id}
            ack = await asyncio.get_running_loop().run_in_executor(None, self._index_submit_receiver.handle_submit, payload,
                                                                   peer.peer_id)
            return MessageType.INDEX_SUBMIT_ACK, asdict(ack)

        async def on_pex(payload, peer: PeerInfo):
            who = peer.peer_id or peer.host
            if not self._pex.check_rate_limit(who):
                return MessageType.PEX_RESPONSE, {"peers": [], "error": "rate_limited"}
            n = payload.get("max_peers", PEX_MAX_PEERS)
            n = n if isinstance(n, int) and 0 < n <= PEX_MAX_PEERS else PEX_MAX_PEERS
            return MessageType.PEX_RESPONSE, {"peers": self._pex.build_response(self._get_connected_peer_addrs(), n)}

        async def on_credit_announce(payload, peer: PeerInfo):
            mgr = self._credit_sync_manager
            if mgr is None or not mgr.has_identity or payload.get("owner_email_hash") != mgr.owner_email_hash:
                return MessageType.ERROR, {"error": "no_match"}
            pid = str(payload.get("peer_id", "")) or peer.peer_id
            if pid:
                mgr.register_same_owner_peer(pid)
            return MessageType.CREDIT_SYNC_EXCHANGE, mgr.build_summary().to_dict()

        async def on_credit_exchange(payload, peer: PeerInfo):
            from infomesh_b200.credits.sync import CreditSummary

            mgr = self._credit_sync_manager
            if mgr is None or not mgr.has_identity:
                return MessageType.ERROR, {"error": "credit_sync_disabled"}
            try:
                mgr.receive_summary(CreditSummary.from_dict(payload), verify_signature=True)
            except Exception as exc:  # noqa: BLE001
                return MessageType.ERROR, {"error": f"bad_summary: {exc}"}
            return MessageType.CREDIT_SYNC_EXCHANGE, mgr.build_summary().to_dict()

        async def on_credit_proof(payload, peer: PeerInfo):
            from infomesh_b200.credits.verification import CreditProofBuilder

            mgr = self._credit_sync_manager
            ledger = getattr(mgr, "_ledger", None) if mgr is not None else None
            if ledger is None:
                return MessageType.ERROR, {"error": "no_ledger"}
            n = payload.get("sample_size", 10)
            proof = CreditProofBuilder(ledger, self._key_pair).build_proof(
                sample_size=n if isinstance(n, int) and 0 < n <= 50 else 10, request_id=str(payload.get("request_id", "")))
            return MessageType.CREDIT_PROOF_RESPONSE, proof

        async def on_llm(payload, peer: PeerInfo):
            if self._llm_handler is None:
                return MessageType.ERROR, {"error": "llm_unavailable"}
            return MessageType.LLM_RESPONSE, await self._llm_handler(payload, peer.peer_id)

        async def on_crawl_lock(payload, peer: PeerInfo):
            url = str(payload.get("url", ""))
            ok = bool(url) and await self._dht.acquire_crawl_lock(url)
            return MessageType.CRAWL_LOCK_ACK, {"url": url, "acquired": bool(ok)}

        t.register(MessageType.PING, on_ping)
        t.register(MessageType.SEARCH_REQUEST, on_search)
        t.register(MessageType.REPLICATE_REQUEST, on_replicate)
        t.register(MessageType.INDEX_SUBMIT, on_index_submit)
        t.register(MessageType.PEX_REQUEST, on_pex)
        t.register(MessageType.CREDIT_SYNC_ANNOUNCE, on_credit_announce)
        t.register(MessageType.CREDIT_SYNC_EXCHANGE, on_credit_exchange)
        t.register(MessageType.CREDIT_PROOF_REQUEST, on_credit_proof)
        t.register(MessageType.LLM_REQUEST, on_llm)
        t.register(MessageType.CRAWL_LOCK, on_crawl_lock)

    def _get_registered_protocols(self) -> list[str]:
        return list(ALL_PROTOCOLS)

    # ------------------------------------------------------------------ bootstrap
    async def _bootstrap(self) -> None:
        cfg = self._config
        addrs: list[str] = [a for a in cfg.network.bootstrap_nodes if a and a != "default"]
        use_bundled = "default" in cfg.network.bootstrap_nodes or not cfg.network.bootstrap_nodes
        results: dict[str, object] = {"configured": len(addrs), "connected": 0, "failed": 0, "sources": []}
        if use_bundled:
            try:
                found = await asyncio.wait_for(BS.discover_bootstrap_nodes(
                    static_nodes=BS.bundled_nodes(), dns_domain=cfg.network.bootstrap_dns_domain,
                    cache_dir=Path(cfg.node.data_dir), use_dns=cfg.network.bootstrap_dns,
                    use_github=cfg.network.bootstrap_github), timeout=15.0)
                addrs += found.addrs
                results["sources"] = found.sources_succeeded
            except Exception as exc:  # noqa: BLE001
                logger.debug("bootstrap_discovery_failed", error=str(exc))
        targets: list[tuple[str, int]] = []
        for a in dict.fromkeys(addrs):
            try:
                h, p, _ = parse_multiaddr(a)
                targets.append((h, p))
            except ValueError:
                results["failed"] = int(results["failed"]) + 1
        ok = await self._kad.bootstrap(targets) if targets else 0
        results["connected"] = ok
        results["failed"] = int(results["failed"]) + len(targets) - ok
        if ok == 0:
            ok += await self._connect_cached_peers()
            results["from_peer_store"] = ok
        self._bootstrap_results = results

    async def _connect_cached_peers(self) -> int:
        if self._peer_store is None:
            return 0
        n = 0
        for cp in self._peer_store.load_recent():
            try:
                h, p, _ = parse_multiaddr(cp.multiaddr)
            except ValueError:
                continue
            if await self._kad.ping(h, p):
                self._peer_store.upsert(cp.peer_id, cp.multiaddr)
                n += 1
            else:
                self._peer_store.record_failure(cp.peer_id)
        if n:
            await self._kad.lookup_nodes(int(self._peer_id, 16))
        return n

    def _get_connected_peer_addrs(self) -> list[tuple[str, str]]:
        return [(c.peer_id, format_multiaddr(c.host, c.port, c.peer_id)) for c in self._kad.connected_contacts()]

    async def _save_connected_peers(self) -> None:
        if self._peer_store is not None and self._kad is not None:
            try:
                self._peer_store.save_connected(self._get_connected_peer_addrs())
            except Exception:  # noqa: BLE001
                logger.debug("peer_store_save_failed")

    # ------------------------------------------------------------------ maintenance
    async def _run_main_loop(self) -> None:
        now = time.time()
        last = dict(refresh=now, status=now, pex=now, credit=now, version=now)
        while not self._stop_async.is_set():
            try:
                await asyncio.wait_for(self._stop_async.wait(), timeout=1.0)
                break
            except asyncio.TimeoutError:
                pass
            now = time.time()
            if now - last["refresh"] >= _ROUTING_REFRESH_INTERVAL:
                last["refresh"] = now
                await self._refresh_routing_table()
                await self._save_connected_peers()
                if self._peer_store is not None:
                    try:
                        self._peer_store.prune()
                    except Exception:  # noqa: BLE001
                        pass
            if now - last["pex"] >= PEX_ROUND_INTERVAL:
                last["pex"] = now
                try:
                    await self._run_pex_round()
                except Exception:  # noqa: BLE001
                    logger.debug("pex_round_failed")
            if now - last["credit"] >= _CREDIT_SYNC_INTERVAL:
                last["credit"] = now
                try:
                    await self._run_credit_sync_round()
                except Exception:  # noqa: BLE001
                    logger.debug("credit_sync_round_failed")
            if now - last["status"] >= _STATUS_WRITE_INTERVAL:
                last["status"] = now
                self._write_status_file()
            try:
                await self._connect_mdns_peers()
            except Exception:  # noqa: BLE001
                pass
            if now - last["version"] >= _VERSION_CHECK_INTERVAL:
                last["version"] = now
                upd = self._peer_version_tracker.check_peer_update()
                if upd is not None:
                    logger.info("update_available_from_peers", current=upd.current, latest=upd.latest)

    async def _refresh_routing_table(self) -> None:
        if self._kad is not None and len(self._kad.table):
            await self._kad.lookup_nodes(int(self._peer_id, 16))

    async def _connect_mdns_peers(self) -> None:
        if self._mdns is None:
            return
        known = set(self.get_connected_peers())
        for pid, peer in self._mdns.discovered_peers.items():
            if pid not in known:
                await self._kad.ping(peer.host, peer.port)

    async def _run_pex_round(self) -> int:
        contacts = self._kad.connected_contacts()
        if not contacts:
            return 0
        self._pex.cleanup_rate_limits()
        import random

        added = 0
        known = set(self.get_connected_peers())
        for c in random.sample(contacts, min(PEX_MAX_PEERS_PER_ROUND, len(contacts))):
            try:
                reply = await self._transport.request((c.host, c.port), MessageType.PEX_REQUEST, {"max_peers": PEX_MAX_PEERS},
                                                      timeout=5.0)
            except Exception:  # noqa: BLE001
                continue
            if not reply or reply[0] != MessageType.PEX_RESPONSE:
                continue
            for info in self._pex.process_response(c.peer_id, reply[1].get("peers", []) or [], known):
                try:
                    h, p, _ = parse_multiaddr(info.multiaddr)
                except ValueError:
                    continue
                got = await self._kad.ping(h, p)
                if got is not None and got.peer_id == info.peer_id:
                    known.add(info.peer_id)
                    added += 1
                    if self._peer_store is not None:
                        self._peer_store.upsert(info.peer_id, info.multiaddr)
        return added

    async def _announce_credit_sync(self) -> None:
        mgr = self._credit_sync_manager
        if mgr is None or not mgr.has_identity:
            return
        from infomesh_b200.credits.sync import CreditSummary

        for pid in self.get_connected_peers():
            try:
                reply = await self._send_to_peer(pid, MessageType.CREDIT_SYNC_ANNOUNCE,
                                                 {"peer_id": self._peer_id, "owner_email_hash": mgr.owner_email_hash}, 5.0)
                if reply and reply[0] == MessageType.CREDIT_SYNC_EXCHANGE:
                    mgr.register_same_owner_peer(pid)
                    mgr.receive_summary(CreditSummary.from_dict(reply[1]), verify_signature=True)
            except Exception:  # noqa: BLE001
                logger.debug("credit_sync_announce_failed", target=pid[:16])

    async def _run_credit_sync_round(self) -> None:
        mgr = self._credit_sync_manager
        if mgr is None or not mgr.has_identity:
            return
        from infomesh_b200.credits.sync import CreditSummary

        mgr.purge_stale()
        for pid in mgr.get_same_owner_peers():
            if not mgr.needs_sync(pid):
                continue
            try:
                reply = await self._send_to_peer(pid, MessageType.CREDIT_SYNC_EXCHANGE, mgr.build_summary().to_dict(), 5.0)
                if reply and reply[0] == MessageType.CREDIT_SYNC_EXCHANGE:
                    mgr.receive_summary(CreditSummary.from_dict(reply[1]), verify_signature=True)
            except Exception:  # noqa: BLE001
                logger.debug("credit_sync_round_failed", target=pid[:16])

    # ------------------------------------------------------------------ status file
    def _write_status_file(self, *, state: str | None = None, error: str = "") -> None:
        path = Path(self._config.node.data_dir) / "p2p_status.json"
        try:
            running = self._state == NodeState.RUNNING
            peer_ids = self.get_connected_peers() if running else []
            ds, ts = (self._dht.stats if self._dht else None), self._throttle.stats
            data = {
                "state": state or str(self._state), "peer_id": self._peer_id, "peers": len(peer_ids), "peer_ids": peer_ids,
                "listen_addrs": self.listen_addrs if running else [], "timestamp": time.time(), "error": error,
                "dht": ({"keys_stored": ds.keys_stored, "keys_published": ds.keys_published,
                         "gets_performed": ds.gets_performed, "puts_performed": ds.puts_performed} if ds else {}),
                "bandwidth": {"upload_bytes": ts.upload_bytes, "download_bytes": ts.download_bytes,
                              "upload_waits": ts.upload_waits, "download_waits": ts.download_waits},
                "bootstrap": self._bootstrap_results, "peer_versions": self._peer_version_tracker.peer_versions,
            }
            path.parent.mkdir(parents=True, exist_ok=True)
            tmp = path.with_suffix(".tmp")
            tmp.write_text(json.dumps(data))
            tmp.replace(path)
        except OSError:
            pass

    # ------------------------------------------------------------------ public API (any thread / any loop)
    def _submit(self, coro) -> "asyncio.Future":
        if self._loop is None or self._state != NodeState.RUNNING:
            coro.close()
            raise RuntimeError("P2P node is not running")
        return asyncio.wrap_future(asyncio.run_coroutine_threadsafe(coro, self._loop))

    def run_sync(self, coro, timeout: float = 30.0):
        if self._loop is None or self._state != NodeState.RUNNING:
            coro.close()
            raise RuntimeError("P2P node is not running")
        return asyncio.run_coroutine_threadsafe(coro, self._loop).result(timeout)

    async def connect(self, addr: str) -> bool:
        h, p, _ = parse_multiaddr(addr)
        if asyncio.get_running_loop() is self._loop:
            return await self._kad.ping(h, p) is not None
        return await self._submit(self._kad.ping(h, p)) is not None

    async def search_network(self, query: str, keywords: list[str], limit: int = 10) -> list[dict[str, object]]:
        if self._router is None or self._state != NodeState.RUNNING:
            return []
        try:
            res = await asyncio.wait_for(self._submit(self._router.route_query(query, keywords, limit)),
                                         timeout=_SEARCH_NETWORK_TIMEOUT)
        except (asyncio.TimeoutError, RuntimeError):
            logger.warning("search_network_timeout", query=query[:60])
            return []
        return [{"url": r.url, "title": r.title, "snippet": r.snippet, "score": r.score, "peer_id": r.peer_id,
                 "doc_id": r.doc_id} for r in res]

    async def publish_document_to_network(self, doc_id: int, url: str, title: str, text: str, score: float = 1.0) -> int:
        return await self.publish_documents_to_network([{"doc_id": doc_id, "url": url, "title": title, "text": text,
                                                          "score": score}])

    async def publish_documents_to_network(self, documents: list[dict[str, object]]) -> int:
        if self._distributed_index is None or self._state != NodeState.RUNNING or not documents:
            return 0
        try:
            return await asyncio.wait_for(self._submit(self._distributed_index.publish_batch(documents)),
                                          timeout=_PUBLISH_TIMEOUT)
        except Exception:  # noqa: BLE001
            logger.warning("publish_network_failed", documents=len(documents))
            return 0

    async def replicate_document(self, doc_id: int, url: str, title: str, text: str, text_hash: str, language: str = "") -> int:
        if self._replicator is None or self._state != NodeState.RUNNING:
            return 0
        return await self._submit(self._replicator.replicate_document(doc_id, url, title, text, text_hash, language))
