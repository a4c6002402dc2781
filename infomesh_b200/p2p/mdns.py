"""LAN discovery: every 30 s each node multicasts ``b"INFOMESH" + msgpack{peer_id, port, ts}`` to 224.0.0.251:5353;
listeners keep peers for 120 s (reference infomesh/p2p/mdns.py:26-252).  One thread runs both duties with a
socket timeout instead of the reference's two threads."""
from __future__ import annotations

import contextlib
import socket
import struct
import threading
import time
from dataclasses import dataclass, field

import msgpack

from infomesh_b200.p2p.protocol import safe_unpackb
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MDNS_GROUP = "224.0.0.251"
MDNS_PORT = 5353
SERVICE_TYPE = "_infomesh._tcp.local."
ANNOUNCE_INTERVAL = 30.0
PEER_TTL = 120.0
MAGIC = b"INFOMESH"


@dataclass
class DiscoveredPeer:
    peer_id: str
    host: str
    port: int
    last_seen: float = field(default_factory=time.monotonic)

    @property
    def is_stale(self) -> bool:
        return time.monotonic() - self.last_seen > PEER_TTL


class MDNSDiscovery:
    def __init__(self, peer_id: str, port: int = 4001, *, group: str = MDNS_GROUP, mdns_port: int = MDNS_PORT):
        self._peer_id, self._port = peer_id, port
        self._group, self._mport = group, mdns_port
        self._peers: dict[str, DiscoveredPeer] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None
        self._sock: socket.socket | None = None

    @property
    def discovered_peers(self) -> dict[str, DiscoveredPeer]:
        with self._lock:
            for pid in [p for p, v in self._peers.items() if v.is_stale]:
                del self._peers[pid]
            return dict(self._peers)

    @property
    def peer_count(self) -> int:
        return len(self.discovered_peers)

    def build_announce(self) -> bytes:
        return MAGIC + msgpack.packb({"peer_id": self._peer_id, "port": self._port, "ts": time.time()}, use_bin_type=True)

    def parse_announce(self, data: bytes, addr: tuple[str, int]) -> DiscoveredPeer | None:
        if len(data) < len(MAGIC) + 3 or not data.startswith(MAGIC):
            return None
        try:
            body = safe_unpackb(data[len(MAGIC):])
        except Exception:  # noqa: BLE001
            return None
        if not isinstance(body, dict):
            return None
        pid, port = body.get("peer_id", ""), body.get("port", 0)
        if not pid or not isinstance(port, int) or not 0 < port < 65536 or pid == self._peer_id:
            return None
        return DiscoveredPeer(str(pid), addr[0], port)

    # reference-compatible private aliases
    _build_announce = build_announce
    _parse_announce = parse_announce

    def observe(self, data: bytes, addr: tuple[str, int]) -> DiscoveredPeer | None:
        peer = self.parse_announce(data, addr)
        if peer:
            with self._lock:
                if peer.peer_id not in self._peers:
                    logger.info("mdns_peer_discovered", peer_id=peer.peer_id[:16], host=peer.host, port=peer.port)
                self._peers[peer.peer_id] = peer
        return peer

    def _open(self) -> socket.socket:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM, socket.IPPROTO_UDP)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        if hasattr(socket, "SO_REUSEPORT"):
            with contextlib.suppress(OSError):
                s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEPORT, 1)
        s.bind(("", self._mport))
        s.setsockopt(socket.IPPROTO_IP, socket.IP_ADD_MEMBERSHIP,
                     struct.pack("4sL", socket.inet_aton(self._group), socket.INADDR_ANY))
        s.setsockopt(socket.IPPROTO_IP, socket.IP_MULTICAST_LOOP, 0)
        s.setsockopt(socket.IPPROTO_IP, socket.IP_MULTICAST_TTL, 1)
        s.settimeout(1.0)
        return s

    def start(self) -> bool:
        if self._thread is not None:
            return True
        try:
            self._sock = self._open()
        except OSError as exc:
            logger.warning("mdns_unavailable", error=str(exc))
            return False
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True, name="mdns")
        self._thread.start()
        return True

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=3.0)
        if self._sock is not None:
            with contextlib.suppress(OSError):
                self._sock.close()
        self._thread = self._sock = None

    def _loop(self) -> None:
        next_announce = 0.0
        while not self._stop.is_set():
            now = time.monotonic()
            try:
                if now >= next_announce:
                    self._sock.sendto(self.build_announce(), (self._group, self._mport))
                    next_announce = now + ANNOUNCE_INTERVAL
                data, addr = self._sock.recvfrom(1024)
                self.observe(data, addr)
            except (TimeoutError, socket.timeout):
                continue
            except OSError:
                if self._stop.is_set():
                    break
                time.sleep(1.0)
