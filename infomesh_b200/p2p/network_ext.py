"""Network extras: STUN-based NAT probe, DNS peer discovery, geographic proximity ordering (haversine), partition
detection with recovery actions, relay selection (reference infomesh/p2p/network_ext.py:28-329).  The STUN probe here
actually parses XOR-MAPPED-ADDRESS instead of assuming the server echoes our own port."""
from __future__ import annotations

import math
import os
import socket
import struct
import time
from dataclasses import dataclass, field

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)
_STUN_COOKIE = 0x2112A442


@dataclass(frozen=True)
class NATInfo:
    nat_type: str            # none | full_cone | restricted | symmetric | unknown
    external_ip: str
    external_port: int
    internal_ip: str
    internal_port: int


def parse_stun_response(data: bytes, txn_id: bytes) -> tuple[str, int] | None:
    if len(data) < 20 or data[8:20] != txn_id:
        return None
    mtype, mlen = struct.unpack("!HH", data[:4])
    if mtype != 0x0101:
        return None
    off, end = 20, min(len(data), 20 + mlen)
    while off + 4 <= end:
        atype, alen = struct.unpack("!HH", data[off:off + 4])
        val = data[off + 4:off + 4 + alen]
        if atype in (0x0020, 0x0001) and alen >= 8 and val[1] == 0x01:
            port, = struct.unpack("!H", val[2:4])
            ip = struct.unpack("!I", val[4:8])[0]
            if atype == 0x0020:
                port ^= _STUN_COOKIE >> 16
                ip ^= _STUN_COOKIE
            return socket.inet_ntoa(struct.pack("!I", ip)), port
        off += 4 + alen + (-alen % 4)
    return None


async def detect_nat_type(stun_server: str = "stun.l.google.com", stun_port: int = 19302, *, timeout: float = 3.0) -> NATInfo:
    import asyncio

    def probe() -> NATInfo:
        iip, iport, eip, eport, kind = "0.0.0.0", 0, "", 0, "unknown"
        try:
            with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:
                s.settimeout(timeout)
                s.bind(("", 0))
                iport = s.getsockname()[1]
                txn = os.urandom(12)
                dest = socket.getaddrinfo(stun_server, stun_port, socket.AF_INET)[0][4]
                s.connect(dest)
                iip = s.getsockname()[0]
                s.send(struct.pack("!HHI", 0x0001, 0, _STUN_COOKIE) + txn)
                try:
                    mapped = parse_stun_response(s.recv(1024), txn)
                    if mapped:
                        eip, eport = mapped
                        kind = "none" if eip == iip else ("full_cone" if eport == iport else "restricted")
                except (TimeoutError, socket.timeout):
                    kind = "symmetric"
        except OSError as exc:
            logger.debug("nat_detection_failed", error=str(exc))
        return NATInfo(kind, eip, eport, iip, iport)

    return await asyncio.get_running_loop().run_in_executor(None, probe)


@dataclass(frozen=True)
class DNSPeer:
    host: str
    port: int
    peer_id: str = ""
    priority: int = 0
    source: str = "dns"          # which record type produced it (after the reference's fields, so positional construction matches)


def discover_peers_dns(domain: str = "infomesh.io", *, default_port: int = 4001) -> list[DNSPeer]:
    """SRV records first, then plain A/AAAA records of ``peers.<domain>``."""
    from infomesh_b200.p2p.bootstrap import SRV_SERVICE, _resolve_srv

    peers: list[DNSPeer] = []
    try:
        peers += [DNSPeer(h, p, source="dns_srv") for h, p in _resolve_srv(f"{SRV_SERVICE}.{domain}")]
    except OSError:
        pass
    if not peers:
        try:
            for info in socket.getaddrinfo(f"peers.{domain}", default_port, proto=socket.IPPROTO_TCP):
                peers.append(DNSPeer(info[4][0], default_port, source="dns_a"))
        except OSError:
            pass
    return list(dict.fromkeys(peers))


@dataclass(frozen=True)
class GeoLocation:
    country: str = ""
    region: str = ""
    city: str = ""
    latitude: float = 0.0
    longitude: float = 0.0


def estimate_geo_distance(loc1: GeoLocation, loc2: GeoLocation) -> float:
    """Great-circle distance in km."""
    p1, p2 = math.radians(loc1.latitude), math.radians(loc2.latitude)
    dphi, dlmb = p2 - p1, math.radians(loc2.longitude - loc1.longitude)
    h = math.sin(dphi / 2) ** 2 + math.cos(p1) * math.cos(p2) * math.sin(dlmb / 2) ** 2
    return 6371.0 * 2 * math.atan2(math.sqrt(h), math.sqrt(1 - h))


def sort_peers_by_proximity(peers: list[tuple[str, GeoLocation]], my_location: GeoLocation) -> list[tuple[str, float]]:
    return sorted(((pid, estimate_geo_distance(my_location, loc)) for pid, loc in peers), key=lambda x: x[1])


@dataclass
class PartitionState:
    is_partitioned: bool = False
    reachable_peers: int = 0
    expected_peers: int = 0
    last_check: float = 0.0
    recovery_attempts: int = 0


class PartitionDetector:
    def __init__(self, threshold: float = 0.5):
        self._threshold = threshold
        self._state = PartitionState()

    def check(self, reachable: int, total: int) -> PartitionState:
        st = self._state
        st.reachable_peers, st.expected_peers, st.last_check = reachable, total, time.time()
        st.is_partitioned = total > 0 and reachable / total < self._threshold
        if not st.is_partitioned:
            st.recovery_attempts = 0
        return st

    def get_recovery_actions(self) -> list[str]:
        if not self._state.is_partitioned:
            return []
        actions = ["Reconnect to bootstrap nodes", "Refresh routing table", "Re-announce local index to DHT"]
        if self._state.recovery_attempts > 3:
            actions.append("Consider restarting the node")
        self._state.recovery_attempts += 1
        return actions


@dataclass
class RelayConfig:
    enabled: bool = False
    max_relay_connections: int = 10
    max_bandwidth_mbps: float = 5.0
    relay_peers: list[str] = field(default_factory=list)


def select_relay(available_relays: list[tuple[str, float]]) -> str | None:
    return min(available_relays, key=lambda r: r[1])[0] if available_relays else None
