"""Peer exchange gossip: PEX_REQUEST {max_peers} -> PEX_RESPONSE {peers:[{peer_id, multiaddr}]}; <= 10 peers per
response, one request / peer / 60 s, never advertises self, accepts only /ip4|/ip6 .. /p2p/ multiaddrs
(reference infomesh/p2p/pex.py:38-171)."""
from __future__ import annotations

import time
from dataclasses import dataclass

PEX_MAX_PEERS = 10
PEX_MIN_INTERVAL = 60
PEX_ROUND_INTERVAL = 300
PEX_MAX_PEERS_PER_ROUND = 3


@dataclass(frozen=True)
class PEXPeerInfo:
    peer_id: str
    multiaddr: str


def _is_valid_multiaddr(maddr: object) -> bool:
    return isinstance(maddr, str) and "/p2p/" in maddr and maddr.startswith(("/ip4/", "/ip6/"))


class PeerExchange:
    def __init__(self, peer_id: str):
        self._me = peer_id
        self._last_request: dict[str, float] = {}

    def check_rate_limit(self, requester_id: str) -> bool:
        now = time.time()
        if now - self._last_request.get(requester_id, 0.0) < PEX_MIN_INTERVAL:
            return False
        self._last_request[requester_id] = now
        return True

    def build_response(self, connected_peers: list[tuple[str, str]], max_peers: int = PEX_MAX_PEERS) -> list[dict[str, str]]:
        out = [{"peer_id": pid, "multiaddr": addr} for pid, addr in connected_peers
               if pid != self._me and _is_valid_multiaddr(addr)]
        return out[:min(max_peers, PEX_MAX_PEERS)]

    def process_response(self, sender_id: str, peers_data: list[dict[str, object]],
                         known_peers: set[str] | None = None) -> list[PEXPeerInfo]:
        known = known_peers or set()
        fresh: list[PEXPeerInfo] = []
        seen: set[str] = set()
        for entry in peers_data[:PEX_MAX_PEERS]:
            if not isinstance(entry, dict):
                continue
            pid, addr = str(entry.get("peer_id", "")), entry.get("multiaddr", "")
            if not pid or pid in (self._me, sender_id) or pid in known or pid in seen or not _is_valid_multiaddr(addr):
                continue
            seen.add(pid)
            fresh.append(PEXPeerInfo(pid, str(addr)))
        return fresh

    def cleanup_rate_limits(self) -> None:
        cutoff = time.time() - PEX_MIN_INTERVAL * 10
        for pid in [p for p, ts in self._last_request.items() if ts < cutoff]:
            del self._last_request[pid]
