"""Who crawls which URL -- decided by every peer on its own, with the same answer everywhere.

Contract (SURVEY §2.1 crawler/ "url assigner"; reference infomesh/crawler/url_assigner.py): a URL belongs to the known
peer whose ``SHA-256(peer_id)`` is closest to ``SHA-256(url)`` under XOR distance (ties broken by peer id); the local peer
is always a candidate and cannot be removed; ``assign`` wraps a URL in a ``CrawlAssignment`` issued by the local peer.

Implementation: peer digests are converted to integers once, when the peer is added, so an ownership query is one hash of
the URL plus an integer XOR per peer -- no hex parsing in the loop."""
from __future__ import annotations

from infomesh_b200.hashing import content_hash
from infomesh_b200.p2p.protocol import CrawlAssignment


def _ring_position(text: str) -> int:
    """Position of ``text`` in the 256-bit key space."""
    return int(content_hash(text), 16)


class UrlAssigner:
    def __init__(self, local_peer_id: str):
        self._local = local_peer_id
        self._positions: dict[str, int] = {}
        self.add_peer(local_peer_id)

    # ---- membership
    def add_peer(self, peer_id: str) -> None:
        if peer_id not in self._positions:
            self._positions[peer_id] = _ring_position(peer_id)

    def remove_peer(self, peer_id: str) -> None:
        if peer_id != self._local:
            self._positions.pop(peer_id, None)

    @property
    def known_peers(self) -> int:
        return len(self._positions)

    # ---- ownership
    def closest_peer(self, url: str) -> str:
        target = _ring_position(url)
        return min(self._positions, key=lambda peer: (self._positions[peer] ^ target, peer))

    def is_local_owner(self, url: str) -> bool:
        return self.closest_peer(url) == self._local

    def filter_local_urls(self, urls: list[str]) -> list[str]:
        return list(filter(self.is_local_owner, urls))

    def assign(self, url: str, *, depth: int = 0) -> CrawlAssignment:
        return CrawlAssignment(url=url, depth=depth, priority=1.0, assigner_peer_id=self._local)
