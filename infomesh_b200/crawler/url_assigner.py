"""Deterministic URL -> owning peer assignment by XOR distance between SHA-256(url) and SHA-256(peer_id)
(reference infomesh/crawler/url_assigner.py:26-152) so peers split the crawl frontier without coordination."""
from __future__ import annotations

from infomesh_b200.hashing import content_hash
from infomesh_b200.p2p.protocol import CrawlAssignment


def _xor_distance(hex_a: str, hex_b: str) -> int:
    return int(hex_a, 16) ^ int(hex_b, 16)


class UrlAssigner:
    def __init__(self, local_peer_id: str):
        self._local = local_peer_id
        self._peers: dict[str, str] = {local_peer_id: content_hash(local_peer_id)}

    def add_peer(self, peer_id: str) -> None:
        self._peers.setdefault(peer_id, content_hash(peer_id))

    def remove_peer(self, peer_id: str) -> None:
        if peer_id != self._local:
            self._peers.pop(peer_id, None)

    @property
    def known_peers(self) -> int:
        return len(self._peers)

    def closest_peer(self, url: str) -> str:
        h = content_hash(url)
        return min(self._peers.items(), key=lambda kv: (_xor_distance(h, kv[1]), kv[0]))[0]

    def is_local_owner(self, url: str) -> bool:
        return self.closest_peer(url) == self._local

    def assign(self, url: str, *, depth: int = 0) -> CrawlAssignment:
        return CrawlAssignment(url=url, depth=depth, priority=1.0, assigner_peer_id=self._local)

    def filter_local_urls(self, urls: list[str]) -> list[str]:
        return [u for u in urls if self.is_local_owner(u)]
