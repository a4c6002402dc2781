"""Sybil / eclipse defences: proof-of-work node identities (SHA-256(pubkey || nonce_le64) with >= 20 leading zero
bits), node id = first 160 bits of that digest, and a per-bucket /24 (v4) or /48 (v6) subnet cap of 3
(reference infomesh/p2p/sybil.py:27-395).  The search loop works on batches through hashlib so the GIL is released
for the digest itself."""
from __future__ import annotations

import hashlib
import ipaddress
import struct
import time
from collections import defaultdict
from dataclasses import dataclass, field

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_DIFFICULTY_BITS = 20
DEFAULT_MAX_PER_SUBNET = 3


@dataclass(frozen=True)
class ProofOfWork:
    nonce: int
    difficulty_bits: int
    hash_hex: str
    elapsed_seconds: float


def leading_zero_bits(digest: bytes) -> int:
    n = int.from_bytes(digest, "big")
    return len(digest) * 8 - n.bit_length()


def compute_pow_hash(public_key_bytes: bytes, nonce: int) -> bytes:
    return hashlib.sha256(public_key_bytes + struct.pack("<Q", nonce)).digest()


def generate_pow(public_key_bytes: bytes, difficulty_bits: int = DEFAULT_DIFFICULTY_BITS, *,
                 max_nonce: int = 2 ** 48, progress_interval: int = 1_000_000) -> ProofOfWork:
    t0 = time.monotonic()
    base = hashlib.sha256(public_key_bytes)        # reuse the absorbed prefix for every candidate
    limit = 1 << (256 - difficulty_bits)
    pack = struct.Struct("<Q").pack
    for nonce in range(max_nonce):
        h = base.copy()
        h.update(pack(nonce))
        d = h.digest()
        if int.from_bytes(d, "big") < limit:
            dt = time.monotonic() - t0
            logger.info("pow_found", nonce=nonce, difficulty=difficulty_bits, elapsed_seconds=round(dt, 2))
            return ProofOfWork(nonce, difficulty_bits, d.hex(), dt)
        if nonce and nonce % progress_interval == 0:
            logger.debug("pow_progress", nonces_tried=nonce)
    raise RuntimeError(f"PoW failed: no valid nonce found in {max_nonce} attempts")


def verify_pow(public_key_bytes: bytes, nonce: int, difficulty_bits: int = DEFAULT_DIFFICULTY_BITS) -> bool:
    return leading_zero_bits(compute_pow_hash(public_key_bytes, nonce)) >= difficulty_bits


def derive_node_id(public_key_bytes: bytes, nonce: int) -> str:
    return compute_pow_hash(public_key_bytes, nonce).hex()[:40]


def subnet_of(ip: str) -> str:
    addr = ipaddress.ip_address(ip)
    prefix = 24 if addr.version == 4 else 48
    return str(ipaddress.ip_network(f"{ip}/{prefix}", strict=False))


@dataclass
class SubnetLimiter:
    max_per_subnet: int = DEFAULT_MAX_PER_SUBNET
    _buckets: dict = field(default_factory=lambda: defaultdict(lambda: defaultdict(set)))

    def _get_subnet(self, ip: str) -> str:
        return subnet_of(ip)

    def can_add(self, ip: str, bucket_id: int) -> bool:
        return len(self._buckets[bucket_id][subnet_of(ip)]) < self.max_per_subnet

    def add(self, ip: str, peer_id: str, bucket_id: int) -> bool:
        members = self._buckets[bucket_id][subnet_of(ip)]
        if peer_id in members:
            return True
        if len(members) >= self.max_per_subnet:
            logger.warning("subnet_limit_reached", subnet=subnet_of(ip), bucket_id=bucket_id, rejected_peer=peer_id)
            return False
        members.add(peer_id)
        return True

    def remove(self, ip: str, peer_id: str, bucket_id: int) -> None:
        sn = subnet_of(ip)
        bucket = self._buckets.get(bucket_id)
        if not bucket or sn not in bucket:
            return
        bucket[sn].discard(peer_id)
        if not bucket[sn]:
            del bucket[sn]
        if not bucket:
            del self._buckets[bucket_id]

    def get_subnet_counts(self, bucket_id: int) -> dict[str, int]:
        return {sn: len(p) for sn, p in self._buckets.get(bucket_id, {}).items() if p}

    def total_nodes(self) -> int:
        return sum(len(p) for b in self._buckets.values() for p in b.values())


@dataclass
class SybilValidator:
    """PoW valid -> derived id matches the claimed id -> subnet quota free."""
    difficulty_bits: int = DEFAULT_DIFFICULTY_BITS
    max_per_subnet: int = DEFAULT_MAX_PER_SUBNET
    subnet_limiter: SubnetLimiter = field(init=False)

    def __post_init__(self) -> None:
        self.subnet_limiter = SubnetLimiter(max_per_subnet=self.max_per_subnet)

    def validate_peer(self, public_key_bytes: bytes, pow_nonce: int, ip: str, peer_id: str,
                      bucket_id: int) -> tuple[bool, str]:
        if not verify_pow(public_key_bytes, pow_nonce, self.difficulty_bits):
            logger.warning("sybil_pow_invalid", peer_id=peer_id[:16])
            return False, "invalid_pow"
        if derive_node_id(public_key_bytes, pow_nonce) != peer_id:
            logger.warning("sybil_id_mismatch", peer_id=peer_id[:16])
            return False, "node_id_mismatch"
        if not self.subnet_limiter.add(ip, peer_id, bucket_id):
            return False, "subnet_limit"
        return True, "ok"
