"""Merkle tree over document / credit-entry hashes with membership proofs and a signable root record
(reference infomesh/trust/merkle.py:33-331): leaves = sha256("leaf:" + h), nodes = sha256(left + right) over the
hex strings, an odd node is paired with itself."""
from __future__ import annotations

import hashlib
import time
from dataclasses import dataclass
from typing import Any

from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class ProofSide:
    LEFT = "left"
    RIGHT = "right"


@dataclass(frozen=True)
class MerkleProof:
    doc_hash: str                                  # leaf hash
    proof_path: tuple[tuple[str, str], ...]        # ((sibling hash, side), ...) leaf -> root
    root_hash: str
    leaf_index: int


@dataclass(frozen=True)
class MerkleRoot:
    root_hash: str
    document_count: int
    built_at: float
    peer_id: str
    signature: bytes = b""


def _hash_pair(left: str, right: str) -> str:
    return hashlib.sha256((left + right).encode("ascii")).hexdigest()


def _hash_leaf(data: str) -> str:
    return hashlib.sha256(("leaf:" + data).encode("utf-8")).hexdigest()


class MerkleTree:
    def __init__(self):
        self._levels: list[list[str]] = []
        self._built_at = 0.0

    @property
    def root_hash(self) -> str:
        return self._levels[-1][0] if self._levels else ""

    @property
    def leaf_count(self) -> int:
        return len(self._levels[0]) if self._levels else 0

    @property
    def built_at(self) -> float:
        return self._built_at

    @property
    def height(self) -> int:
        return len(self._levels)

    def build(self, document_hashes: list[str]) -> str:
        if not document_hashes:
            raise ValueError("Cannot build Merkle tree from empty hash list")
        self._built_at = time.time()
        level = [_hash_leaf(h) for h in document_hashes]
        self._levels = [level]
        while len(level) > 1:
            padded = level + [level[-1]] if len(level) % 2 else level
            level = [_hash_pair(padded[i], padded[i + 1]) for i in range(0, len(padded), 2)]
            self._levels.append(level)
        return self.root_hash

    def get_proof(self, leaf_index: int) -> MerkleProof:
        if not self._levels:
            raise RuntimeError("Merkle tree not built yet")
        if not 0 <= leaf_index < self.leaf_count:
            raise IndexError(f"leaf_index {leaf_index} out of range [0, {self.leaf_count})")
        path: list[tuple[str, str]] = []
        idx = leaf_index
        for level in self._levels[:-1]:
            if idx % 2 == 0:
                sib = level[idx + 1] if idx + 1 < len(level) else level[idx]
                path.append((sib, ProofSide.RIGHT))
            else:
                path.append((level[idx - 1], ProofSide.LEFT))
            idx //= 2
        return MerkleProof(self._levels[0][leaf_index], tuple(path), self.root_hash, leaf_index)

    @staticmethod
    def verify_proof(proof: MerkleProof) -> bool:
        cur = proof.doc_hash
        for sib, side in proof.proof_path:
            cur = _hash_pair(sib, cur) if side == ProofSide.LEFT else _hash_pair(cur, sib)
        return cur == proof.root_hash

    @staticmethod
    def verify_document(document_hash: str, proof: MerkleProof) -> bool:
        return _hash_leaf(document_hash) == proof.doc_hash and MerkleTree.verify_proof(proof)

    def root_payload(self, peer_id: str) -> bytes:
        return f"{self.root_hash}|{self.leaf_count}|{self._built_at}|{peer_id}".encode()

    def create_root_record(self, peer_id: str, key_pair: KeyPairLike | None = None) -> MerkleRoot:
        sig = key_pair.sign(self.root_payload(peer_id)) if key_pair is not None else b""
        return MerkleRoot(self.root_hash, self.leaf_count, self._built_at, peer_id, sig)


def verify_root_record(root: MerkleRoot, public_key: bytes) -> bool:
    from infomesh_b200.p2p.keys import verify_with_public_key

    payload = f"{root.root_hash}|{root.document_count}|{root.built_at}|{root.peer_id}".encode()
    return verify_with_public_key(public_key, payload, root.signature)


def serialize_merkle_root(root: MerkleRoot) -> dict[str, Any]:
    return {"root_hash": root.root_hash, "document_count": root.document_count, "built_at": root.built_at,
            "peer_id": root.peer_id, "signature": root.signature.hex()}


def deserialize_merkle_root(data: dict[str, Any]) -> MerkleRoot:
    return MerkleRoot(data["root_hash"], data["document_count"], data["built_at"], data["peer_id"],
                      bytes.fromhex(data.get("signature", "")))


def serialize_proof(proof: MerkleProof) -> dict[str, Any]:
    return {"doc_hash": proof.doc_hash, "proof_path": [(h, s) for h, s in proof.proof_path],
            "root_hash": proof.root_hash, "leaf_index": proof.leaf_index}


def deserialize_proof(data: dict[str, Any]) -> MerkleProof:
    return MerkleProof(data["doc_hash"], tuple((h, s) for h, s in data["proof_path"]), data["root_hash"],
                       data["leaf_index"])
