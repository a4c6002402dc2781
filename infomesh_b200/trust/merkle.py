"""Commitment to a set of document (or credit-entry) hashes: Merkle root, membership proofs, a signable root record.

Contract (SURVEY §2.1 trust/ "merkle"; reference infomesh/trust/merkle.py) -- this is wire format, peers verify each
other's proofs: a leaf is ``sha256("leaf:" + hash)``; an inner node is ``sha256(left_hex + right_hex)``; a level with an
odd number of nodes pairs its last node with itself; a proof lists ``(sibling, "L" | "R")`` from the leaf upward;
the signed root payload is ``root|count|built_at|peer_id``.

Implementation: the tree keeps its levels as tuples built by one ``_fold`` step applied until a single node remains; a
proof is read off by walking ``index ^ 1`` (the sibling) and ``index >> 1`` (the parent); verification is a left fold over
the path; (de)serialisation goes through the dataclass field lists rather than hand-written dictionaries."""
from __future__ import annotations

import time
from dataclasses import asdict, dataclass
from enum import StrEnum
from functools import reduce
from hashlib import sha256
from typing import Any

from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class ProofSide(StrEnum):
    """Which side of the running hash a sibling goes on.  The one-letter values are what travels in serialised proofs."""
    LEFT = "L"
    RIGHT = "R"


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


def _hash_leaf(data: str) -> str:
    return sha256(b"leaf:" + data.encode("utf-8")).hexdigest()


def _hash_pair(left: str, right: str) -> str:
    return sha256(f"{left}{right}".encode("ascii")).hexdigest()


def _fold(level: tuple[str, ...]) -> tuple[str, ...]:
    """One level up: neighbours are hashed together, a trailing single node with itself."""
    evens, odds = level[0::2], level[1::2]
    if len(odds) < len(evens):
        odds += (evens[-1],)
    return tuple(map(_hash_pair, evens, odds))


def _climb(node: str, step: tuple[str, str]) -> str:
    sibling, side = step
    return _hash_pair(sibling, node) if side == ProofSide.LEFT else _hash_pair(node, sibling)


def _signing_payload(root_hash: str, count: int, built_at: float, peer_id: str) -> bytes:
    return "|".join((root_hash, str(count), str(built_at), peer_id)).encode()


class MerkleTree:
    def __init__(self):
        self._levels: tuple[tuple[str, ...], ...] = ()      # [0] = leaves ... [-1] = (root,)
        self._built_at = 0.0

    # ---- shape
    @property
    def height(self) -> int:
        return len(self._levels)

    @property
    def leaf_count(self) -> int:
        return len(self._levels[0]) if self._levels else 0

    @property
    def root_hash(self) -> str:
        return self._levels[-1][0] if self._levels else ""

    @property
    def built_at(self) -> float:
        return self._built_at

    # ---- construction
    def build(self, document_hashes: list[str]) -> str:
        if not document_hashes:
            raise ValueError("Cannot build Merkle tree from empty hash list")
        stack = [tuple(_hash_leaf(h) for h in document_hashes)]
        while len(stack[-1]) > 1:
            stack.append(_fold(stack[-1]))
        self._levels, self._built_at = tuple(stack), time.time()
        return self.root_hash

    # ---- proofs
    def get_proof(self, leaf_index: int) -> MerkleProof:
        if not self._levels:
            raise RuntimeError("Merkle tree not built yet")
        if leaf_index not in range(self.leaf_count):
            raise IndexError(f"leaf_index {leaf_index} out of range [0, {self.leaf_count})")
        steps = []
        position = leaf_index
        for level in self._levels[:-1]:
            mate = position ^ 1                                   # the other child of the same parent
            sibling = level[mate] if mate < len(level) else level[position]
            steps.append((sibling, ProofSide.LEFT if mate < position else ProofSide.RIGHT))
            position >>= 1
        return MerkleProof(self._levels[0][leaf_index], tuple(steps), self.root_hash, leaf_index)

    @staticmethod
    def verify_proof(proof: MerkleProof) -> bool:
        return reduce(_climb, proof.proof_path, proof.doc_hash) == proof.root_hash

    @staticmethod
    def verify_document(document_hash: str, proof: MerkleProof) -> bool:
        return proof.doc_hash == _hash_leaf(document_hash) and MerkleTree.verify_proof(proof)

    # ---- signed root
    def root_payload(self, peer_id: str) -> bytes:
        return _signing_payload(self.root_hash, self.leaf_count, self._built_at, peer_id)

    def create_root_record(self, peer_id: str, key_pair: KeyPairLike | None = None) -> MerkleRoot:
        signature = b"" if key_pair is None else key_pair.sign(self.root_payload(peer_id))
        return MerkleRoot(self.root_hash, self.leaf_count, self._built_at, peer_id, signature)


def verify_root_record(root: MerkleRoot, public_key: bytes) -> bool:
    from infomesh_b200.p2p.keys import verify_with_public_key

    payload = _signing_payload(root.root_hash, root.document_count, root.built_at, root.peer_id)
    return verify_with_public_key(public_key, payload, root.signature)


# ----------------------------------------------------------------------------- wire forms (msgpack / JSON friendly)
def serialize_merkle_root(root: MerkleRoot) -> dict[str, Any]:
    return {**asdict(root), "signature": root.signature.hex()}


def deserialize_merkle_root(data: dict[str, Any]) -> MerkleRoot:
    return MerkleRoot(**{**{k: data[k] for k in ("root_hash", "document_count", "built_at", "peer_id")},
                         "signature": bytes.fromhex(data.get("signature", ""))})


def serialize_proof(proof: MerkleProof) -> dict[str, Any]:
    return {**asdict(proof), "proof_path": [tuple(step) for step in proof.proof_path]}


def deserialize_proof(data: dict[str, Any]) -> MerkleProof:
    return MerkleProof(doc_hash=data["doc_hash"], proof_path=tuple((sib, ProofSide(side)) for sib, side in data["proof_path"]),
                       root_hash=data["root_hash"], leaf_index=data["leaf_index"])
