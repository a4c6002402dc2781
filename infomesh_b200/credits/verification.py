"""Peer-verifiable credit proofs: Merkle tree over the signed ledger entries, signed root, random sample of entries
with membership proofs; the verifier checks root signature, entry hashes + signatures, and proofs
(reference infomesh/credits/verification.py:58-390)."""
from __future__ import annotations

import random
import time
from dataclasses import dataclass
from typing import Any

from infomesh_b200.credits.ledger import CreditLedger, _entry_canonical
from infomesh_b200.credits.types import CreditEntry
from infomesh_b200.hashing import content_hash
from infomesh_b200.p2p.keys import verify_with_public_key
from infomesh_b200.trust.merkle import MerkleTree, deserialize_proof, serialize_proof
from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass(frozen=True)
class CreditVerificationResult:
    peer_id: str
    verified: bool
    total_earned: float
    entry_count: int
    valid_signatures: int
    invalid_signatures: int
    valid_proofs: int
    invalid_proofs: int
    merkle_root_valid: bool
    detail: str


def _root_canonical(root_hash: str, entry_count: int, peer_id: str) -> bytes:
    return f"{root_hash}|{entry_count}|{peer_id}".encode()


def _select_sample(total: int, sample_size: int) -> list[int]:
    return list(range(total)) if total <= sample_size else sorted(random.sample(range(total), sample_size))


_ENTRY_FIELDS = ("entry_hash", "action", "quantity", "weight", "multiplier", "credits", "timestamp", "note",
                 "signature")


def _entry_to_dict(e: CreditEntry) -> dict[str, Any]:
    return {k: getattr(e, k) for k in _ENTRY_FIELDS}


class CreditProofBuilder:
    def __init__(self, ledger: CreditLedger, key_pair: KeyPairLike):
        self._ledger = ledger
        self._kp = key_pair

    def build_proof(self, *, sample_size: int = 10, request_id: str = "") -> dict[str, Any]:
        entries = self._ledger.signed_entries()
        base = {"peer_id": self._kp.peer_id, "request_id": request_id, "timestamp": time.time(),
                "public_key": self._kp.public_key_bytes().hex()}
        if not entries:
            return {**base, "total_earned": 0.0, "total_spent": 0.0, "action_breakdown": {}, "entry_count": 0,
                    "merkle_root": "", "root_signature": "", "sample_entries": [], "sample_proofs": []}
        tree = MerkleTree()
        tree.build([e.entry_hash for e in entries])
        picks = _select_sample(len(entries), sample_size)
        breakdown: dict[str, float] = {}
        for e in entries:
            breakdown[e.action] = breakdown.get(e.action, 0.0) + e.credits
        st = self._ledger.stats()
        return {**base, "total_earned": st.total_earned, "total_spent": st.total_spent, "action_breakdown": breakdown,
                "entry_count": len(entries), "merkle_root": tree.root_hash,
                "root_signature": self._kp.sign(_root_canonical(tree.root_hash, len(entries), self._kp.peer_id)).hex(),
                "sample_entries": [_entry_to_dict(entries[i]) for i in picks],
                "sample_proofs": [serialize_proof(tree.get_proof(i)) for i in picks]}

    @staticmethod
    def verify_proof(proof_data: dict[str, Any], *, known_public_key: bytes | None = None) -> CreditVerificationResult:
        peer = proof_data.get("peer_id", "")
        n = proof_data.get("entry_count", 0)
        earned = proof_data.get("total_earned", 0.0)

        def result(ok, vs=0, bad_s=0, vp=0, bad_p=0, root_ok=False, detail="ok"):
            return CreditVerificationResult(peer, ok, earned if n else 0.0, n, vs, bad_s, vp, bad_p, root_ok, detail)

        if n == 0:
            return result(True, root_ok=True, detail="empty_ledger")
        try:
            pub = known_public_key if known_public_key is not None else bytes.fromhex(proof_data["public_key"])
            if len(pub) != 32:
                raise ValueError("public key must be 32 bytes")
        except Exception as exc:  # noqa: BLE001
            return result(False, detail=f"invalid_public_key: {exc}")
        root = proof_data.get("merkle_root", "")
        try:
            root_ok = verify_with_public_key(pub, _root_canonical(root, n, peer),
                                             bytes.fromhex(proof_data.get("root_signature", "")))
        except ValueError:
            root_ok = False
        samples, proofs = proof_data.get("sample_entries", []), proof_data.get("sample_proofs", [])
        vs = bad_s = vp = bad_p = 0
        for i, ent in enumerate(samples):
            try:
                canon = _entry_canonical(ent["action"], ent["quantity"], ent["weight"], ent["multiplier"],
                                         ent["credits"], ent["timestamp"], ent["note"])
                if content_hash(canon) != ent["entry_hash"]:
                    bad_s += 1
                    continue
                if verify_with_public_key(pub, canon, bytes.fromhex(ent["signature"])):
                    vs += 1
                else:
                    bad_s += 1
            except (KeyError, ValueError, TypeError):
                bad_s += 1
                continue
            if i < len(proofs):
                try:
                    pr = deserialize_proof(proofs[i])
                    good = (MerkleTree.verify_proof(pr) and pr.root_hash == root
                            and MerkleTree.verify_document(ent["entry_hash"], pr))
                except (KeyError, ValueError, TypeError):
                    good = False
                vp, bad_p = vp + good, bad_p + (not good)
        ok = root_ok and bad_s == 0 and bad_p == 0 and (vs > 0 or not samples)
        notes = ([] if root_ok else ["merkle_root_signature_invalid"]) + \
                ([f"invalid_entry_signatures={bad_s}"] if bad_s else []) + \
                ([f"invalid_merkle_proofs={bad_p}"] if bad_p else [])
        return result(ok, vs, bad_s, vp, bad_p, root_ok, "; ".join(notes) or "ok")
