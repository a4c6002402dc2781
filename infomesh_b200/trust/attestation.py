"""Signed content attestations: ``sign(url|raw_hash|text_hash|crawled_at)`` proves which peer crawled what
(reference infomesh/trust/attestation.py:25-259)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any

from infomesh_b200.hashing import content_hash
from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass(frozen=True)
class ContentAttestation:
    url: str
    raw_hash: str
    text_hash: str
    peer_id: str
    signature: bytes
    crawled_at: float
    content_length: int


@dataclass(frozen=True)
class VerificationResult:
    url: str
    raw_match: bool
    text_match: bool
    signature_valid: bool
    verified: bool
    detail: str


def _attestation_payload(url: str, raw_hash: str, text_hash: str, crawled_at: float) -> bytes:
    return f"{url}|{raw_hash}|{text_hash}|{crawled_at}".encode()


def create_attestation(url: str, raw_body: bytes, extracted_text: str, key_pair: KeyPairLike, *,
                       crawled_at: float | None = None) -> ContentAttestation:
    raw_h, text_h, ts = content_hash(raw_body), content_hash(extracted_text), crawled_at or time.time()
    return ContentAttestation(url, raw_h, text_h, key_pair.peer_id,
                              key_pair.sign(_attestation_payload(url, raw_h, text_h, ts)), ts,
                              len(extracted_text.encode("utf-8")))


def verify_attestation(attestation: ContentAttestation, key_pair: KeyPairLike, *, raw_body: bytes | None = None,
                       extracted_text: str | None = None) -> VerificationResult:
    """Signature always; hashes only for the material the verifier re-obtained."""
    a = attestation
    sig_ok = key_pair.verify(_attestation_payload(a.url, a.raw_hash, a.text_hash, a.crawled_at), a.signature)
    raw_ok = raw_body is None or content_hash(raw_body) == a.raw_hash
    text_ok = extracted_text is None or content_hash(extracted_text) == a.text_hash
    notes = [n for n, ok in (("signature_invalid", sig_ok), ("raw_hash_mismatch", raw_ok),
                             ("text_hash_mismatch", text_ok)) if not ok]
    return VerificationResult(a.url, raw_ok, text_ok, sig_ok, not notes, "; ".join(notes) or "ok")


def verify_attestation_with_key(attestation: ContentAttestation, public_key: bytes) -> bool:
    from infomesh_b200.p2p.keys import verify_with_public_key

    a = attestation
    return verify_with_public_key(public_key, _attestation_payload(a.url, a.raw_hash, a.text_hash, a.crawled_at),
                                  a.signature)


def serialize_attestation(att: ContentAttestation) -> dict[str, Any]:
    return {"url": att.url, "raw_hash": att.raw_hash, "text_hash": att.text_hash, "peer_id": att.peer_id,
            "signature": att.signature.hex(), "crawled_at": att.crawled_at, "content_length": att.content_length}


def deserialize_attestation(data: dict[str, Any]) -> ContentAttestation:
    return ContentAttestation(data["url"], data["raw_hash"], data["text_hash"], data["peer_id"],
                              bytes.fromhex(data["signature"]), data["crawled_at"], data["content_length"])


def verify_merkle_root(root: Any, key_pair: KeyPairLike) -> bool:
    """Check a ``trust.merkle.MerkleRoot`` signature with the publisher's key."""
    if not getattr(root, "signature", b""):
        return False
    payload = f"{root.root_hash}|{root.document_count}|{root.built_at}|{root.peer_id}".encode()
    return key_pair.verify(payload, root.signature)
