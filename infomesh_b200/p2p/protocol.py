"""Wire protocol: protocol ids, message type codes, message dataclasses and the length-prefixed msgpack codec.

Byte-compatible with reference infomesh/p2p/protocol.py:28-455 — frame ``[u32 BE length][msgpack {type, payload}]``,
10 MiB cap, bounded unpack limits for untrusted input, DHT key derivation ``/infomesh/{kw,url}/<sha256>``.
"""
from __future__ import annotations

import time
from dataclasses import asdict, dataclass, field
from enum import IntEnum
from typing import Any

import msgpack

from infomesh_b200.hashing import content_hash

PROTOCOL_SEARCH = "/infomesh/search/1.0.0"
PROTOCOL_INDEX = "/infomesh/index/1.0.0"
PROTOCOL_CRAWL = "/infomesh/crawl/1.0.0"
PROTOCOL_REPLICATE = "/infomesh/replicate/1.0.0"
PROTOCOL_PING = "/infomesh/ping/1.0.0"
PROTOCOL_CREDIT = "/infomesh/credit/1.0.0"
PROTOCOL_CREDIT_SYNC = "/infomesh/credit-sync/1.0.0"
PROTOCOL_INDEX_SUBMIT = "/infomesh/index-submit/1.0.0"
PROTOCOL_PEX = "/infomesh/pex/1.0.0"
PROTOCOL_LLM = "/infomesh/llm/1.0.0"
ALL_PROTOCOLS = (PROTOCOL_SEARCH, PROTOCOL_INDEX, PROTOCOL_CRAWL, PROTOCOL_REPLICATE, PROTOCOL_PING, PROTOCOL_CREDIT,
                 PROTOCOL_CREDIT_SYNC, PROTOCOL_INDEX_SUBMIT, PROTOCOL_PEX, PROTOCOL_LLM)


class MessageType(IntEnum):
    PING = 0
    PONG = 1
    SEARCH_REQUEST = 10
    SEARCH_RESPONSE = 11
    INDEX_PUBLISH = 20
    INDEX_PUBLISH_ACK = 21
    INDEX_QUERY = 22
    INDEX_QUERY_RESPONSE = 23
    CRAWL_ASSIGN = 30
    CRAWL_ASSIGN_ACK = 31
    CRAWL_LOCK = 32
    CRAWL_LOCK_ACK = 33
    CRAWL_UNLOCK = 34
    REPLICATE_REQUEST = 40
    REPLICATE_RESPONSE = 41
    ATTESTATION_PUBLISH = 50
    ATTESTATION_PUBLISH_ACK = 51
    KEY_REVOCATION = 60
    KEY_REVOCATION_ACK = 61
    CREDIT_PROOF_REQUEST = 70
    CREDIT_PROOF_RESPONSE = 71
    CREDIT_SYNC_ANNOUNCE = 72
    CREDIT_SYNC_EXCHANGE = 73
    INDEX_SUBMIT = 80
    INDEX_SUBMIT_ACK = 81
    PEX_REQUEST = 90
    PEX_RESPONSE = 91
    ERROR = 99
    SIGNED_ENVELOPE = 100
    # --- built-in transport / Kademlia RPCs (the reference delegates these to py-libp2p's own protocols)
    HELLO = 110
    HELLO_ACK = 111
    DHT_FIND_NODE = 120
    DHT_NODES = 121
    DHT_FIND_VALUE = 122
    DHT_VALUE = 123
    DHT_STORE = 124
    DHT_STORE_ACK = 125
    LLM_REQUEST = 130
    LLM_RESPONSE = 131


def _now() -> float:
    return time.time()


@dataclass(frozen=True)
class PeerPointer:
    """keyword -> where a matching document lives (DHT inverted-index value)."""
    peer_id: str
    doc_id: int
    url: str
    score: float
    title: str = ""


@dataclass(frozen=True)
class SearchRequest:
    query: str
    keywords: list[str]
    limit: int = 10
    request_id: str = ""
    timestamp: float = field(default_factory=_now)


@dataclass(frozen=True)
class SearchResult:
    url: str
    title: str
    snippet: str
    score: float
    peer_id: str = ""
    doc_id: int = 0


@dataclass(frozen=True)
class SearchResponse:
    request_id: str
    results: list[dict[str, Any]]
    peer_id: str = ""
    elapsed_ms: float = 0.0


@dataclass(frozen=True)
class IndexPublish:
    keyword: str
    pointers: list[dict[str, Any]]
    peer_id: str = ""
    timestamp: float = field(default_factory=_now)
    signature: bytes = b""


@dataclass(frozen=True)
class CrawlLock:
    url: str
    url_hash: str = ""
    peer_id: str = ""
    timestamp: float = field(default_factory=_now)
    ttl_seconds: int = 300


@dataclass(frozen=True)
class CrawlAssignment:
    url: str
    depth: int = 0
    priority: float = 1.0
    assigner_peer_id: str = ""


@dataclass(frozen=True)
class ReplicateRequest:
    doc_id: int
    url: str
    title: str
    text: str
    text_hash: str
    language: str = ""
    source_peer_id: str = ""
    replica_index: int = 0


@dataclass(frozen=True)
class Attestation:
    url: str
    raw_hash: str
    text_hash: str
    peer_id: str
    timestamp: float = field(default_factory=_now)
    signature: bytes = b""


@dataclass(frozen=True)
class CreditProofRequest:
    requester_peer_id: str
    request_id: str = ""
    sample_size: int = 10
    timestamp: float = field(default_factory=_now)


@dataclass(frozen=True)
class CreditProofResponse:
    peer_id: str
    request_id: str
    total_earned: float
    total_spent: float
    action_breakdown: dict[str, Any]
    entry_count: int
    merkle_root: str
    root_signature: str
    sample_entries: list[dict[str, Any]]
    sample_proofs: list[dict[str, Any]]
    timestamp: float = field(default_factory=_now)
    public_key: str = ""


@dataclass(frozen=True)
class CreditSyncAnnounce:
    peer_id: str
    owner_email_hash: str
    timestamp: float = field(default_factory=_now)


@dataclass(frozen=True)
class CreditSyncExchange:
    peer_id: str
    owner_email_hash: str
    total_earned: float
    total_spent: float
    contribution_score: float
    entry_count: int
    tier: str
    timestamp: float = field(default_factory=_now)
    signature: str = ""


@dataclass(frozen=True)
class KeyRevocationRecord:
    old_peer_id: str
    new_peer_id: str
    old_public_key: bytes
    new_public_key: bytes
    reason: str = "rotation"
    timestamp: float = field(default_factory=_now)
    old_key_signature: bytes = b""
    new_key_signature: bytes = b""


@dataclass(frozen=True)
class IndexSubmit:
    url: str
    title: str
    text: str
    raw_html_hash: str
    text_hash: str
    language: str = ""
    crawled_at: float = field(default_factory=_now)
    peer_id: str = ""
    signature: bytes = b""
    discovered_links: list[str] = field(default_factory=list)


@dataclass(frozen=True)
class IndexSubmitAck:
    url: str
    doc_id: int = 0
    success: bool = True
    error: str = ""
    peer_id: str = ""


# ------------------------------------------------------------------ codec
MAX_MESSAGE_SIZE = 10 * 1024 * 1024
_PREFIX = 4
_SAFE_UNPACK = dict(max_map_len=2**16, max_array_len=2**16, max_str_len=2**20, max_bin_len=2**20)


def encode_message(msg_type: MessageType | int, payload: dict[str, Any]) -> bytes:
    body = msgpack.packb({"type": int(msg_type), "payload": payload}, use_bin_type=True)
    if len(body) > MAX_MESSAGE_SIZE:
        raise ValueError(f"Message too large: {len(body)} > {MAX_MESSAGE_SIZE}")
    return len(body).to_bytes(_PREFIX, "big") + body


def safe_unpackb(data: bytes) -> Any:
    return msgpack.unpackb(data, raw=False, **_SAFE_UNPACK)


def decode_message(data: bytes) -> tuple[MessageType, dict[str, Any]]:
    """Accepts a framed message (or bare msgpack when the first 4 bytes are not a plausible length)."""
    if len(data) < _PREFIX:
        raise ValueError(f"Message too short: {len(data)} bytes")
    if len(data) > MAX_MESSAGE_SIZE + _PREFIX:
        raise ValueError(f"Message exceeds max size: {len(data)} bytes")
    n = int.from_bytes(data[:_PREFIX], "big")
    try:
        obj = safe_unpackb(data[_PREFIX:_PREFIX + n]) if 0 < n <= MAX_MESSAGE_SIZE else safe_unpackb(data)
    except Exception as exc:  # noqa: BLE001 — msgpack raises several unrelated types
        raise ValueError(f"Malformed message: {exc}") from exc
    if not isinstance(obj, dict) or "type" not in obj or "payload" not in obj:
        raise ValueError("Message missing 'type' or 'payload' field")
    try:
        return MessageType(obj["type"]), obj["payload"]
    except ValueError as exc:
        raise ValueError(f"Unknown message type: {obj['type']!r}") from exc


def read_frame_length(prefix: bytes) -> int:
    """Validate a 4-byte length prefix read from a stream."""
    if len(prefix) != _PREFIX:
        raise ValueError("short length prefix")
    n = int.from_bytes(prefix, "big")
    if n <= 0 or n > MAX_MESSAGE_SIZE:
        raise ValueError(f"invalid frame length {n}")
    return n


def dataclass_to_payload(obj: object) -> dict[str, Any]:
    return asdict(obj)  # type: ignore[call-overload]


def url_to_dht_key(url: str) -> str:
    return f"/infomesh/url/{content_hash(url)}"


def keyword_to_dht_key(keyword: str) -> str:
    return f"/infomesh/kw/{content_hash(keyword.lower())}"


def encode_signed_envelope(envelope_dict: dict[str, Any]) -> bytes:
    return encode_message(MessageType.SIGNED_ENVELOPE, envelope_dict)


def decode_signed_envelope(data: bytes) -> dict[str, Any] | None:
    kind, payload = decode_message(data)
    return payload if kind == MessageType.SIGNED_ENVELOPE else None
