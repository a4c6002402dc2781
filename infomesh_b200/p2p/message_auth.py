"""Signed envelopes for peer messages: Ed25519 over ``peer_id|nonce(8B BE)|timestamp(%.6f)|payload``, strictly
increasing per-sender nonces, 5-minute freshness window, isolated-peer rejection
(reference infomesh/p2p/message_auth.py:44-341)."""
from __future__ import annotations

import time
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Callable

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MAX_MESSAGE_AGE_SECONDS = 300.0
MAX_NONCE_HISTORY = 10_000


class VerificationError(Exception):
    pass


@dataclass(frozen=True)
class SignedEnvelope:
    payload: bytes
    peer_id: str
    signature: bytes
    nonce: int
    timestamp: float = field(default_factory=time.time)


def _canonical_bytes(peer_id: str, nonce: int, timestamp: float, payload: bytes) -> bytes:
    return b"|".join((peer_id.encode(), nonce.to_bytes(8, "big"), f"{timestamp:.6f}".encode(), payload))


class NonceCounter:
    def __init__(self, start: int = 0):
        self._value = start

    def next(self) -> int:
        self._value += 1
        return self._value

    @property
    def current(self) -> int:
        return self._value


def sign_envelope(payload: bytes, key_pair: Any, nonce_counter: NonceCounter, *, now: float | None = None) -> SignedEnvelope:
    ts = now or time.time()
    nonce = nonce_counter.next()
    sig = key_pair.sign(_canonical_bytes(key_pair.peer_id, nonce, ts, payload))
    return SignedEnvelope(payload, key_pair.peer_id, sig, nonce, ts)


def envelope_to_dict(env: SignedEnvelope) -> dict[str, Any]:
    return {"payload": env.payload, "peer_id": env.peer_id, "signature": env.signature, "nonce": env.nonce,
            "timestamp": env.timestamp}


def envelope_from_dict(d: dict[str, Any]) -> SignedEnvelope:
    return SignedEnvelope(d["payload"], d["peer_id"], d["signature"], d["nonce"], d["timestamp"])


class PeerKeyRegistry:
    def __init__(self):
        self._keys: dict[str, bytes] = {}

    def register(self, peer_id: str, public_key: bytes) -> None:
        self._keys[peer_id] = public_key

    def get(self, peer_id: str) -> bytes | None:
        return self._keys.get(peer_id)

    def remove(self, peer_id: str) -> None:
        self._keys.pop(peer_id, None)

    def __contains__(self, peer_id: str) -> bool:
        return peer_id in self._keys

    def __len__(self) -> int:
        return len(self._keys)


class NonceTracker:
    """Highest nonce per sender; LRU-evicts beyond MAX_NONCE_HISTORY senders."""

    def __init__(self):
        self._highest: OrderedDict[str, int] = OrderedDict()

    def check_and_record(self, peer_id: str, nonce: int) -> bool:
        if nonce <= self._highest.get(peer_id, 0):
            return False
        self._highest[peer_id] = nonce
        self._highest.move_to_end(peer_id)
        while len(self._highest) > MAX_NONCE_HISTORY:
            self._highest.popitem(last=False)
        return True

    def highest(self, peer_id: str) -> int:
        return self._highest.get(peer_id, 0)


def _verify_raw(public_key_bytes: bytes, data: bytes, signature: bytes) -> bool:
    try:
        from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PublicKey

        Ed25519PublicKey.from_public_bytes(public_key_bytes).verify(signature, data)
        return True
    except Exception:  # noqa: BLE001 — InvalidSignature / ValueError / missing lib all mean "not verified"
        return False


def verify_envelope(envelope: SignedEnvelope, key_registry: PeerKeyRegistry, nonce_tracker: NonceTracker, *,
                    is_isolated_fn: Callable[[str], bool] | None = None, now: float | None = None,
                    max_age: float = MAX_MESSAGE_AGE_SECONDS) -> bytes:
    """isolation -> known key -> freshness -> signature -> nonce.  The nonce is only recorded for envelopes whose
    signature verified, so a forged envelope cannot burn a victim's nonce range."""
    ts = now or time.time()
    pid = envelope.peer_id
    if is_isolated_fn is not None and is_isolated_fn(pid):
        raise VerificationError(f"peer {pid[:16]} is isolated")
    pub = key_registry.get(pid)
    if pub is None:
        raise VerificationError(f"unknown public key for peer {pid[:16]}")
    age = abs(ts - envelope.timestamp)
    if age > max_age:
        raise VerificationError(f"message too old ({age:.0f}s > {max_age:.0f}s)")
    if envelope.nonce <= nonce_tracker.highest(pid):
        raise VerificationError(f"replayed nonce {envelope.nonce} from {pid[:16]}")
    if not _verify_raw(pub, _canonical_bytes(pid, envelope.nonce, envelope.timestamp, envelope.payload),
                       envelope.signature):
        raise VerificationError(f"invalid signature from {pid[:16]}")
    nonce_tracker.check_and_record(pid, envelope.nonce)
    return envelope.payload
