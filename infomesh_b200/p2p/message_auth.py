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


MAX_KNOWN_KEYS = 10_000          # LRU bound of the trust-on-first-use key table
NONCE_WINDOW = 1024              # how far behind the highest nonce an unseen nonce is still accepted


class PeerKeyRegistry:
    """peer id -> Ed25519 public key, least-recently-used eviction past ``capacity`` so a flood of fresh identities cannot
    grow the table without bound (pinned entries -- operator-configured peers -- are never evicted)."""

    def __init__(self, capacity: int = MAX_KNOWN_KEYS):
        self._keys: OrderedDict[str, bytes] = OrderedDict()
        self._pinned: set[str] = set()
        self._cap = max(1, int(capacity))

    def register(self, peer_id: str, public_key: bytes, *, pinned: bool = False) -> None:
        self._keys[peer_id] = public_key
        self._keys.move_to_end(peer_id)
        if pinned:
            self._pinned.add(peer_id)
        if len(self._keys) > self._cap:
            for victim in list(self._keys):
                if len(self._keys) <= self._cap:
                    break
                if victim not in self._pinned and victim != peer_id:
                    del self._keys[victim]

    def get(self, peer_id: str) -> bytes | None:
        pub = self._keys.get(peer_id)
        if pub is not None:
            self._keys.move_to_end(peer_id)
        return pub

    def remove(self, peer_id: str) -> None:
        self._keys.pop(peer_id, None)
        self._pinned.discard(peer_id)

    def __contains__(self, peer_id: str) -> bool:
        return peer_id in self._keys

    def __len__(self) -> int:
        return len(self._keys)


class NonceTracker:
    """Replay filter per sender: a sliding window below the highest nonce seen.

    A strictly increasing rule rejects honest traffic when frames of one sender arrive out of order (each request uses
    its own TCP connection here, so a large REPLICATE can land after a later PING).  A nonce is accepted iff it has not
    been seen and is not older than ``NONCE_WINDOW`` below the sender's highest; senders are LRU-evicted beyond
    MAX_NONCE_HISTORY."""

    def __init__(self, window: int = NONCE_WINDOW):
        self._highest: OrderedDict[str, int] = OrderedDict()
        self._seen: dict[str, set[int]] = {}
        self._window = max(0, int(window))

    def acceptable(self, peer_id: str, nonce: int) -> bool:
        hi = self._highest.get(peer_id, 0)
        if nonce > hi:
            return True
        return nonce > hi - self._window and nonce > 0 and nonce not in self._seen.get(peer_id, ())

    def check_and_record(self, peer_id: str, nonce: int) -> bool:
        if not self.acceptable(peer_id, nonce):
            return False
        hi = max(self._highest.get(peer_id, 0), nonce)
        self._highest[peer_id] = hi
        self._highest.move_to_end(peer_id)
        seen = self._seen.setdefault(peer_id, set())
        seen.add(nonce)
        if len(seen) > 2 * self._window + 2:
            floor = hi - self._window
            self._seen[peer_id] = {n for n in seen if n > floor}
        while len(self._highest) > MAX_NONCE_HISTORY:
            old, _ = self._highest.popitem(last=False)
            self._seen.pop(old, None)
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
    if not nonce_tracker.acceptable(pid, envelope.nonce):
        raise VerificationError(f"replayed nonce {envelope.nonce} from {pid[:16]}")
    if not _verify_raw(pub, _canonical_bytes(pid, envelope.nonce, envelope.timestamp, envelope.payload),
                       envelope.signature):
        raise VerificationError(f"invalid signature from {pid[:16]}")
    nonce_tracker.check_and_record(pid, envelope.nonce)
    return envelope.payload
