"""GDPR Art. 17 deletion requests: signed, persisted, with a permanent re-crawl blocklist
(reference infomesh/trust/gdpr.py:39-644).  Unsigned / unverifiable requests are never actioned."""
from __future__ import annotations

import json
import time
from dataclasses import dataclass, field
from enum import StrEnum
from typing import Any

from infomesh_b200.db import SQLiteStore
from infomesh_b200.hashing import content_hash, short_hash
from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DELETION_DHT_PREFIX: str = "/infomesh/gdpr/"
MAX_REASON_LENGTH: int = 5_000


class DeletionStatus(StrEnum):
    PENDING = "pending"
    ACKNOWLEDGED = "acknowledged"
    DELETED = "deleted"
    INVALID = "invalid"


class DeletionBasis(StrEnum):
    RIGHT_TO_ERASURE = "right_to_erasure"
    CONSENT_WITHDRAWN = "consent_withdrawn"
    OBJECTION = "objection"
    UNLAWFUL_PROCESSING = "unlawful_processing"
    LEGAL_OBLIGATION = "legal_obligation"


@dataclass(frozen=True)
class DeletionRequest:
    request_id: str
    url: str
    requester_id: str
    basis: DeletionBasis
    reason: str
    signature: bytes
    created_at: float
    personal_data_fields: list[str] = field(default_factory=list)


@dataclass(frozen=True)
class DeletionConfirmation:
    request_id: str
    peer_id: str
    status: DeletionStatus
    deleted_at: float | None = None
    detail: str = ""


@dataclass
class DeletionRecord:
    request: DeletionRequest
    confirmations: list[DeletionConfirmation] = field(default_factory=list)
    propagated_to: list[str] = field(default_factory=list)


def deletion_dht_key(url: str) -> str:
    return f"{DELETION_DHT_PREFIX}{content_hash(url)}"


def _generate_request_id(url: str, peer_id: str, timestamp: float) -> str:
    return short_hash(f"gdpr|{url}|{peer_id}|{timestamp}".encode(), length=24)


def _request_payload(request_id: str, url: str, basis: str, reason: str, created_at: float) -> bytes:
    return f"{request_id}|{url}|{basis}|{reason}|{created_at}".encode()


def serialize_request(request: DeletionRequest) -> dict[str, Any]:
    return {"request_id": request.request_id, "url": request.url, "requester_id": request.requester_id, "basis": request.basis.value,
            "reason": request.reason, "signature": request.signature.hex(), "created_at": request.created_at,
            "personal_data_fields": list(request.personal_data_fields)}


def deserialize_request(data: dict[str, Any]) -> DeletionRequest:
    return DeletionRequest(data["request_id"], data["url"], data["requester_id"], DeletionBasis(data["basis"]), data["reason"],
                           bytes.fromhex(data["signature"]), data["created_at"], list(data.get("personal_data_fields", [])))


class _GDPRStore(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS gdpr_requests (request_id TEXT PRIMARY KEY, body TEXT NOT NULL);
        CREATE TABLE IF NOT EXISTS gdpr_confirmations (id INTEGER PRIMARY KEY AUTOINCREMENT, request_id TEXT NOT NULL,
            peer_id TEXT NOT NULL, status TEXT NOT NULL, deleted_at REAL, detail TEXT NOT NULL DEFAULT '');
        CREATE TABLE IF NOT EXISTS gdpr_propagations (request_id TEXT NOT NULL, peer_id TEXT NOT NULL,
            PRIMARY KEY (request_id, peer_id));
        CREATE TABLE IF NOT EXISTS gdpr_blocklist (url TEXT PRIMARY KEY, added_at REAL NOT NULL);
    """

    def run(self, sql: str, args: tuple = ()) -> None:
        with self._lock:
            self._conn.execute(sql, args)
            self._conn.commit()

    def load_all(self) -> list[DeletionRecord]:
        out = []
        for rid, body in self._conn.execute("SELECT request_id, body FROM gdpr_requests"):
            confs = [DeletionConfirmation(rid, r[0], DeletionStatus(r[1]), r[2], r[3]) for r in self._conn.execute(
                "SELECT peer_id, status, deleted_at, detail FROM gdpr_confirmations WHERE request_id = ? ORDER BY id", (rid,))]
            props = [r[0] for r in self._conn.execute("SELECT peer_id FROM gdpr_propagations WHERE request_id = ?", (rid,))]
            out.append(DeletionRecord(deserialize_request(json.loads(body)), confs, props))
        return out

    def load_blocklist(self) -> list[str]:
        return [r[0] for r in self._conn.execute("SELECT url FROM gdpr_blocklist")]


class DeletionManager:
    def __init__(self, db_path: str | None = None):
        self._records: dict[str, DeletionRecord] = {}
        self._by_url: dict[str, str] = {}
        self._blocklist: set[str] = set()
        self._store = _GDPRStore(db_path) if db_path is not None else None
        if self._store is not None:
            for rec in self._store.load_all():
                self._records[rec.request.request_id] = rec
                self._by_url[rec.request.url] = rec.request.request_id
            self._blocklist.update(self._store.load_blocklist())

    def _register(self, req: DeletionRequest) -> None:
        if req.request_id not in self._records:
            self._records[req.request_id] = DeletionRecord(req)
            if self._store:
                self._store.run("INSERT OR REPLACE INTO gdpr_requests VALUES (?, ?)",
                                (req.request_id, json.dumps(serialize_request(req))))
        self._by_url[req.url] = req.request_id
        self._blocklist.add(req.url)
        if self._store:
            self._store.run("INSERT OR IGNORE INTO gdpr_blocklist VALUES (?, ?)", (req.url, time.time()))

    def create_request(self, url: str, basis: DeletionBasis, reason: str, key_pair: KeyPairLike, *,
                       personal_data_fields: list[str] | None = None, now: float | None = None) -> DeletionRequest:
        now = now or time.time()
        reason = reason[:MAX_REASON_LENGTH]
        rid = _generate_request_id(url, key_pair.peer_id, now)
        req = DeletionRequest(rid, url, key_pair.peer_id, basis, reason,
                              key_pair.sign(_request_payload(rid, url, basis.value, reason, now)), now,
                              personal_data_fields or [])
        self._register(req)
        return req

    def verify_request(self, request: DeletionRequest, key_pair: KeyPairLike) -> bool:
        r = request
        return key_pair.verify(_request_payload(r.request_id, r.url, r.basis.value, r.reason, r.created_at), r.signature)

    def receive_request(self, request: DeletionRequest, requester_key: KeyPairLike | None = None) -> bool:
        if requester_key is None or not self.verify_request(request, requester_key):
            logger.warning("gdpr_request_rejected", request_id=request.request_id,
                           reason="no_key_provided" if requester_key is None else "invalid_signature")
            return False
        self._register(request)
        return True

    def confirm_deletion(self, request_id: str, peer_id: str, *, now: float | None = None) -> DeletionConfirmation | None:
        rec = self._records.get(request_id)
        if rec is None:
            return None
        now = now or time.time()
        conf = DeletionConfirmation(request_id, peer_id, DeletionStatus.DELETED, now, f"deleted at {now:.0f}")
        rec.confirmations.append(conf)
        if self._store:
            self._store.run("INSERT INTO gdpr_confirmations (request_id, peer_id, status, deleted_at, detail) "
                            "VALUES (?, ?, ?, ?, ?)", (request_id, peer_id, conf.status.value, now, conf.detail))
        return conf

    def record_propagation(self, request_id: str, peer_id: str) -> None:
        rec = self._records.get(request_id)
        if rec and peer_id not in rec.propagated_to:
            rec.propagated_to.append(peer_id)
            if self._store:
                self._store.run("INSERT OR IGNORE INTO gdpr_propagations VALUES (?, ?)", (request_id, peer_id))

    def is_blocked(self, url: str) -> bool:
        return url in self._blocklist

    def unblock(self, url: str, *, admin_key: KeyPairLike) -> bool:
        """Reverse an erroneous request (court order, mistake); the record is kept, marked INVALID."""
        if url not in self._blocklist:
            return False
        self._blocklist.discard(url)
        if self._store:
            self._store.run("DELETE FROM gdpr_blocklist WHERE url = ?", (url,))
        rid = self._by_url.pop(url, None)
        rec = self._records.get(rid or "")
        if rec is not None:
            rec.confirmations.append(DeletionConfirmation(rid, getattr(admin_key, "peer_id", "admin"),
                                                          DeletionStatus.INVALID, time.time(), "admin_unblock"))
        return True

    def get_request_for_url(self, url: str) -> DeletionRequest | None:
        rec = self._records.get(self._by_url.get(url, ""))
        return rec.request if rec else None

    def get_record(self, request_id: str) -> DeletionRecord | None:
        return self._records.get(request_id)

    def list_pending(self, peer_id: str) -> list[DeletionRequest]:
        return [r.request for r in self._records.values()
                if not any(c.peer_id == peer_id and c.status == DeletionStatus.DELETED for c in r.confirmations)]

    def list_all(self) -> list[DeletionRequest]:
        return [r.request for r in self._records.values()]

    @property
    def blocklist_size(self) -> int:
        return len(self._blocklist)

    def close(self) -> None:
        if self._store:
            self._store.close()
