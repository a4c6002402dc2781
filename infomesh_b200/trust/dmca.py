"""DMCA takedown notices: signed by the requester, 24 h compliance deadline, acknowledgements / propagation
tracked per peer, persisted in SQLite so a restart cannot shed obligations (reference infomesh/trust/dmca.py:36-576)."""
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

COMPLIANCE_DEADLINE_HOURS: float = 24.0
TAKEDOWN_DHT_PREFIX: str = "/infomesh/takedown/"
MAX_NOTICE_LENGTH: int = 10_000


class TakedownStatus(StrEnum):
    PENDING = "pending"
    ACKNOWLEDGED = "acknowledged"
    COMPLIED = "complied"
    EXPIRED = "expired"
    INVALID = "invalid"


@dataclass(frozen=True)
class TakedownNotice:
    notice_id: str
    url: str
    requester_id: str
    reason: str
    signature: bytes
    created_at: float
    deadline: float
    contact_info: str = ""


@dataclass(frozen=True)
class TakedownAck:
    notice_id: str
    peer_id: str
    status: TakedownStatus
    complied_at: float | None = None
    detail: str = ""


@dataclass
class TakedownRecord:
    notice: TakedownNotice
    acknowledgments: list[TakedownAck] = field(default_factory=list)
    propagated_to: list[str] = field(default_factory=list)


def takedown_dht_key(url: str) -> str:
    return f"{TAKEDOWN_DHT_PREFIX}{content_hash(url)}"


def _generate_notice_id(url: str, peer_id: str, timestamp: float) -> str:
    return short_hash(f"takedown|{url}|{peer_id}|{timestamp}".encode(), length=24)


def _notice_payload(notice_id: str, url: str, reason: str, created_at: float) -> bytes:
    return f"{notice_id}|{url}|{reason}|{created_at}".encode()


def serialize_notice(notice: TakedownNotice) -> dict[str, Any]:
    return {"notice_id": notice.notice_id, "url": notice.url, "requester_id": notice.requester_id, "reason": notice.reason,
            "signature": notice.signature.hex(), "created_at": notice.created_at, "deadline": notice.deadline,
            "contact_info": notice.contact_info}


def deserialize_notice(data: dict[str, Any]) -> TakedownNotice:
    return TakedownNotice(data["notice_id"], data["url"], data["requester_id"], data["reason"], bytes.fromhex(data["signature"]),
                          data["created_at"], data["deadline"], data.get("contact_info", ""))


class _TakedownStore(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS takedown_notices (notice_id TEXT PRIMARY KEY, body TEXT NOT NULL);
        CREATE TABLE IF NOT EXISTS takedown_acks (id INTEGER PRIMARY KEY AUTOINCREMENT, notice_id TEXT NOT NULL,
            peer_id TEXT NOT NULL, status TEXT NOT NULL, complied_at REAL, detail TEXT NOT NULL DEFAULT '');
        CREATE TABLE IF NOT EXISTS takedown_propagations (notice_id TEXT NOT NULL, peer_id TEXT NOT NULL,
            PRIMARY KEY (notice_id, peer_id));
    """

    def save_notice(self, n: TakedownNotice) -> None:
        with self._lock:
            self._conn.execute("INSERT OR REPLACE INTO takedown_notices VALUES (?, ?)",
                               (n.notice_id, json.dumps(serialize_notice(n))))
            self._conn.commit()

    def save_ack(self, a: TakedownAck) -> None:
        with self._lock:
            self._conn.execute("INSERT INTO takedown_acks (notice_id, peer_id, status, complied_at, detail) "
                               "VALUES (?, ?, ?, ?, ?)", (a.notice_id, a.peer_id, a.status.value, a.complied_at, a.detail))
            self._conn.commit()

    def save_propagation(self, notice_id: str, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO takedown_propagations VALUES (?, ?)", (notice_id, peer_id))
            self._conn.commit()

    def load_all(self) -> list[TakedownRecord]:
        out = []
        for nid, body in self._conn.execute("SELECT notice_id, body FROM takedown_notices"):
            acks = [TakedownAck(nid, r[0], TakedownStatus(r[1]), r[2], r[3]) for r in self._conn.execute(
                "SELECT peer_id, status, complied_at, detail FROM takedown_acks WHERE notice_id = ? ORDER BY id", (nid,))]
            props = [r[0] for r in self._conn.execute("SELECT peer_id FROM takedown_propagations WHERE notice_id = ?", (nid,))]
            out.append(TakedownRecord(deserialize_notice(json.loads(body)), acks, props))
        return out


class TakedownManager:
    MAX_NOTICES_PER_HOUR: int = 10

    def __init__(self, db_path: str | None = None):
        self._records: dict[str, TakedownRecord] = {}
        self._by_url: dict[str, str] = {}
        self._rate: dict[str, list[float]] = {}
        self._store = _TakedownStore(db_path) if db_path is not None else None
        if self._store is not None:
            for rec in self._store.load_all():
                self._records[rec.notice.notice_id] = rec
                self._by_url[rec.notice.url] = rec.notice.notice_id

    def _rate_ok(self, requester: str, now: float) -> bool:
        recent = [t for t in self._rate.get(requester, []) if t > now - 3600]
        self._rate[requester] = recent
        return len(recent) < self.MAX_NOTICES_PER_HOUR

    def create_notice(self, url: str, reason: str, key_pair: KeyPairLike, *, contact_info: str = "",
                      now: float | None = None) -> TakedownNotice:
        now = now or time.time()
        if not self._rate_ok(key_pair.peer_id, now):
            raise ValueError(f"Rate limit exceeded: max {self.MAX_NOTICES_PER_HOUR} takedown notices per hour")
        reason = reason[:MAX_NOTICE_LENGTH]
        nid = _generate_notice_id(url, key_pair.peer_id, now)
        notice = TakedownNotice(nid, url, key_pair.peer_id, reason, key_pair.sign(_notice_payload(nid, url, reason, now)),
                                now, now + COMPLIANCE_DEADLINE_HOURS * 3600, contact_info)
        self._register(notice)
        self._rate.setdefault(key_pair.peer_id, []).append(now)
        return notice

    def _register(self, notice: TakedownNotice) -> None:
        self._records.setdefault(notice.notice_id, TakedownRecord(notice))
        self._by_url[notice.url] = notice.notice_id
        if self._store:
            self._store.save_notice(notice)

    def verify_notice(self, notice: TakedownNotice, key_pair: KeyPairLike) -> bool:
        return key_pair.verify(_notice_payload(notice.notice_id, notice.url, notice.reason, notice.created_at),
                               notice.signature)

    def receive_notice(self, notice: TakedownNotice, requester_key: KeyPairLike | None) -> bool:
        """Accept a notice propagated by a peer only when its signature checks out."""
        if requester_key is None or not self.verify_notice(notice, requester_key):
            logger.warning("takedown_rejected", notice_id=notice.notice_id)
            return False
        self._register(notice)
        return True

    def acknowledge(self, notice_id: str, peer_id: str, *, status: TakedownStatus = TakedownStatus.ACKNOWLEDGED,
                    now: float | None = None) -> TakedownAck | None:
        rec = self._records.get(notice_id)
        if rec is None:
            return None
        now = now or time.time()
        ack = TakedownAck(notice_id, peer_id, status, now if status == TakedownStatus.COMPLIED else None,
                          f"acknowledged at {now:.0f}")
        rec.acknowledgments.append(ack)
        if self._store:
            self._store.save_ack(ack)
        return ack

    def mark_complied(self, notice_id: str, peer_id: str, *, now: float | None = None) -> TakedownAck | None:
        return self.acknowledge(notice_id, peer_id, status=TakedownStatus.COMPLIED, now=now)

    def record_propagation(self, notice_id: str, peer_id: str) -> None:
        rec = self._records.get(notice_id)
        if rec and peer_id not in rec.propagated_to:
            rec.propagated_to.append(peer_id)
            if self._store:
                self._store.save_propagation(notice_id, peer_id)

    def is_taken_down(self, url: str) -> bool:
        return url in self._by_url

    def get_notice_for_url(self, url: str) -> TakedownNotice | None:
        rec = self._records.get(self._by_url.get(url, ""))
        return rec.notice if rec else None

    def get_record(self, notice_id: str) -> TakedownRecord | None:
        return self._records.get(notice_id)

    def _peer_status(self, rec: TakedownRecord, peer_id: str) -> TakedownStatus | None:
        states = [a.status for a in rec.acknowledgments if a.peer_id == peer_id]
        if TakedownStatus.COMPLIED in states:
            return TakedownStatus.COMPLIED
        return states[-1] if states else None

    def check_compliance(self, notice_id: str, peer_id: str, *, now: float | None = None) -> TakedownStatus:
        rec = self._records.get(notice_id)
        if rec is None:
            return TakedownStatus.INVALID
        st = self._peer_status(rec, peer_id)
        if st is not None:
            return st
        return TakedownStatus.EXPIRED if (now or time.time()) > rec.notice.deadline else TakedownStatus.PENDING

    def list_active(self) -> list[TakedownNotice]:
        return [r.notice for r in self._records.values()]

    def list_non_compliant(self, peer_id: str, *, now: float | None = None) -> list[TakedownNotice]:
        return [r.notice for r in self._records.values()
                if self._peer_status(r, peer_id) not in (TakedownStatus.COMPLIED, TakedownStatus.INVALID)]

    def close(self) -> None:
        if self._store:
            self._store.close()
