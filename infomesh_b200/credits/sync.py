"""Cross-node credit aggregation for nodes of the same owner (matched by SHA-256 of the normalised e-mail):
signed per-node summaries are exchanged and summed (reference infomesh/credits/sync.py:47-585).
Summaries expire after 72 h, at most 20 peers per owner, resync every 300 s."""
from __future__ import annotations

import time
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Any

from infomesh_b200.credits.ledger import CreditLedger
from infomesh_b200.db import SQLiteStore
from infomesh_b200.hashing import content_hash
from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

SUMMARY_TTL_HOURS: float = 72.0
SYNC_INTERVAL_SECONDS: float = 300.0
MAX_PEER_SUMMARIES: int = 20
MAX_CLOCK_SKEW_SECONDS = 300.0


def _num(d: dict[str, Any], key: str, cast=float):
    v = d.get(key)
    return cast(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else cast(0)


@dataclass(frozen=True)
class CreditSummary:
    peer_id: str
    owner_email_hash: str
    total_earned: float
    total_spent: float
    contribution_score: float
    entry_count: int
    tier: str
    timestamp: float
    signature: str = ""
    public_key: str = ""      # hex; lets a receiver check the signature (peer_id must derive from it)

    def to_dict(self) -> dict[str, object]:
        return asdict(self)

    @classmethod
    def from_dict(cls, d: dict[str, Any]) -> "CreditSummary":
        return cls(str(d.get("peer_id", "")), str(d.get("owner_email_hash", "")), _num(d, "total_earned"),
                   _num(d, "total_spent"), _num(d, "contribution_score"), _num(d, "entry_count", int),
                   str(d.get("tier", "Tier 1")), _num(d, "timestamp"), str(d.get("signature", "")),
                   str(d.get("public_key", "")))

    def canonical(self) -> bytes:
        return (f"{self.peer_id}|{self.owner_email_hash}|{self.total_earned}|{self.total_spent}|"
                f"{self.contribution_score}|{self.timestamp}").encode()


@dataclass
class AggregatedCreditStats:
    total_earned: float = 0.0
    total_spent: float = 0.0
    balance: float = 0.0
    contribution_score: float = 0.0
    node_count: int = 1
    peer_summaries: list[CreditSummary] = field(default_factory=list)


class CreditSyncStore(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS peer_credit_summaries (
            peer_id TEXT PRIMARY KEY, owner_email_hash TEXT NOT NULL, total_earned REAL NOT NULL,
            total_spent REAL NOT NULL, contribution_score REAL NOT NULL, entry_count INTEGER NOT NULL DEFAULT 0,
            tier TEXT NOT NULL DEFAULT '', timestamp REAL NOT NULL, signature TEXT NOT NULL DEFAULT '',
            received_at REAL NOT NULL);
        CREATE INDEX IF NOT EXISTS idx_pcs_owner ON peer_credit_summaries(owner_email_hash);
    """

    def __init__(self, db_path: Path | str | None = None, **store_options):
        super().__init__(db_path, **store_options)        # check_same_thread / row_factory / extra_pragmas of SQLiteStore

    def store_summary(self, summary: CreditSummary) -> None:
        with self._lock:
            self._conn.execute(
                "INSERT OR REPLACE INTO peer_credit_summaries (peer_id, owner_email_hash, total_earned, total_spent, "
                "contribution_score, entry_count, tier, timestamp, signature, received_at) "
                "VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                (summary.peer_id, summary.owner_email_hash, summary.total_earned, summary.total_spent, summary.contribution_score, summary.entry_count,
                 summary.tier, summary.timestamp, summary.signature, time.time()))
            self._conn.commit()

    def get_peer_summaries(self, owner_email_hash: str, *, include_stale: bool = False) -> list[CreditSummary]:
        cutoff = 0.0 if include_stale else time.time() - SUMMARY_TTL_HOURS * 3600
        rows = self._conn.execute(
            "SELECT peer_id, owner_email_hash, total_earned, total_spent, contribution_score, entry_count, tier, "
            "timestamp, signature FROM peer_credit_summaries WHERE owner_email_hash = ? AND received_at >= ? "
            "ORDER BY timestamp DESC", (owner_email_hash, cutoff)).fetchall()
        return [CreditSummary(*tuple(r)) for r in rows]

    def purge_stale(self) -> int:
        with self._lock:
            cur = self._conn.execute("DELETE FROM peer_credit_summaries WHERE received_at < ?",
                                     (time.time() - SUMMARY_TTL_HOURS * 3600,))
            self._conn.commit()
        return cur.rowcount

    def remove_peer(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("DELETE FROM peer_credit_summaries WHERE peer_id = ?", (peer_id,))
            self._conn.commit()

    def peer_count(self, owner_email_hash: str) -> int:
        return int(self._conn.execute("SELECT COUNT(*) FROM peer_credit_summaries WHERE owner_email_hash = ?",
                                      (owner_email_hash,)).fetchone()[0])

    def has_peer(self, peer_id: str) -> bool:
        return self._conn.execute("SELECT 1 FROM peer_credit_summaries WHERE peer_id = ?",
                                  (peer_id,)).fetchone() is not None


class CreditSyncManager:
    def __init__(self, ledger: CreditLedger, store: CreditSyncStore, owner_email: str,
                 key_pair: KeyPairLike | None = None, local_peer_id: str = ""):
        self._ledger, self._store, self._kp, self._peer_id = ledger, store, key_pair, local_peer_id
        self._owner_hash = content_hash(owner_email.lower().strip()) if owner_email else ""
        self._peers: dict[str, float] = {}

    @property
    def owner_email_hash(self) -> str:
        return self._owner_hash

    @property
    def has_identity(self) -> bool:
        return bool(self._owner_hash)

    def build_summary(self) -> CreditSummary:
        st = self._ledger.stats()
        draft = CreditSummary(self._peer_id, self._owner_hash, st.total_earned, st.total_spent,
                              st.contribution_score, int(st.total_earned + st.total_spent),
                              getattr(st.tier, "value", str(st.tier)), time.time())
        sig = pub = ""
        if self._kp is not None:
            try:
                sig, pub = self._kp.sign(draft.canonical()).hex(), self._kp.public_key_bytes().hex()
            except Exception:  # noqa: BLE001
                logger.warning("credit_summary_sign_failed")
        return CreditSummary(**{**asdict(draft), "signature": sig, "public_key": pub})

    def _signature_ok(self, s: CreditSummary) -> bool:
        """Unsigned summaries are tolerated (older peers); a signature, when present, must verify."""
        if not s.signature:
            return True
        try:
            sig = bytes.fromhex(s.signature)
        except ValueError:
            return False
        if s.public_key:
            from infomesh_b200.p2p.keys import peer_id_from_public_key, verify_with_public_key

            try:
                pub = bytes.fromhex(s.public_key)
            except ValueError:
                return False
            return verify_with_public_key(pub, s.canonical(), sig) and (
                len(s.peer_id) != 40 or peer_id_from_public_key(pub) == s.peer_id)
        return True

    def receive_summary(self, summary: CreditSummary, *, verify_signature: bool = True) -> bool:
        if not self.has_identity or summary.owner_email_hash != self._owner_hash or summary.peer_id == self._peer_id:
            return False
        if summary.timestamp > time.time() + MAX_CLOCK_SKEW_SECONDS:
            logger.warning("credit_sync_future_timestamp", peer_id=summary.peer_id[:16])
            return False
        if verify_signature and not self._signature_ok(summary):
            logger.warning("credit_sync_bad_signature", peer_id=summary.peer_id[:16])
            return False
        if not self._store.has_peer(summary.peer_id) and self._store.peer_count(self._owner_hash) >= MAX_PEER_SUMMARIES:
            logger.warning("credit_sync_max_peers_reached")
            return False
        self._store.store_summary(summary)
        self._peers[summary.peer_id] = time.time()
        return True

    def aggregated_stats(self) -> AggregatedCreditStats:
        st = self._ledger.stats()
        peers = self._store.get_peer_summaries(self._owner_hash) if self.has_identity else []
        earned = st.total_earned + sum(p.total_earned for p in peers)
        spent = st.total_spent + sum(p.total_spent for p in peers)
        score = st.contribution_score + sum(p.contribution_score for p in peers)
        return AggregatedCreditStats(earned, spent, earned - spent, score, 1 + len(peers), peers)

    def needs_sync(self, peer_id: str) -> bool:
        return time.time() - self._peers.get(peer_id, 0.0) > SYNC_INTERVAL_SECONDS

    def register_same_owner_peer(self, peer_id: str) -> None:
        if peer_id != self._peer_id:
            self._peers[peer_id] = 0.0

    def get_same_owner_peers(self) -> list[str]:
        return list(self._peers)

    def purge_stale(self) -> int:
        return self._store.purge_stale()

    def close(self) -> None:
        self._store.close()
