"""SQLite cache of peers we connected to (``<data_dir>/peer_store.db``) so a restart can rejoin the mesh without any
bootstrap node: upsert on success, failure counting, freshest-first loading that hides peers failing > 80 % of
>= 5 attempts, age / size pruning (reference infomesh/p2p/peer_store.py:36-228)."""
from __future__ import annotations

import time
from dataclasses import dataclass
from pathlib import Path

from infomesh_b200.db import SQLiteStore

DEFAULT_MAX_PEERS = 200
DEFAULT_MAX_AGE_HOURS = 168
DEFAULT_LOAD_LIMIT = 20


@dataclass(frozen=True)
class CachedPeer:
    peer_id: str
    multiaddr: str
    last_seen: float
    success_count: int
    fail_count: int

    @property
    def success_rate(self) -> float:
        n = self.success_count + self.fail_count
        return self.success_count / n if n else 0.0


class PeerStore(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS peers (
            peer_id TEXT PRIMARY KEY, multiaddr TEXT NOT NULL, last_seen REAL NOT NULL,
            success_count INTEGER NOT NULL DEFAULT 1, fail_count INTEGER NOT NULL DEFAULT 0);
        CREATE INDEX IF NOT EXISTS idx_peers_last_seen ON peers (last_seen DESC);
    """
    _UPSERT = ("INSERT INTO peers (peer_id, multiaddr, last_seen, success_count, fail_count) VALUES (?, ?, ?, 1, 0) "
               "ON CONFLICT(peer_id) DO UPDATE SET multiaddr = excluded.multiaddr, last_seen = excluded.last_seen, "
               "success_count = success_count + 1")

    def __init__(self, data_dir: Path | str):
        super().__init__(Path(data_dir) / "peer_store.db")

    def upsert(self, peer_id: str, multiaddr: str) -> None:
        with self._lock:
            self._conn.execute(self._UPSERT, (peer_id, multiaddr, time.time()))
            self._conn.commit()

    def save_connected(self, peers: list[tuple[str, str]]) -> None:
        now = time.time()
        with self._lock:
            self._conn.executemany(self._UPSERT, [(pid, addr, now) for pid, addr in peers])
            self._conn.commit()

    def record_failure(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("UPDATE peers SET fail_count = fail_count + 1 WHERE peer_id = ?", (peer_id,))
            self._conn.commit()

    def remove(self, peer_id: str) -> None:
        with self._lock:
            self._conn.execute("DELETE FROM peers WHERE peer_id = ?", (peer_id,))
            self._conn.commit()

    def load_recent(self, limit: int = DEFAULT_LOAD_LIMIT) -> list[CachedPeer]:
        rows = self._conn.execute(
            "SELECT peer_id, multiaddr, last_seen, success_count, fail_count FROM peers "
            "WHERE (success_count + fail_count) < 5 "
            "   OR CAST(success_count AS REAL) / (success_count + fail_count) > 0.2 "
            "ORDER BY last_seen DESC LIMIT ?", (limit,)).fetchall()
        return [CachedPeer(*tuple(r)) for r in rows]

    def count(self) -> int:
        return int(self._conn.execute("SELECT COUNT(*) FROM peers").fetchone()[0])

    def prune(self, max_age_hours: float = DEFAULT_MAX_AGE_HOURS, max_peers: int = DEFAULT_MAX_PEERS) -> int:
        with self._lock:
            removed = self._conn.execute("DELETE FROM peers WHERE last_seen < ?",
                                         (time.time() - max_age_hours * 3600,)).rowcount
            extra = self.count() - max_peers
            if extra > 0:
                self._conn.execute("DELETE FROM peers WHERE peer_id IN "
                                   "(SELECT peer_id FROM peers ORDER BY last_seen ASC LIMIT ?)", (extra,))
                removed += extra
            self._conn.commit()
        return removed
