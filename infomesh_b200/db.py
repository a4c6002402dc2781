"""SQLite base class shared by the ledger / trust / peer stores (reference infomesh/db.py:37-98):
WAL journal, 5 s busy timeout, schema bootstrap, context-manager close."""
from __future__ import annotations

import sqlite3
import threading
from pathlib import Path


class SQLiteStore:
    _SCHEMA: str = ""

    def __init__(self, db_path: Path | str | None = None, *, check_same_thread: bool = False, row_factory: type | None = None,
                 extra_pragmas: list[str] | None = None):
        self._path = str(db_path) if db_path else ":memory:"
        if self._path != ":memory:":
            Path(self._path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(self._path, check_same_thread=check_same_thread)
        self._conn.row_factory = row_factory if row_factory is not None else sqlite3.Row      # stores here read columns by name
        self._lock = threading.RLock()
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA busy_timeout=5000")
        for pragma in extra_pragmas or ():
            self._conn.execute(pragma)
        if self._SCHEMA:
            self._conn.executescript(self._SCHEMA)
            self._conn.commit()
        self._post_init()

    def _post_init(self) -> None:  # hook for migrations
        return None

    @property
    def conn(self) -> sqlite3.Connection:
        return self._conn

    @property
    def path(self) -> str:
        return self._path

    def close(self) -> None:
        try:
            self._conn.close()
        except sqlite3.Error:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()
