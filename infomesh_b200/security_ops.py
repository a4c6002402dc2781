"""Operational security for the admin surfaces: API keys that can be rotated without downtime, and an audit trail.

Contract (SURVEY §2.1 "security ops"; reference infomesh/security_ops.py): keys are stored only as SHA-256 digests, in a
JSON file readable by the owner alone (0600); a key may expire (``ttl_days``), be revoked by label, or be rotated -- the
successor is added under ``<old>-rotated-<unix time>`` while the predecessor keeps working for a grace period; validation
compares in constant time and does not reveal which entry matched; a corrupt key file loads as "no keys".  The audit log
appends one JSON object per line (``ts, action, source, client, details<=500 chars, ok``), rolls over to ``.log.1`` past
the size limit, and never raises.

Implementation: key records know their own life-cycle (``usable_at``); the file is read and written through a pair of
functions that own the JSON shape; validation folds ``hmac.compare_digest`` over all records; the audit reader tails the
file with a bounded deque instead of splitting the whole text."""
from __future__ import annotations

import hmac
import json
import threading
import time
from collections import deque
from dataclasses import asdict, dataclass, fields
from hashlib import sha256
from pathlib import Path

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_DAY = 86400.0
_DETAIL_LIMIT = 500


# ----------------------------------------------------------------------------- API keys
@dataclass
class APIKeyEntry:
    key_hash: str
    label: str
    created_at: float
    expires_at: float | None = None
    revoked: bool = False

    def usable_at(self, instant: float) -> bool:
        return not self.revoked and (self.expires_at is None or instant < self.expires_at)

    def active(self, now: float | None = None) -> bool:
        return self.usable_at(time.time() if now is None else now)


def _fingerprint(secret: str) -> str:
    return sha256(secret.encode()).hexdigest()


def _entry_from_json(item: dict) -> APIKeyEntry:
    return APIKeyEntry(key_hash=str(item["key_hash"]), label=str(item.get("label", "")), created_at=float(item.get("created_at", 0)),
                       expires_at=item.get("expires_at"), revoked=bool(item.get("revoked", False)))


class APIKeyManager:
    def __init__(self, keys_file: Path | None = None):
        self._file = Path(keys_file) if keys_file else None
        self._keys: list[APIKeyEntry] = self._read_file()

    # ---- persistence
    def _read_file(self) -> list[APIKeyEntry]:
        if self._file is None or not self._file.exists():
            return []
        try:
            return [_entry_from_json(item) for item in json.loads(self._file.read_text("utf-8"))]
        except (ValueError, KeyError, TypeError, OSError) as exc:
            logger.warning("api_keys_load_failed", error=str(exc))
            return []

    def _write_file(self) -> None:
        if self._file is None:
            return
        try:
            self._file.parent.mkdir(parents=True, exist_ok=True)
            scratch = self._file.with_name(f"{self._file.name}.tmp")
            scratch.write_text(json.dumps([asdict(entry) for entry in self._keys], indent=2), encoding="utf-8")
            scratch.chmod(0o600)                # before it becomes visible under the real name
            scratch.replace(self._file)
        except OSError as exc:
            logger.warning("api_keys_save_failed", error=str(exc))

    # ---- life-cycle
    def add_key(self, key: str, label: str = "", ttl_days: int | None = None) -> APIKeyEntry:
        born = time.time()
        entry = APIKeyEntry(key_hash=_fingerprint(key), label=label or f"key-{len(self._keys) + 1}", created_at=born,
                            expires_at=born + ttl_days * _DAY if ttl_days else None)
        self._keys.append(entry)
        self._write_file()
        return entry

    def revoke(self, label: str) -> bool:
        matched = [entry for entry in self._keys if entry.label == label]
        for entry in matched:
            entry.revoked = True
        if matched:
            self._write_file()
        return bool(matched)

    def rotate(self, old_label: str, new_key: str, *, grace_days: int = 7) -> APIKeyEntry:
        stamp = time.time()
        for entry in self._keys:
            if entry.label == old_label and entry.usable_at(stamp):
                entry.expires_at = stamp + grace_days * _DAY      # clients get a grace window to switch
        return self.add_key(new_key, f"{old_label}-rotated-{int(stamp)}")

    # ---- queries
    def validate(self, key: str) -> bool:
        probe, instant = _fingerprint(key), time.time()
        verdicts = [hmac.compare_digest(entry.key_hash, probe) & entry.usable_at(instant) for entry in self._keys]
        return any(verdicts)                    # the list is built in full first: no early exit on a match

    def list_keys(self) -> list[dict[str, object]]:
        instant = time.time()
        return [{"label": e.label, "created": e.created_at, "expires": e.expires_at, "revoked": e.revoked, "active": e.usable_at(instant)}
                for e in self._keys]


# ----------------------------------------------------------------------------- audit trail
@dataclass
class AuditEntry:
    timestamp: float
    action: str
    source: str     # api | cli | mcp
    client: str
    details: str = ""
    success: bool = True


# wire key in the log line  <-  AuditEntry field
_LINE_KEYS = {"timestamp": "ts", "success": "ok"}
_FIELD_DEFAULTS = {"timestamp": 0.0, "action": "", "source": "", "client": "", "details": "", "success": True}


class AuditLogger:
    def __init__(self, log_path: Path | None = None, max_size_mb: int = 50):
        self._path = Path(log_path) if log_path else None
        self._limit = max_size_mb << 20
        self._mutex = threading.Lock()

    def _roll_if_large(self) -> None:
        if self._path.exists() and self._path.stat().st_size > self._limit:
            self._path.replace(self._path.with_suffix(".log.1"))

    def log(self, action: str, source: str = "api", client: str = "localhost", details: str = "", success: bool = True) -> None:
        if self._path is None:
            return
        record = {"ts": time.time(), "action": action, "source": source, "client": client, "details": details[:_DETAIL_LIMIT], "ok": success}
        try:
            with self._mutex:
                self._path.parent.mkdir(parents=True, exist_ok=True)
                self._roll_if_large()
                with self._path.open("a", encoding="utf-8") as sink:
                    sink.write(json.dumps(record) + "\n")
        except OSError:                         # auditing must never take the service down
            pass

    def recent(self, limit: int = 50) -> list[AuditEntry]:
        if self._path is None or not self._path.exists():
            return []
        try:
            with self._path.open(encoding="utf-8") as source:
                tail = deque(source, maxlen=limit)
        except OSError:
            return []
        entries = []
        for line in tail:
            try:
                raw = json.loads(line)
            except ValueError:
                continue
            values = {f.name: type(_FIELD_DEFAULTS[f.name])(raw.get(_LINE_KEYS.get(f.name, f.name), _FIELD_DEFAULTS[f.name]))
                      for f in fields(AuditEntry)}
            entries.append(AuditEntry(**values))
        return entries
