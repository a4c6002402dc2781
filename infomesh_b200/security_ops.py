"""Operational security: rotating API keys stored as SHA-256 digests (TTL, revoke, constant-time compare) and an
append-only JSON-lines audit log with size rotation (reference infomesh/security_ops.py:23-233)."""
from __future__ import annotations

import hashlib
import hmac
import json
import threading
import time
from dataclasses import asdict, dataclass
from pathlib import Path

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass
class APIKeyEntry:
    key_hash: str
    label: str
    created_at: float
    expires_at: float | None = None
    revoked: bool = False

    def active(self, now: float | None = None) -> bool:
        return not self.revoked and (self.expires_at is None or (now or time.time()) < self.expires_at)


class APIKeyManager:
    def __init__(self, keys_file: Path | None = None):
        self._file = Path(keys_file) if keys_file else None
        self._keys: list[APIKeyEntry] = []
        self._load()

    @staticmethod
    def _digest(key: str) -> str:
        return hashlib.sha256(key.encode()).hexdigest()

    def add_key(self, key: str, label: str = "", ttl_days: int | None = None) -> APIKeyEntry:
        now = time.time()
        entry = APIKeyEntry(self._digest(key), label or f"key-{len(self._keys) + 1}", now, now + ttl_days * 86400 if ttl_days else None)
        self._keys.append(entry)
        self._save()
        return entry

    def validate(self, key: str) -> bool:
        digest, now = self._digest(key), time.time()
        ok = False
        for e in self._keys:       # no early exit: timing independent of which key matched
            ok |= e.active(now) and hmac.compare_digest(e.key_hash, digest)
        return ok

    def revoke(self, label: str) -> bool:
        hit = [e for e in self._keys if e.label == label]
        for e in hit:
            e.revoked = True
        if hit:
            self._save()
        return bool(hit)

    def rotate(self, old_label: str, new_key: str, *, grace_days: int = 7) -> APIKeyEntry:
        """Add ``new_key`` and let the old one expire after a grace period instead of cutting clients off."""
        for e in self._keys:
            if e.label == old_label and e.active():
                e.expires_at = time.time() + grace_days * 86400
        return self.add_key(new_key, f"{old_label}-rotated-{int(time.time())}")

    def list_keys(self) -> list[dict[str, object]]:
        now = time.time()
        return [{"label": e.label, "created": e.created_at, "expires": e.expires_at, "revoked": e.revoked, "active": e.active(now)}
                for e in self._keys]

    def _load(self) -> None:
        if not self._file or not self._file.exists():
            return
        try:
            for it in json.loads(self._file.read_text("utf-8")):
                self._keys.append(APIKeyEntry(str(it["key_hash"]), str(it.get("label", "")), float(it.get("created_at", 0)),
                                              it.get("expires_at"), bool(it.get("revoked", False))))
        except (json.JSONDecodeError, KeyError, TypeError, OSError) as exc:
            logger.warning("api_keys_load_failed", error=str(exc))

    def _save(self) -> None:
        if not self._file:
            return
        try:
            self._file.parent.mkdir(parents=True, exist_ok=True)
            tmp = self._file.with_name(self._file.name + ".tmp")
            tmp.write_text(json.dumps([asdict(e) for e in self._keys], indent=2), encoding="utf-8")
            tmp.chmod(0o600)
            tmp.replace(self._file)
        except OSError as exc:
            logger.warning("api_keys_save_failed", error=str(exc))


@dataclass
class AuditEntry:
    timestamp: float
    action: str
    source: str     # api | cli | mcp
    client: str
    details: str = ""
    success: bool = True


class AuditLogger:
    """Never raises: auditing must not take the service down."""

    def __init__(self, log_path: Path | None = None, max_size_mb: int = 50):
        self._path = Path(log_path) if log_path else None
        self._max_bytes = max_size_mb * 2 ** 20
        self._lock = threading.Lock()

    def log(self, action: str, source: str = "api", client: str = "localhost", details: str = "", success: bool = True) -> None:
        if not self._path:
            return
        line = json.dumps({"ts": time.time(), "action": action, "source": source, "client": client, "details": details[:500],
                           "ok": success})
        try:
            with self._lock:
                self._path.parent.mkdir(parents=True, exist_ok=True)
                if self._path.exists() and self._path.stat().st_size > self._max_bytes:
                    self._path.replace(self._path.with_suffix(".log.1"))
                with open(self._path, "a", encoding="utf-8") as f:
                    f.write(line + "\n")
        except OSError:
            pass

    def recent(self, limit: int = 50) -> list[AuditEntry]:
        if not self._path or not self._path.exists():
            return []
        out = []
        try:
            for line in self._path.read_text("utf-8").strip().split("\n")[-limit:]:
                try:
                    d = json.loads(line)
                except json.JSONDecodeError:
                    continue
                out.append(AuditEntry(float(d.get("ts", 0)), str(d.get("action", "")), str(d.get("source", "")),
                                      str(d.get("client", "")), str(d.get("details", "")), bool(d.get("ok", True))))
        except OSError:
            pass
        return out
