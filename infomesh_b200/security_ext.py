"""Extended security: TLS config, HS256 JWT verification (stdlib only), role-based tool access, IP allow/block lists,
webhook HMAC signatures, SQLite API audit log (reference infomesh/security_ext.py:27-371)."""
from __future__ import annotations

import base64
import hashlib
import hmac
import ipaddress
import json
import sqlite3
import threading
import time
from dataclasses import dataclass, field
from enum import StrEnum
from pathlib import Path
from typing import Any


@dataclass(frozen=True)
class TLSConfig:
    enabled: bool = False
    cert_file: str = ""
    key_file: str = ""
    ca_file: str = ""

    def validate(self) -> list[str]:
        if not self.enabled:
            return []
        errs = []
        for label, path in (("cert_file", self.cert_file), ("key_file", self.key_file)):
            if not path:
                errs.append(f"TLS {label} is required")
            elif not Path(path).exists():
                errs.append(f"TLS {label} not found: {path}")
        return errs

    def ssl_context(self) -> Any:
        if not self.enabled:
            return None
        import ssl

        ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
        ctx.minimum_version = ssl.TLSVersion.TLSv1_2
        ctx.load_cert_chain(self.cert_file, self.key_file)
        if self.ca_file:
            ctx.load_verify_locations(self.ca_file)
        return ctx


def _b64url(data: str) -> bytes:
    return base64.urlsafe_b64decode(data + "=" * (-len(data) % 4))


def verify_jwt_token(token: str, secret: str, *, algorithms: list[str] | None = None) -> dict[str, object] | None:
    """HS256 only; the header's ``alg`` must be in ``algorithms`` (so ``none`` can never pass); ``exp``/``nbf`` honoured."""
    allowed = algorithms or ["HS256"]
    try:
        head_b, body_b, sig_b = token.split(".")
        header = json.loads(_b64url(head_b))
        if header.get("alg") != "HS256" or "HS256" not in allowed:
            return None
        want = hmac.new(secret.encode(), f"{head_b}.{body_b}".encode(), hashlib.sha256).digest()
        if not hmac.compare_digest(want, _b64url(sig_b)):
            return None
        payload = json.loads(_b64url(body_b))
        if not isinstance(payload, dict):
            return None
        now = time.time()
        exp, nbf = payload.get("exp"), payload.get("nbf")
        if exp is not None and (not isinstance(exp, (int, float)) or exp < now):
            return None
        if isinstance(nbf, (int, float)) and nbf > now:
            return None
        return payload
    except Exception:  # noqa: BLE001 — any malformed token is simply invalid
        return None


def make_jwt_token(payload: dict[str, object], secret: str) -> str:
    """Companion of :func:`verify_jwt_token` for tests and the CLI (``infomesh keys token``)."""
    enc = lambda b: base64.urlsafe_b64encode(b).rstrip(b"=").decode()  # noqa: E731
    head = enc(json.dumps({"alg": "HS256", "typ": "JWT"}, separators=(",", ":")).encode())
    body = enc(json.dumps(payload, separators=(",", ":")).encode())
    sig = enc(hmac.new(secret.encode(), f"{head}.{body}".encode(), hashlib.sha256).digest())
    return f"{head}.{body}.{sig}"


class Role(StrEnum):
    ADMIN = "admin"
    READER = "reader"
    CRAWLER = "crawler"


_A, _R, _C = Role.ADMIN, Role.READER, Role.CRAWLER
_TOOL_ROLES: dict[str, set[Role]] = {
    "search": {_A, _R, _C}, "search_local": {_A, _R, _C}, "web_search": {_A, _R, _C}, "fetch_page": {_A, _R},
    "crawl_url": {_A, _C}, "network_stats": {_A, _R}, "status": {_A, _R}, "batch_search": {_A, _R}, "suggest": {_A, _R, _C},
    "register_webhook": {_A}, "analytics": {_A},
}


def check_role(tool_name: str, user_role: str | None) -> bool:
    if user_role is None:
        return True                    # RBAC not configured
    allowed = _TOOL_ROLES.get(tool_name)
    if allowed is None:
        return True
    try:
        return Role(user_role) in allowed
    except ValueError:
        return False


@dataclass
class IPFilter:
    """Entries may be single addresses or CIDR blocks.  The blocklist wins; a non-empty allowlist is exclusive."""
    allowlist: set[str] = field(default_factory=set)
    blocklist: set[str] = field(default_factory=set)

    @staticmethod
    def _hit(ip: str, entries: set[str]) -> bool:
        if ip in entries:
            return True
        try:
            addr = ipaddress.ip_address(ip)
        except ValueError:
            return False
        for e in entries:
            if "/" in e:
                try:
                    if addr in ipaddress.ip_network(e, strict=False):
                        return True
                except ValueError:
                    continue
        return False

    def is_allowed(self, ip: str) -> bool:
        if self._hit(ip, self.blocklist):
            return False
        return not self.allowlist or self._hit(ip, self.allowlist)

    def add_allow(self, ip: str) -> None:
        self.allowlist.add(ip)

    def add_block(self, ip: str) -> None:
        self.blocklist.add(ip)

    def remove_allow(self, ip: str) -> None:
        self.allowlist.discard(ip)

    def remove_block(self, ip: str) -> None:
        self.blocklist.discard(ip)


def sign_webhook_payload(payload: dict[str, object], secret: str) -> str:
    body = json.dumps(payload, sort_keys=True, ensure_ascii=False).encode("utf-8")
    return "sha256=" + hmac.new(secret.encode("utf-8"), body, hashlib.sha256).hexdigest()


def verify_webhook_signature(payload: dict[str, object], signature: str, secret: str) -> bool:
    return hmac.compare_digest(sign_webhook_payload(payload, secret), signature)


class AuditLog:
    _SECRET_ARGS = frozenset({"api_key", "password", "secret", "token"})

    def __init__(self, db_path: Path | str | None = None):
        self._db_path = str(db_path) if db_path else ":memory:"
        if self._db_path != ":memory:":
            Path(self._db_path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(self._db_path, check_same_thread=False)
        self._conn.row_factory = sqlite3.Row
        self._lock = threading.Lock()
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("CREATE TABLE IF NOT EXISTS audit_log (id INTEGER PRIMARY KEY AUTOINCREMENT, timestamp REAL NOT NULL, "
                           "tool_name TEXT NOT NULL, api_key_hash TEXT, client_ip TEXT, arguments_json TEXT, "
                           "success INTEGER NOT NULL DEFAULT 1, latency_ms REAL DEFAULT 0)")
        self._conn.commit()

    def log(self, tool_name: str, *, api_key: str | None = None, client_ip: str | None = None,
            arguments: dict[str, Any] | None = None, success: bool = True, latency_ms: float = 0) -> None:
        key_hash = hashlib.sha256(api_key.encode()).hexdigest()[:16] if api_key else None
        args = json.dumps({k: v for k, v in arguments.items() if k not in self._SECRET_ARGS}, default=str)[:1000] if arguments else None
        with self._lock:
            self._conn.execute("INSERT INTO audit_log (timestamp, tool_name, api_key_hash, client_ip, arguments_json, success, "
                               "latency_ms) VALUES (?, ?, ?, ?, ?, ?, ?)",
                               (time.time(), tool_name, key_hash, client_ip, args, int(success), latency_ms))
            self._conn.commit()

    def query(self, *, limit: int = 100, tool_name: str | None = None, since: float | None = None) -> list[dict[str, object]]:
        sql, params = "SELECT * FROM audit_log WHERE 1=1", []
        if tool_name:
            sql += " AND tool_name = ?"
            params.append(tool_name)
        if since:
            sql += " AND timestamp >= ?"
            params.append(since)
        with self._lock:
            return [dict(r) for r in self._conn.execute(sql + " ORDER BY timestamp DESC LIMIT ?", [*params, limit])]

    def close(self) -> None:
        self._conn.close()

    def __enter__(self) -> "AuditLog":
        return self

    def __exit__(self, *args: object) -> None:
        self.close()
