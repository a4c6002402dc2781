"""Content hashing helpers (SHA-256) — parity with reference infomesh/hashing.py:13-40."""
from __future__ import annotations

import hashlib


def content_hash(data: str | bytes) -> str:
    """Hex SHA-256 of text (UTF-8) or bytes."""
    if isinstance(data, str):
        data = data.encode("utf-8")
    return hashlib.sha256(data).hexdigest()


def short_hash(data: str | bytes, length: int = 16) -> str:
    """Prefix of :func:`content_hash` (cache keys, log correlation)."""
    return content_hash(data)[:max(1, length)]
