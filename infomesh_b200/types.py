"""Structural typing seams so tests can inject fakes (reference infomesh/types.py:16-84)."""
from __future__ import annotations

from collections.abc import Callable
from typing import Any, Protocol, runtime_checkable


@runtime_checkable
class KeyPairLike(Protocol):
    @property
    def peer_id(self) -> str: ...

    def sign(self, data: bytes) -> bytes: ...

    def verify(self, data: bytes, signature: bytes) -> bool: ...

    def public_key_bytes(self) -> bytes: ...


@runtime_checkable
class VectorStoreLike(Protocol):
    def add_document(self, doc_id: int, url: str, title: str, text: str, language: str | None = None) -> None: ...

    def search(self, query: str, limit: int = 10) -> list[Any]: ...

    def delete_document(self, doc_id: int) -> None: ...

    def get_stats(self) -> dict[str, Any]: ...

    def close(self) -> None: ...


AuthorityFn = Callable[[str], float]
