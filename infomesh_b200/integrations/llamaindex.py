"""LlamaIndex-style reader: ``load_data(query)`` / ``lazy_load_data`` -> ``LlamaDocument(text, metadata, id_)``
(reference infomesh/integrations/llamaindex.py:18-96)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Iterator

from infomesh_b200.integrations._base import ClientBacked, result_meta


@dataclass
class LlamaDocument:
    text: str
    metadata: dict[str, Any] = field(default_factory=dict)
    id_: str = ""
    extra_info: dict[str, Any] = field(default_factory=dict)


class InfoMeshReader(ClientBacked):
    def __init__(self, data_dir: str = "~/.infomesh", limit: int = 5, **kw: Any):
        super().__init__(data_dir, **kw)
        self._limit = limit

    def load_data(self, query: str, **kwargs: Any) -> list[LlamaDocument]:
        return [LlamaDocument(r.snippet, result_meta(r), id_=r.url)
                for r in self._ensure_client().search(query, limit=kwargs.get("limit", self._limit))]

    def lazy_load_data(self, query: str, **kwargs: Any) -> Iterator[LlamaDocument]:
        yield from self.load_data(query, **kwargs)
