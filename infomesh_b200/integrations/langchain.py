"""LangChain-style retriever: ``invoke`` / ``ainvoke`` / ``get_relevant_documents`` -> ``Document(page_content,
metadata)`` (reference infomesh/integrations/langchain.py:19-118)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

from infomesh_b200.integrations._base import ClientBacked, result_meta


@dataclass
class Document:
    page_content: str
    metadata: dict[str, Any] = field(default_factory=dict)


class InfoMeshRetriever(ClientBacked):
    def __init__(self, data_dir: str = "~/.infomesh", limit: int = 5, **kw: Any):
        super().__init__(data_dir, **kw)
        self._limit = limit

    @staticmethod
    def _docs(results) -> list[Document]:
        return [Document(r.snippet, result_meta(r)) for r in results]

    def invoke(self, query: str, **kwargs: Any) -> list[Document]:
        return self._docs(self._ensure_client().search(query, limit=kwargs.get("limit", self._limit)))

    def get_relevant_documents(self, query: str) -> list[Document]:
        return self.invoke(query)

    async def ainvoke(self, query: str, **kwargs: Any) -> list[Document]:
        return self._docs(await self._ensure_client().search_async(query, limit=kwargs.get("limit", self._limit)))

    def batch(self, queries: list[str], **kwargs: Any) -> list[list[Document]]:
        """Runnable.batch: a single device pass when the client is GPU-backed."""
        return [self._docs(rs) for rs in self._ensure_client().search_many(queries, limit=kwargs.get("limit", self._limit))]
