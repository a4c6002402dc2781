"""Haystack-style document store: ``query`` / ``count_documents`` / ``write_documents``
(reference infomesh/integrations/haystack.py:18-114).  ``write_documents`` indexes documents that carry content
directly and crawls the ones that only carry a URL."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

from infomesh_b200.integrations._base import ClientBacked


@dataclass
class HaystackDocument:
    content: str
    meta: dict[str, Any] = field(default_factory=dict)
    id: str = ""
    score: float | None = None


class InfoMeshDocumentStore(ClientBacked):
    def query(self, query: str, *, top_k: int = 10, **kwargs: Any) -> list[HaystackDocument]:
        return [HaystackDocument(r.snippet, {"title": r.title, "url": r.url, "source": "infomesh"}, id=r.url, score=r.score)
                for r in self._ensure_client().search(query, limit=top_k)]

    def count_documents(self) -> int:
        n = self._ensure_client().get_stats().get("total_documents", 0)
        return int(n) if isinstance(n, (int, float)) else 0

    def write_documents(self, documents: list[HaystackDocument]) -> int:
        client, ok = self._ensure_client(), 0
        for d in documents:
            url = str(d.meta.get("url", "") or d.id)
            if not url:
                continue
            if d.content and len(d.content) >= 50:
                ok += client.add_document(url, str(d.meta.get("title", "")), d.content) is not None
            else:
                ok += client.crawl(url).success
        return ok
