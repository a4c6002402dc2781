"""Reciprocal Rank Fusion of keyword (FTS5 / GPU BM25) and vector results, keyed by URL.

``RRF(d) = sum_s w_s / (60 + rank_s(d))`` (reference infomesh/search/merge.py:20,37-133).  The batched GPU form is
``ops.fuse.rrf_fuse`` (keyed by document id).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

_RRF_K = 60


@dataclass(frozen=True)
class MergedResult:
    doc_id: str
    url: str
    title: str
    snippet: str
    fts_score: float | None
    vector_score: float | None
    combined_score: float
    source: str  # "fts" | "vector" | "hybrid"


def merge_results(fts_results: list[Any], vector_results: list[Any], *, limit: int = 10, fts_weight: float = 1.0,
                  vector_weight: float = 1.0, rrf_k: int = _RRF_K) -> list[MergedResult]:
    """Fuse two ranked lists.  ``fts_results`` carry ``.snippet``/``.score``; vector results ``.text_preview``/``.score``."""
    slots: dict[str, dict[str, Any]] = {}
    for rank, r in enumerate(fts_results, start=1):
        e = slots.setdefault(r.url, {"doc_id": str(r.doc_id), "url": r.url, "title": r.title, "snippet": "",
                                     "fts": None, "vec": None, "rrf": 0.0})
        if e["fts"] is None:
            e["fts"] = float(r.score)
            e["snippet"] = r.snippet or e["snippet"]
            e["rrf"] += fts_weight / (rrf_k + rank)
    for rank, r in enumerate(vector_results, start=1):
        e = slots.setdefault(r.url, {"doc_id": str(r.doc_id), "url": r.url, "title": r.title, "snippet": "",
                                     "fts": None, "vec": None, "rrf": 0.0})
        if e["vec"] is None:
            e["vec"] = float(r.score)
            if not e["snippet"]:
                e["snippet"] = getattr(r, "text_preview", "") or ""
            if not e["title"]:
                e["title"] = r.title
            e["rrf"] += vector_weight / (rrf_k + rank)
    merged = [MergedResult(doc_id=e["doc_id"], url=e["url"], title=e["title"], snippet=e["snippet"],
                           fts_score=e["fts"], vector_score=e["vec"], combined_score=round(e["rrf"], 6),
                           source="hybrid" if e["fts"] is not None and e["vec"] is not None
                           else ("fts" if e["fts"] is not None else "vector")) for e in slots.values()]
    merged.sort(key=lambda m: m.combined_score, reverse=True)
    return merged[:limit]
