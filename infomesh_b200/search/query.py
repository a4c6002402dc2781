"""Query orchestration: local (FTS5 -> rank -> passage), hybrid (keyword + vector -> RRF), distributed
(local + peers -> dedupe -> rank).  Mirrors reference infomesh/search/query.py:54-531; the GPU engine plugs in
through ``gpu_engine=`` (batched device pipeline, ``engine.hybrid``) while the CPU path stays the oracle.
"""
from __future__ import annotations

import re
import time
from collections.abc import Callable
from dataclasses import dataclass, replace
from math import isfinite
from typing import TYPE_CHECKING, Any

from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.index.ranking import RankedResult, rank_local_results
from infomesh_b200.types import VectorStoreLike
from infomesh_b200.utils.log import get_logger

if TYPE_CHECKING:  # pragma: no cover
    from infomesh_b200.index.distributed import DistributedIndex

logger = get_logger(__name__)

_FTS_SPECIALS = re.compile(r'["\(\)\*\{\}\^:]')
_FTS_OPERATORS = re.compile(r"\b(AND|OR|NOT|NEAR)\b", re.IGNORECASE)
_WS = re.compile(r"\s+")
MAX_QUERY_CHARS = 1000


@dataclass(frozen=True)
class QueryResult:
    results: list[RankedResult]
    total: int
    elapsed_ms: float
    source: str  # "local" | "network"


@dataclass(frozen=True)
class HybridResult:
    results: list[Any]  # MergedResult
    total: int
    elapsed_ms: float
    source: str  # "hybrid" | "fts" | "vector"


@dataclass
class DistributedResult:
    results: list[RankedResult]
    total: int
    elapsed_ms: float
    source: str  # "distributed" | "local_only"
    local_count: int = 0
    remote_count: int = 0


def _sanitize_fts_query(query: str) -> str:
    """Strip FTS5 syntax (quotes, parens, ``* { } ^ :``) and boolean / NEAR operators; never return empty."""
    query = query[:MAX_QUERY_CHARS]
    cleaned = _WS.sub(" ", _FTS_OPERATORS.sub(" ", _FTS_SPECIALS.sub(" ", query))).strip()
    if cleaned:
        return cleaned
    alnum = _WS.sub(" ", re.sub(r"[^a-zA-Z0-9\s]", " ", query)).strip()[:100]
    return alnum or "infomesh"


sanitize_fts_query = _sanitize_fts_query


def search_local(store: LocalStore, query: str, *, limit: int = 10, offset: int = 0,
                 authority_fn: Callable[[str], float] | None = None, language: str | None = None,
                 date_from: float | None = None, date_to: float | None = None,
                 include_domains: list[str] | None = None, exclude_domains: list[str] | None = None) -> QueryResult:
    from infomesh_b200.search.cjk import tokenize_query_cjk
    from infomesh_b200.search.nlp import expand_query
    from infomesh_b200.search.passage import _tokenize

    t0 = time.monotonic()
    filters = dict(language=language, date_from=date_from, date_to=date_to, include_domains=include_domains,
                   exclude_domains=exclude_domains)
    fts_query = _sanitize_fts_query(tokenize_query_cjk(query))
    rows = store.search(fts_query, limit=limit * 2, offset=offset, **filters)
    if len(rows) < limit:  # sparse: widen with synonyms (each expansion is its own AND query)
        seen = {r.url for r in rows}
        for term in expand_query(query, max_expansions=3):
            tq = _sanitize_fts_query(term)
            if not tq or tq == "infomesh":
                continue
            for r in store.search(tq, limit=limit, offset=0, **filters):
                if r.url not in seen:
                    seen.add(r.url)
                    rows.append(r)
    ranked = rank_local_results(rows, authority_fn=authority_fn, query_tokens=_tokenize(query), limit=limit)
    if ranked:
        _enhance_snippets(store, ranked, query)
    elapsed = (time.monotonic() - t0) * 1000
    logger.info("query_local", query=query, raw=len(rows), ranked=len(ranked), elapsed_ms=round(elapsed, 1))
    return QueryResult(ranked, len(ranked), elapsed, "local")


def _enhance_snippets(store: LocalStore, results: list[RankedResult], query: str, *, max_enhance: int = 10) -> None:
    """Swap weak FTS5 snippets (short or without a query term) for the best passage of the full text."""
    from infomesh_b200.search.passage import _tokenize, select_best_passage

    wanted = set(_tokenize(query))
    for i, r in enumerate(results[:max_enhance]):
        if len(r.snippet) >= 80 and wanted & set(_tokenize(r.snippet)):
            continue
        try:
            doc = store.get_document(int(r.doc_id))
        except (TypeError, ValueError):
            continue
        if doc is None or not doc.text:
            continue
        passage = select_best_passage(doc.text, query, max_length=300)
        if passage and len(passage) > len(r.snippet):
            results[i] = replace(r, snippet=passage)


def search_hybrid(store: LocalStore, vector_store: VectorStoreLike, query: str, *, limit: int = 10,
                  fts_weight: float = 1.0, vector_weight: float = 1.0,
                  authority_fn: Callable[[str], float] | None = None) -> HybridResult:
    """Keyword top-``limit`` + vector top-``limit`` fused with RRF (k = 60)."""
    from infomesh_b200.search.merge import merge_results

    if not (hasattr(vector_store, "search") and hasattr(vector_store, "add_document")):
        raise TypeError(f"vector_store must be a VectorStore, got {type(vector_store).__name__}")
    t0 = time.monotonic()
    fts = store.search(_sanitize_fts_query(query), limit=limit)
    vec = vector_store.search(query, limit=limit)
    merged = merge_results(fts, vec, limit=limit, fts_weight=fts_weight, vector_weight=vector_weight)
    elapsed = (time.monotonic() - t0) * 1000
    has_fts = any(m.fts_score is not None for m in merged)
    has_vec = any(m.vector_score is not None for m in merged)
    source = "hybrid" if has_fts and has_vec else ("vector" if has_vec else "fts")
    logger.info("query_hybrid", query=query, fts_count=len(fts), vec_count=len(vec), merged_count=len(merged),
                elapsed_ms=round(elapsed, 1))
    return HybridResult(merged, len(merged), elapsed, source)


# ------------------------------------------------------------------ distributed
def _safe_remote_int(value: object, *, default: int = 0) -> int:
    if isinstance(value, bool) or not isinstance(value, (int, float, str)):
        return default
    try:
        f = float(value)
        return int(f) if isfinite(f) else default
    except (TypeError, ValueError, OverflowError):
        return default


def _safe_remote_float(value: object, *, default: float = 0.0) -> float:
    if isinstance(value, bool) or not isinstance(value, (int, float, str)):
        return default
    try:
        f = float(value)
    except (TypeError, ValueError):
        return default
    return f if isfinite(f) else default


def _make_remote_result(*, url: str, title: str, snippet: str, score: object, doc_id: object,
                        peer_id: str) -> RankedResult:
    """Remote hits carry only the peer's own score; every local signal is zero."""
    return RankedResult(doc_id=_safe_remote_int(doc_id), url=url, title=title, snippet=snippet, bm25_score=0.0,
                        freshness_score=0.0, trust_score=0.0, authority_score=0.0,
                        combined_score=_safe_remote_float(score), crawled_at=0.0, peer_id=peer_id)


async def search_distributed(store: LocalStore, distributed_index: "DistributedIndex | None", query: str, *,
                             limit: int = 10, authority_fn: Callable[[str], float] | None = None,
                             vector_store: VectorStoreLike | None = None,
                             network_search_fn: Callable[[str, list[str], int], Any] | None = None
                             ) -> DistributedResult:
    """Local search + (peer fan-out through ``network_search_fn`` | DHT pointer stubs), deduped by URL."""
    from infomesh_b200.index.distributed import extract_keywords

    t0 = time.monotonic()
    query = _sanitize_fts_query(query)
    local = search_local(store, query, limit=limit, authority_fn=authority_fn)
    keywords = extract_keywords(query, max_keywords=10)
    remote: list[RankedResult] = []
    if keywords and network_search_fn is not None:
        try:
            for r in await network_search_fn(query, keywords, limit):
                if isinstance(r, dict) and r.get("url"):
                    remote.append(_make_remote_result(
                        url=str(r["url"]), title=str(r.get("title", "")), snippet=str(r.get("snippet", "")),
                        score=r.get("score", 0.0), doc_id=r.get("doc_id", 0), peer_id=str(r.get("peer_id", ""))))
        except Exception:  # noqa: BLE001 — search is never blocked by the network
            logger.exception("network_search_failed")
    elif keywords and distributed_index is not None:
        try:
            for ptr in await distributed_index.query(keywords):
                remote.append(_make_remote_result(url=ptr.url, title=ptr.title, snippet="", score=ptr.score,
                                                  doc_id=ptr.doc_id, peer_id=ptr.peer_id))
        except Exception:  # noqa: BLE001
            logger.exception("dht_query_failed")
    remote_count = len(remote)
    seen: set[str] = set()
    merged: list[RankedResult] = []
    for r in [*local.results, *remote]:  # local first wins on duplicates
        if r.url not in seen:
            seen.add(r.url)
            merged.append(r)
    merged.sort(key=lambda r: r.combined_score, reverse=True)
    merged = merged[:limit]
    elapsed = (time.monotonic() - t0) * 1000
    source = "distributed" if remote_count > 0 else "local_only"
    logger.info("query_distributed", query=query, local_count=local.total, remote_count=remote_count,
                merged=len(merged), elapsed_ms=round(elapsed, 1))
    return DistributedResult(merged, len(merged), elapsed, source, local.total, remote_count)
