"""The three ways a query is answered: from this node's store, from store + vector index, or from the whole network.

Contract (reference infomesh/search/query.py):

* ``search_local``: sanitise the text for FTS5 (CJK-aware), fetch ``2 x limit`` BM25 rows, widen with up to three synonym
  expansions when fewer than ``limit`` came back, rank with the six-signal ranker, then replace weak snippets (short, or
  without a query term) by the best passage of the document.
* ``search_hybrid``: keyword top-``limit`` and vector top-``limit`` fused by reciprocal-rank fusion; ``source`` says which
  of the two lists actually contributed.
* ``search_distributed``: local results plus remote ones -- from the peer fan-out callback when given, else from DHT
  pointer lookups -- de-duplicated by URL with local copies winning, sorted by combined score; remote failures never fail
  the search; remote numbers are untrusted input and are coerced defensively.

The sanitiser never returns an empty string (fallback: the alphanumeric residue, then the literal ``infomesh``).  The GPU
engine serves the same contracts through ``engine.gpu_index``; this module stays the CPU oracle.

Implementation: the sanitiser is a character translation table plus one operator pattern; timing is a small stopwatch
object; URL de-duplication is one ordered-dict helper used by both the synonym widening and the network merge; the two
remote sources are generator functions with one shared error boundary."""
from __future__ import annotations

import re
import time
from collections.abc import AsyncIterator, Callable, Iterable
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

MAX_QUERY_CHARS = 1000
_FALLBACK_QUERY = "infomesh"
_FTS_SYNTAX_TO_SPACE = str.maketrans({ch: " " for ch in '"()*{}^:'})
_FTS_KEYWORDS = re.compile(r"\b(?:AND|OR|NOT|NEAR)\b", re.IGNORECASE)
_NOT_PLAIN = re.compile(r"[^a-zA-Z0-9\s]")


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


class _Stopwatch:
    def __init__(self):
        self._t0 = time.monotonic()

    @property
    def ms(self) -> float:
        return (time.monotonic() - self._t0) * 1000.0


def _squeeze(text: str) -> str:
    return " ".join(text.split())


def _sanitize_fts_query(query: str) -> str:
    """Plain words only: FTS5 punctuation and the AND / OR / NOT / NEAR keywords are blanked out."""
    clipped = query[:MAX_QUERY_CHARS]
    words = _squeeze(_FTS_KEYWORDS.sub(" ", clipped.translate(_FTS_SYNTAX_TO_SPACE)))
    return words or _squeeze(_NOT_PLAIN.sub(" ", clipped))[:100] or _FALLBACK_QUERY


sanitize_fts_query = _sanitize_fts_query


def _first_per_url(items: Iterable[Any]) -> list[Any]:
    """Order-preserving de-duplication on ``.url``: the first occurrence wins."""
    unique: dict[str, Any] = {}
    for item in items:
        unique.setdefault(item.url, item)
    return list(unique.values())


# ----------------------------------------------------------------------------- local
def search_local(store: LocalStore, query: str, *, limit: int = 10, offset: int = 0,
                 authority_fn: Callable[[str], float] | None = None, language: str | None = None,
                 date_from: float | None = None, date_to: float | None = None,
                 include_domains: list[str] | None = None, exclude_domains: list[str] | None = None) -> QueryResult:
    from infomesh_b200.search.cjk import tokenize_query_cjk
    from infomesh_b200.search.nlp import expand_query
    from infomesh_b200.search.passage import _tokenize

    watch = _Stopwatch()
    narrowing = {"language": language, "date_from": date_from, "date_to": date_to, "include_domains": include_domains,
                 "exclude_domains": exclude_domains}
    rows = list(store.search(_sanitize_fts_query(tokenize_query_cjk(query)), limit=2 * limit, offset=offset, **narrowing))
    if len(rows) < limit:
        # thin result: each synonym expansion is run as its own (implicit AND) query and appended behind the originals
        variants = (_sanitize_fts_query(term) for term in expand_query(query, max_expansions=3))
        for variant in variants:
            if variant and variant != _FALLBACK_QUERY:
                rows.extend(store.search(variant, limit=limit, offset=0, **narrowing))
        rows = _first_per_url(rows)
    ranked = rank_local_results(rows, authority_fn=authority_fn, query_tokens=_tokenize(query), limit=limit)
    if ranked:
        _enhance_snippets(store, ranked, query)
    logger.info("query_local", query=query, raw=len(rows), ranked=len(ranked), elapsed_ms=round(watch.ms, 1))
    return QueryResult(ranked, len(ranked), watch.ms, "local")


def _enhance_snippets(store: LocalStore, results: list[RankedResult], query: str, *, max_enhance: int = 10) -> None:
    """In place: a snippet under 80 characters, or without any query term, is replaced by the best passage when longer."""
    from infomesh_b200.search.passage import _tokenize, select_best_passage

    query_terms = frozenset(_tokenize(query))

    def good_enough(snippet: str) -> bool:
        return len(snippet) >= 80 and not query_terms.isdisjoint(_tokenize(snippet))

    for slot, hit in enumerate(results[:max_enhance]):
        if good_enough(hit.snippet):
            continue
        try:
            document = store.get_document(int(hit.doc_id))
        except (TypeError, ValueError):
            continue
        body = getattr(document, "text", "")
        better = select_best_passage(body, query, max_length=300) if body else ""
        if better and len(better) > len(hit.snippet):
            results[slot] = replace(hit, snippet=better)


# ----------------------------------------------------------------------------- hybrid
def search_hybrid(store: LocalStore, vector_store: VectorStoreLike, query: str, *, limit: int = 10,
                  fts_weight: float = 1.0, vector_weight: float = 1.0,
                  authority_fn: Callable[[str], float] | None = None) -> HybridResult:
    from infomesh_b200.search.merge import merge_results

    if not all(hasattr(vector_store, method) for method in ("search", "add_document")):
        raise TypeError(f"vector_store must be a VectorStore, got {type(vector_store).__name__}")
    watch = _Stopwatch()
    keyword_hits = store.search(_sanitize_fts_query(query), limit=limit)
    vector_hits = vector_store.search(query, limit=limit)
    fused = merge_results(keyword_hits, vector_hits, limit=limit, fts_weight=fts_weight, vector_weight=vector_weight)
    from_keywords = any(m.fts_score is not None for m in fused)
    from_vectors = any(m.vector_score is not None for m in fused)
    origin = {(True, True): "hybrid", (False, True): "vector"}.get((from_keywords, from_vectors), "fts")
    logger.info("query_hybrid", query=query, fts_count=len(keyword_hits), vec_count=len(vector_hits), merged_count=len(fused),
                elapsed_ms=round(watch.ms, 1))
    return HybridResult(fused, len(fused), watch.ms, origin)


# ----------------------------------------------------------------------------- distributed
def _finite(value: object) -> float | None:
    """A finite float out of untrusted input (numbers or numeric strings; booleans and containers are rejected)."""
    if isinstance(value, bool) or not isinstance(value, (int, float, str)):
        return None
    try:
        number = float(value)
    except (TypeError, ValueError, OverflowError):
        return None
    return number if isfinite(number) else None


def _safe_remote_int(value: object, *, default: int = 0) -> int:
    number = _finite(value)
    return default if number is None else int(number)


def _safe_remote_float(value: object, *, default: float = 0.0) -> float:
    number = _finite(value)
    return default if number is None else number


def _make_remote_result(*, url: str, title: str, snippet: str, score: object, doc_id: object, peer_id: str) -> RankedResult:
    """A remote hit is ranked by its peer's score alone: none of the local signals is known for it."""
    return RankedResult(doc_id=_safe_remote_int(doc_id), url=url, title=title, snippet=snippet, bm25_score=0.0, freshness_score=0.0,
                        trust_score=0.0, authority_score=0.0, combined_score=_safe_remote_float(score), crawled_at=0.0, peer_id=peer_id)


async def _from_peers(fan_out, query: str, keywords: list[str], limit: int) -> AsyncIterator[RankedResult]:
    for item in await fan_out(query, keywords, limit):
        if isinstance(item, dict) and item.get("url"):
            yield _make_remote_result(url=str(item["url"]), title=str(item.get("title", "")), snippet=str(item.get("snippet", "")),
                                      score=item.get("score", 0.0), doc_id=item.get("doc_id", 0), peer_id=str(item.get("peer_id", "")))


async def _from_dht(index: "DistributedIndex", keywords: list[str]) -> AsyncIterator[RankedResult]:
    for pointer in await index.query(keywords):
        yield _make_remote_result(url=pointer.url, title=pointer.title, snippet="", score=pointer.score, doc_id=pointer.doc_id,
                                  peer_id=pointer.peer_id)


async def search_distributed(store: LocalStore, distributed_index: "DistributedIndex | None", query: str, *, limit: int = 10,
                             authority_fn: Callable[[str], float] | None = None, vector_store: VectorStoreLike | None = None,
                             network_search_fn: Callable[[str, list[str], int], Any] | None = None) -> DistributedResult:
    from infomesh_b200.index.distributed import extract_keywords

    watch = _Stopwatch()
    query = _sanitize_fts_query(query)
    mine = search_local(store, query, limit=limit, authority_fn=authority_fn)
    keywords = extract_keywords(query, max_keywords=10)
    if not keywords:
        stream, failure_event = None, ""
    elif network_search_fn is not None:
        stream, failure_event = _from_peers(network_search_fn, query, keywords, limit), "network_search_failed"
    elif distributed_index is not None:
        stream, failure_event = _from_dht(distributed_index, keywords), "dht_query_failed"
    else:
        stream, failure_event = None, ""
    theirs: list[RankedResult] = []
    if stream is not None:
        try:
            async for hit in stream:
                theirs.append(hit)
        except Exception:  # noqa: BLE001 -- the network must never be able to fail a search
            logger.exception(failure_event)
    combined = sorted(_first_per_url([*mine.results, *theirs]), key=lambda r: r.combined_score, reverse=True)[:limit]
    logger.info("query_distributed", query=query, local_count=mine.total, remote_count=len(theirs), merged=len(combined),
                elapsed_ms=round(watch.ms, 1))
    return DistributedResult(combined, len(combined), watch.ms, "distributed" if theirs else "local_only", mine.total, len(theirs))
