"""Composite ranking: BM25 saturation-normalisation, freshness decay, trust, authority, title / URL bonuses.

Constants and formulas follow reference infomesh/index/ranking.py:24-148,171-285 (weights .40/.15/.10/.15 +
bonuses .15/.05, 7-day half-life with a 0.05 floor, ``s / (s + max_s)``).  The batched GPU form of the same
weighted sum lives in the merge epilogue (K12).
"""
from __future__ import annotations

import math
import time
from collections.abc import Callable
from dataclasses import dataclass
from typing import Any

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

WEIGHT_BM25 = 0.40
WEIGHT_FRESHNESS = 0.15
WEIGHT_TRUST = 0.10
WEIGHT_AUTHORITY = 0.15
WEIGHT_TITLE_MATCH = 0.15
WEIGHT_URL_PATH = 0.05
FRESHNESS_HALF_LIFE_SECONDS: float = 7 * 24 * 3600
MIN_FRESHNESS: float = 0.05
DEFAULT_TRUST: float = 0.50


@dataclass(frozen=True)
class RankedResult:
    doc_id: str | int
    url: str
    title: str
    snippet: str
    bm25_score: float
    freshness_score: float
    trust_score: float
    authority_score: float
    combined_score: float
    crawled_at: float
    peer_id: str | None = None
    title_match_score: float = 0.0
    url_path_score: float = 0.0


@dataclass
class _RawCandidate:
    doc_id: str | int
    url: str
    title: str
    snippet: str
    bm25_raw: float
    crawled_at: float
    trust: float = DEFAULT_TRUST
    authority: float = 0.0
    peer_id: str | None = None
    title_match: float = 0.0
    url_path: float = 0.0


RawCandidate = _RawCandidate


def freshness_score(crawled_at: float, *, now: float | None = None,
                    half_life: float = FRESHNESS_HALF_LIFE_SECONDS) -> float:
    """Exponential decay 2^(-age / half_life), floored at ``MIN_FRESHNESS``; future timestamps count as fresh."""
    now = time.time() if now is None else now
    age = max(0.0, now - crawled_at)
    return max(MIN_FRESHNESS, math.pow(2.0, -age / half_life))


def normalize_bm25(score: float, *, max_score: float = 1.0) -> float:
    """Saturation ``s / (s + k)`` with ``k = max_score`` (a score equal to it maps to 0.5)."""
    if score <= 0:
        return 0.0
    return score / (score + max(max_score, 0.0))


def combined_score(bm25: float, freshness: float, trust: float, authority: float = 0.0, *, title_match: float = 0.0,
                   url_path: float = 0.0, w_bm25: float = WEIGHT_BM25, w_fresh: float = WEIGHT_FRESHNESS,
                   w_trust: float = WEIGHT_TRUST, w_authority: float = WEIGHT_AUTHORITY,
                   w_title: float = WEIGHT_TITLE_MATCH, w_url: float = WEIGHT_URL_PATH) -> float:
    return (w_bm25 * bm25 + w_fresh * freshness + w_trust * trust + w_authority * authority
            + w_title * title_match + w_url * url_path)


def rank_results(candidates: list[_RawCandidate], *, limit: int = 10, now: float | None = None,
                 weights: dict[str, float] | None = None) -> list[RankedResult]:
    if not candidates:
        return []
    now = now or time.time()
    top = max(c.bm25_raw for c in candidates) or 1.0
    w = weights or {}
    ranked: list[RankedResult] = []
    for c in candidates:
        nb = normalize_bm25(c.bm25_raw, max_score=top)
        fr = freshness_score(c.crawled_at, now=now)
        total = combined_score(nb, fr, c.trust, c.authority, title_match=c.title_match, url_path=c.url_path, **w)
        ranked.append(RankedResult(
            doc_id=c.doc_id, url=c.url, title=c.title, snippet=c.snippet, bm25_score=round(nb, 6),
            freshness_score=round(fr, 6), trust_score=round(c.trust, 6), authority_score=round(c.authority, 6),
            combined_score=round(total, 6), crawled_at=c.crawled_at, peer_id=c.peer_id,
            title_match_score=round(c.title_match, 6), url_path_score=round(c.url_path, 6)))
    ranked.sort(key=lambda r: r.combined_score, reverse=True)
    logger.debug("results_ranked", candidates=len(candidates), returned=min(limit, len(ranked)))
    return ranked[:limit]


def rank_local_results(results: list[Any], *, trust: float = DEFAULT_TRUST,
                       authority_fn: Callable[[str], float] | None = None, query_tokens: list[str] | None = None,
                       limit: int = 10, now: float | None = None) -> list[RankedResult]:
    """Rank ``LocalStore.search`` rows; title / URL bonuses need ``query_tokens``."""
    from infomesh_b200.search.passage import title_match_score, url_path_score

    cands = []
    for r in results:
        auth = 0.0
        if authority_fn is not None:
            try:
                auth = float(authority_fn(r.url))
            except Exception:  # noqa: BLE001 — a broken authority source must not break search
                auth = 0.0
        cands.append(_RawCandidate(
            doc_id=r.doc_id, url=r.url, title=r.title, snippet=r.snippet, bm25_raw=r.score, crawled_at=r.crawled_at,
            trust=trust, authority=auth,
            title_match=title_match_score(r.title, query_tokens) if query_tokens else 0.0,
            url_path=url_path_score(r.url, query_tokens) if query_tokens else 0.0))
    return rank_results(cands, limit=limit, now=now)
