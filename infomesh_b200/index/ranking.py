"""Composite ranking of search candidates (reference infomesh/index/ranking.py).

Behavioural contract (Appendix B): six signals -- BM25 squashed by ``s / (s + max_s)`` over the batch, freshness
``max(0.05, 2^(-age / 7 d))``, trust (default 0.5), domain authority, title match, URL-path match -- combined with weights
0.40 / 0.15 / 0.10 / 0.15 / 0.15 / 0.05, sorted descending, top ``limit``; every reported component rounded to 6 places.

Implementation: the signals are rows of a declarative table (name, default weight, how to read it from a candidate), a
batch is turned into one ``[n, 6]`` signal matrix and scored with a single matrix-vector product (NumPy).  The same matrix
form is what the device epilogue consumes (``ops.fuse.rank_fuse`` when a CUDA batch path asks for it), so CPU and GPU
ranking share one definition of the signals."""
from __future__ import annotations

import time
from collections.abc import Callable, Sequence
from dataclasses import dataclass
from typing import Any

import numpy as np

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

# (keyword of combined_score / ``weights`` override, default weight) in signal-matrix column order
SIGNALS: tuple[tuple[str, float], ...] = (
    ("w_bm25", WEIGHT_BM25), ("w_fresh", WEIGHT_FRESHNESS), ("w_trust", WEIGHT_TRUST), ("w_authority", WEIGHT_AUTHORITY),
    ("w_title", WEIGHT_TITLE_MATCH), ("w_url", WEIGHT_URL_PATH))
DEFAULT_WEIGHTS = np.array([w for _, w in SIGNALS], dtype=np.float64)


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


# ----------------------------------------------------------------------------- scalar forms (public API)
def freshness_score(crawled_at: float, *, now: float | None = None, half_life: float = FRESHNESS_HALF_LIFE_SECONDS) -> float:
    """``2^(-age / half_life)`` floored at ``MIN_FRESHNESS``; timestamps in the future count as brand new."""
    age = max(0.0, (time.time() if now is None else now) - crawled_at)
    return max(MIN_FRESHNESS, 2.0 ** (-age / half_life))


def normalize_bm25(score: float, *, max_score: float = 1.0) -> float:
    """Saturating squash ``s / (s + k)`` with ``k = max_score``: the batch's best hit maps to 0.5, nothing reaches 1."""
    return score / (score + max(max_score, 0.0)) if score > 0 else 0.0


def combined_score(bm25: float, freshness: float, trust: float, authority: float = 0.0, *, title_match: float = 0.0,
                   url_path: float = 0.0, w_bm25: float = WEIGHT_BM25, w_fresh: float = WEIGHT_FRESHNESS,
                   w_trust: float = WEIGHT_TRUST, w_authority: float = WEIGHT_AUTHORITY,
                   w_title: float = WEIGHT_TITLE_MATCH, w_url: float = WEIGHT_URL_PATH) -> float:
    signals = (bm25, freshness, trust, authority, title_match, url_path)
    weights = (w_bm25, w_fresh, w_trust, w_authority, w_title, w_url)
    return sum(w * s for w, s in zip(weights, signals))


# ----------------------------------------------------------------------------- batch form
def weight_vector(overrides: dict[str, float] | None = None) -> np.ndarray:
    """Default weights with ``{"w_bm25": ..., ...}`` overrides applied (unknown keys are an error, as with kwargs)."""
    w = DEFAULT_WEIGHTS.copy()
    if overrides:
        names = [n for n, _ in SIGNALS]
        for key, value in overrides.items():
            if key not in names:
                raise TypeError(f"combined_score() got an unexpected keyword argument {key!r}")
            w[names.index(key)] = float(value)
    return w


def signal_matrix(candidates: Sequence[_RawCandidate], now: float) -> np.ndarray:
    """``[n, 6]`` float64 matrix in :data:`SIGNALS` column order (BM25 already squashed by the batch maximum)."""
    raw = np.fromiter((c.bm25_raw for c in candidates), dtype=np.float64, count=len(candidates))
    top = float(raw.max()) if raw.size else 0.0
    top = top if top else 1.0
    age = np.maximum(0.0, now - np.fromiter((c.crawled_at for c in candidates), dtype=np.float64, count=len(candidates)))
    sig = np.empty((len(candidates), len(SIGNALS)), dtype=np.float64)
    sig[:, 0] = np.where(raw > 0, raw / (raw + max(top, 0.0)), 0.0)
    sig[:, 1] = np.maximum(MIN_FRESHNESS, np.exp2(-age / FRESHNESS_HALF_LIFE_SECONDS))
    sig[:, 2] = [c.trust for c in candidates]
    sig[:, 3] = [c.authority for c in candidates]
    sig[:, 4] = [c.title_match for c in candidates]
    sig[:, 5] = [c.url_path for c in candidates]
    return sig


def rank_results(candidates: list[_RawCandidate], *, limit: int = 10, now: float | None = None,
                 weights: dict[str, float] | None = None) -> list[RankedResult]:
    if not candidates:
        return []
    stamp = now or time.time()
    sig = signal_matrix(candidates, stamp)
    total = sig @ weight_vector(weights)
    order = np.argsort(-np.round(total, 6), kind="stable")[:limit]      # ties keep retrieval order, like a stable sort
    out = []
    for i in order:
        c, row = candidates[int(i)], np.round(sig[int(i)], 6)
        out.append(RankedResult(doc_id=c.doc_id, url=c.url, title=c.title, snippet=c.snippet, bm25_score=float(row[0]),
                                freshness_score=float(row[1]), trust_score=float(row[2]), authority_score=float(row[3]),
                                combined_score=float(round(float(total[int(i)]), 6)), crawled_at=c.crawled_at, peer_id=c.peer_id,
                                title_match_score=float(row[4]), url_path_score=float(row[5])))
    logger.debug("results_ranked", candidates=len(candidates), returned=len(out))
    return out


def rank_local_results(results: list[Any], *, trust: float = DEFAULT_TRUST, authority_fn: Callable[[str], float] | None = None,
                       query_tokens: list[str] | None = None, limit: int = 10, now: float | None = None) -> list[RankedResult]:
    """Rank ``LocalStore.search`` rows.  ``authority_fn`` failures count as zero authority (a broken side table must not
    break search); the title / URL-path bonuses apply only when ``query_tokens`` are given."""
    from infomesh_b200.search.passage import title_match_score, url_path_score

    def authority_of(url: str) -> float:
        if authority_fn is None:
            return 0.0
        try:
            return float(authority_fn(url))
        except Exception:  # noqa: BLE001
            return 0.0

    batch = [_RawCandidate(doc_id=r.doc_id, url=r.url, title=r.title, snippet=r.snippet, bm25_raw=r.score, crawled_at=r.crawled_at,
                           trust=trust, authority=authority_of(r.url),
                           title_match=title_match_score(r.title, query_tokens) if query_tokens else 0.0,
                           url_path=url_path_score(r.url, query_tokens) if query_tokens else 0.0)
             for r in results]
    return rank_results(batch, limit=limit, now=now)
