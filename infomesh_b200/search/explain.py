"""Explain mode: per-result breakdown of the six ranking signals, their weights and contributions, plus the pipeline
steps a query went through (reference infomesh/search/explain.py:23-164)."""
from __future__ import annotations

from dataclasses import dataclass, field

from infomesh_b200.index import ranking as R
from infomesh_b200.index.ranking import RankedResult

_WEIGHTS = {"bm25": R.WEIGHT_BM25, "freshness": R.WEIGHT_FRESHNESS, "trust": R.WEIGHT_TRUST, "authority": R.WEIGHT_AUTHORITY,
            "title_match": R.WEIGHT_TITLE_MATCH, "url_path": R.WEIGHT_URL_PATH}
DEFAULT_PIPELINE = ["sanitize_fts_query", "fts5_search", "bm25_ranking", "freshness_decay", "trust_scoring", "authority_scoring",
                    "combined_ranking"]
GPU_PIPELINE = ["tokenize", "encoder_forward", "sim_topk_dense", "bm25_intersect_score", "topk_exchange_merge", "rrf_fuse",
                "cross_encoder_rerank", "rerank_select"]


@dataclass
class ScoreExplanation:
    url: str
    title: str
    combined_score: float
    components: dict[str, float] = field(default_factory=dict)
    weights: dict[str, float] = field(default_factory=dict)
    weighted: dict[str, float] = field(default_factory=dict)
    notes: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, object]:
        r4 = lambda d: {k: round(v, 4) for k, v in d.items()}  # noqa: E731
        base = ("bm25", "freshness", "trust", "authority")       # the four signals the reference's tool output breaks a score into
        return {"url": self.url, "title": self.title, "combined_score": round(self.combined_score, 4), "components": r4(self.components),
                "weights": r4(self.weights), "weighted_contributions": r4(self.weighted), "notes": self.notes,
                "breakdown": {k: round(self.weighted.get(k, 0.0), 4) for k in base},
                "dominant_factor": max(self.weighted, key=lambda k: self.weighted.get(k, 0.0)) if self.weighted else ""}


@dataclass
class QueryExplanation:
    query: str
    sanitized_query: str
    total_results: int
    elapsed_ms: float
    results: list[ScoreExplanation]
    pipeline: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, object]:
        return {"query": self.query, "sanitized_query": self.sanitized_query, "total_results": self.total_results,
                "elapsed_ms": round(self.elapsed_ms, 1), "pipeline": self.pipeline, "results": [r.to_dict() for r in self.results]}


_NOTES = (("bm25", lambda v: v > 0.8, "Strong keyword match"), ("freshness", lambda v: v > 0.8, "Recently crawled"),
          ("freshness", lambda v: v < 0.2, "Stale content — may need recrawl"), ("trust", lambda v: v > 0.8, "High-trust peer"),
          ("authority", lambda v: v > 0.5, "High domain authority"), ("title_match", lambda v: v > 0.5, "Query matches title"),
          ("url_path", lambda v: v > 0.3, "Query matches URL path"))


def explain_result(result: RankedResult) -> ScoreExplanation:
    comp = {"bm25": result.bm25_score, "freshness": result.freshness_score, "trust": result.trust_score,
            "authority": result.authority_score, "title_match": result.title_match_score, "url_path": result.url_path_score}
    return ScoreExplanation(result.url, result.title, result.combined_score, comp, dict(_WEIGHTS),
                            {k: comp[k] * _WEIGHTS[k] for k in comp}, [msg for key, test, msg in _NOTES if test(comp[key])])


def explain_query(query: str, sanitized: str, results: list[RankedResult], elapsed_ms: float, *,
                  pipeline: list[str] | None = None) -> QueryExplanation:
    return QueryExplanation(query, sanitized, len(results), elapsed_ms, [explain_result(r) for r in results],
                            list(pipeline) if pipeline is not None else list(DEFAULT_PIPELINE))
