"""Search-quality tooling: NDCG@k / MRR and A/B comparison of rankers, per-category ranking weight profiles, cache
pre-warm queries, per-domain diversification, temporal hints ("last 3 days", "2024"), rule-based intent classes
(reference infomesh/search/quality.py:26-408)."""
from __future__ import annotations

import math
import re
import time
from collections import Counter
from dataclasses import dataclass, field
from urllib.parse import urlparse


def ndcg_at_k(relevance_scores: list[float], k: int = 10) -> float:
    if not relevance_scores:
        return 0.0
    dcg = lambda rels: sum(r / math.log2(i + 2) for i, r in enumerate(rels[:k]))  # noqa: E731
    ideal = dcg(sorted(relevance_scores, reverse=True))
    return dcg(relevance_scores) / ideal if ideal > 0 else 0.0


def mrr(ranks: list[int]) -> float:
    return sum(1.0 / r for r in ranks if r > 0) / len(ranks) if ranks else 0.0


@dataclass
class ABTestResult:
    test_name: str
    query: str
    variant_a_ndcg: float
    variant_b_ndcg: float
    winner: str             # "A" | "B" | "tie" (|diff| <= 0.01)
    improvement_pct: float


class ABTest:
    def __init__(self, name: str):
        self.name = name
        self.results: list[ABTestResult] = []

    def compare(self, query: str, scores_a: list[float], scores_b: list[float], k: int = 10) -> ABTestResult:
        a, b = ndcg_at_k(scores_a, k), ndcg_at_k(scores_b, k)
        diff = b - a
        res = ABTestResult(self.name, query, round(a, 4), round(b, 4), "B" if diff > 0.01 else "A" if diff < -0.01 else "tie",
                           round(diff / a * 100, 2) if a > 0 else 0.0)
        self.results.append(res)
        return res

    def summary(self) -> dict[str, object]:
        wins = Counter(r.winner for r in self.results)
        n = len(self.results)
        return {"test": self.name, "total": n, "A_wins": wins.get("A", 0), "B_wins": wins.get("B", 0), "ties": wins.get("tie", 0),
                "avg_improvement": round(sum(r.improvement_pct for r in self.results) / n, 2) if n else 0.0}


@dataclass(frozen=True)
class RankingProfile:
    name: str
    bm25_weight: float = 0.40
    freshness_weight: float = 0.15
    trust_weight: float = 0.10
    authority_weight: float = 0.15
    title_weight: float = 0.15
    url_weight: float = 0.05


RANKING_PROFILES: dict[str, RankingProfile] = {
    "default": RankingProfile("default"),
    "tech-docs": RankingProfile("tech-docs", bm25_weight=0.50, freshness_weight=0.05, title_weight=0.20, url_weight=0.10),
    "news": RankingProfile("news", bm25_weight=0.25, freshness_weight=0.45, trust_weight=0.15, authority_weight=0.10, title_weight=0.05),
    "academic": RankingProfile("academic", bm25_weight=0.35, freshness_weight=0.05, trust_weight=0.20, authority_weight=0.25,
                               title_weight=0.10, url_weight=0.05),
}
_CATEGORY_HINTS = (("tech-docs", ("docs.", "documentation", "readthedocs", "devdocs", "developer.", "api.")),
                   ("news", ("news", "bbc", "reuters", "cnn", "nytimes")),
                   ("academic", ("arxiv", "scholar", "academic", "ieee", "springer", "pubmed")))


def get_profile(name: str) -> RankingProfile:
    return RANKING_PROFILES.get(name, RANKING_PROFILES["default"])


def detect_domain_category(url: str) -> str:
    low = url.lower()
    return next((cat for cat, hints in _CATEGORY_HINTS if any(h in low for h in hints)), "default")


@dataclass
class PrewarmConfig:
    popular_queries: list[str] = field(default_factory=list)
    max_queries: int = 100
    interval_seconds: int = 3600


DEFAULT_PREWARM_QUERIES = ["python tutorial", "javascript async await", "react hooks", "docker compose", "kubernetes deployment",
                           "git rebase", "sql join", "css flexbox", "rust ownership", "golang goroutine"]


def _host(r: dict[str, object]) -> str:
    try:
        return urlparse(str(r.get("url", ""))).netloc or "unknown"
    except ValueError:
        return "unknown"


@dataclass
class ResultCluster:
    domain: str
    results: list[dict[str, object]] = field(default_factory=list)
    representative_title: str = ""


def cluster_results(results: list[dict[str, object]], max_per_domain: int = 3) -> list[ResultCluster]:
    by: dict[str, ResultCluster] = {}
    for r in results:
        c = by.setdefault(_host(r), ResultCluster(_host(r), representative_title=str(r.get("title", ""))))
        if len(c.results) < max_per_domain:
            c.results.append(r)
    return list(by.values())


def diversify_results(results: list[dict[str, object]], max_per_domain: int = 3) -> list[dict[str, object]]:
    """Round-robin over domains in first-appearance order, at most ``max_per_domain`` from each."""
    queues: dict[str, list[dict[str, object]]] = {}
    for r in results:
        queues.setdefault(_host(r), []).append(r)
    for q in queues.values():
        del q[max_per_domain:]
    out: list[dict[str, object]] = []
    while queues:
        for dom in list(queues):
            out.append(queues[dom].pop(0))
            if not queues[dom]:
                del queues[dom]
    return out


_TEMPORAL = ((r"\b(?:today|tonight)\b", 1), (r"\byesterday\b", 2), (r"\bthis\s+week\b", 7), (r"\blast\s+week\b", 14),
             (r"\bthis\s+month\b", 30), (r"\blast\s+month\b", 60), (r"\bthis\s+year\b", 365), (r"\blast\s+year\b", 730),
             (r"\blast\s+(\d+)\s+days?\b", -1), (r"\b(?:latest|newest|recent)\b", 7), (r"\b20[2-3]\d\b", -2))


def extract_temporal_hint(query: str) -> int | None:
    """Recency window in days implied by the query wording, else None."""
    for pat, days in _TEMPORAL:
        m = re.search(pat, query, re.IGNORECASE)
        if not m:
            continue
        if days == -1:
            return int(m.group(1))
        if days == -2:
            return max(1, (time.localtime().tm_year - int(m.group(0)) + 1) * 365)
        return days
    return None


class QueryIntentClassifier:
    INTENTS = {
        "how_to": [r"\bhow\s+(?:to|do|can|does)\b", r"\btutorial\b", r"\bguide\b", r"\bstep.by.step\b"],
        "definition": [r"\bwhat\s+is\b", r"\bdefin(?:e|ition)\b", r"\bmeaning\s+of\b"],
        "comparison": [r"\bvs\.?\b", r"\bversus\b", r"\bcompare\b", r"\bdifference\s+between\b", r"\bor\b.*\bwhich\b"],
        "error_debug": [r"\berror\b", r"\bexception\b", r"\btraceback\b", r"\bfailed?\b", r"\bnot\s+work", r"\bbug\b"],
        "api_reference": [r"\bapi\b", r"\bfunction\b.*\bsignature\b", r"\bmethod\b.*\bparameter", r"\breturn\s+type\b"],
        "navigational": [r"\blogin\b", r"\bofficial\b", r"\bhomepage\b", r"\bdownload\b", r"\.(?:com|org|io|dev)$"],
    }
    _compiled = {k: [re.compile(p, re.IGNORECASE) for p in v] for k, v in INTENTS.items()}

    def classify(self, query: str) -> str:
        return next((name for name, pats in self._compiled.items() if any(p.search(query) for p in pats)), "informational")

    def classify_with_confidence(self, query: str) -> tuple[str, float]:
        hits = {name: sum(bool(p.search(query)) for p in pats) for name, pats in self._compiled.items()}
        best = max(hits, key=hits.get)
        return (best, round(min(1.0, hits[best] / 3.0), 2)) if hits[best] else ("informational", 0.3)
