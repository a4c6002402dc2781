"""Facets (domain / language / recency buckets), keyword clustering, term highlighting and near-duplicate removal for
result lists (reference infomesh/search/facets.py:21-254)."""
from __future__ import annotations

import re
import time
from collections import Counter
from dataclasses import dataclass, field
from urllib.parse import urlparse

from infomesh_b200.index.ranking import RankedResult

_WORD = re.compile(r"\w+")
_BUCKETS = ((1, "today"), (7, "this_week"), (30, "this_month"), (365, "this_year"))


@dataclass
class FacetCounts:
    domains: dict[str, int] = field(default_factory=dict)
    languages: dict[str, int] = field(default_factory=dict)
    date_ranges: dict[str, int] = field(default_factory=dict)

    def to_dict(self) -> dict[str, dict[str, int]]:
        return {"domains": dict(sorted(self.domains.items(), key=lambda kv: kv[1], reverse=True)[:20]),
                "languages": dict(self.languages), "date_ranges": dict(self.date_ranges)}


def _age_bucket(crawled_at: float, now: float) -> str:
    if not crawled_at or crawled_at > now + 86400:
        return "unknown"
    days = (now - crawled_at) / 86400
    return next((name for limit, name in _BUCKETS if days <= limit), "older")


def compute_facets(results: list[RankedResult], *, max_domains: int = 20, languages: dict[str, str] | None = None) -> FacetCounts:
    """``languages`` optionally maps url -> language code (RankedResult itself carries no language)."""
    now = time.time()
    dom, lang, dates = Counter(), Counter(), Counter()
    for r in results:
        host = urlparse(r.url).netloc
        if host:
            dom[host] += 1
        if languages and r.url in languages:
            lang[languages[r.url]] += 1
        dates[_age_bucket(r.crawled_at, now)] += 1
    return FacetCounts(dict(dom.most_common(max_domains)), dict(lang), dict(dates))


@dataclass
class ResultCluster:
    label: str
    results: list[RankedResult]
    score: float = 0.0


def cluster_results(results: list[RankedResult], *, max_clusters: int = 5, min_cluster_size: int = 2) -> list[ResultCluster]:
    """Greedy: the most shared >3-letter keywords each claim the still-unassigned results containing them."""
    if len(results) < min_cluster_size:
        return []
    toks = [{w for w in _WORD.findall(f"{r.title} {r.snippet}".lower()) if len(w) > 3} for r in results]
    freq = Counter(w for t in toks for w in t)
    free = set(range(len(results)))
    clusters: list[ResultCluster] = []
    for kw, cnt in freq.most_common(max_clusters * 3):
        if len(clusters) >= max_clusters:
            break
        if cnt < min_cluster_size:
            continue
        members = [i for i in sorted(free) if kw in toks[i]]
        if len(members) >= min_cluster_size:
            free -= set(members)
            rs = [results[i] for i in members]
            clusters.append(ResultCluster(kw, rs, sum(r.combined_score for r in rs) / len(rs)))
    return sorted(clusters, key=lambda c: c.score, reverse=True)


def highlight_snippet(snippet: str, query: str, *, marker: str = "**") -> str:
    terms = {t for t in query.lower().split() if t}
    if not terms:
        return snippet
    pat = re.compile(r"\b(" + "|".join(re.escape(t) for t in sorted(terms, key=len, reverse=True)) + r")\b", re.IGNORECASE)
    return pat.sub(lambda m: f"{marker}{m.group(0)}{marker}", snippet)


def dedup_results(results: list[RankedResult], *, similarity_threshold: float = 0.7) -> list[RankedResult]:
    """Drop repeated URLs (trailing slash ignored) and results whose title+snippet word-Jaccard with an already kept
    result reaches the threshold."""
    kept: list[RankedResult] = []
    kept_tok: list[set[str]] = []
    urls: set[str] = set()
    for r in results:
        u = r.url.rstrip("/")
        if u in urls:
            continue
        t = set(f"{r.title} {r.snippet}".lower().split())
        if t and any(k and len(t & k) / len(t | k) >= similarity_threshold for k in kept_tok):
            continue
        urls.add(u)
        kept.append(r)
        kept_tok.append(t)
    return kept
