"""Cross-validation of peer search responses: a URL is trusted when >= 50 % of the queried peers return it with
consistent scores (max deviation <= 3x median) and snippets (pairwise Jaccard >= 0.2); single-source URLs are flagged
fabricated (reference infomesh/search/cross_validate.py:27-287)."""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from itertools import combinations

AGREEMENT_THRESHOLD = 0.5
MIN_PEERS_FOR_VALIDATION = 2
SNIPPET_SIMILARITY_THRESHOLD = 0.20
SCORE_DEVIATION_RATIO = 3.0
VERDICT_TRUSTED, VERDICT_UNVERIFIED, VERDICT_SUSPICIOUS, VERDICT_FABRICATED = "trusted", "unverified", "suspicious", "fabricated"
_WORD = re.compile(r"\w+")


@dataclass(frozen=True)
class PeerResult:
    peer_id: str
    url: str
    title: str
    snippet: str
    score: float


@dataclass(frozen=True)
class ValidatedResult:
    url: str
    title: str
    snippet: str
    score: float
    verdict: str
    agreement_ratio: float
    appearing_peers: list[str]
    score_deviation: float
    detail: str


@dataclass(frozen=True)
class CrossValidationReport:
    query: str
    total_peers: int
    results: list[ValidatedResult]
    suspicious_count: int
    fabricated_count: int
    detail: str


@dataclass
class _UrlAggregation:
    url: str
    title: str
    snippet: str
    peers: list[str] = field(default_factory=list)
    scores: list[float] = field(default_factory=list)
    snippets: list[str] = field(default_factory=list)


def snippet_similarity(a: str, b: str) -> float:
    wa, wb = set(_WORD.findall(a.lower())), set(_WORD.findall(b.lower()))
    return len(wa & wb) / len(wa | wb) if wa and wb else 0.0


def _median(scores: list[float]) -> float:
    return sorted(scores)[len(scores) // 2]


def _score_deviation(scores: list[float]) -> float:
    if len(scores) <= 1:
        return 0.0
    med = _median(scores)
    return max(abs(s - med) for s in scores) / med if med > 0 else 0.0


def _has_snippet_mismatch(snippets: list[str]) -> bool:
    real = [s for s in snippets if s.strip()]
    return any(snippet_similarity(a, b) < SNIPPET_SIMILARITY_THRESHOLD for a, b in combinations(real, 2))


def cross_validate_results(query: str, peer_results: dict[str, list[PeerResult]]) -> CrossValidationReport:
    n = len(peer_results)
    if n < MIN_PEERS_FOR_VALIDATION:
        seen: dict[str, ValidatedResult] = {}
        for pid, rs in peer_results.items():
            for r in rs:
                seen.setdefault(r.url, ValidatedResult(r.url, r.title, r.snippet, r.score, VERDICT_UNVERIFIED, 1.0, [pid], 0.0,
                                                       "insufficient peers for validation"))
        return CrossValidationReport(query, n, list(seen.values()), 0, 0, "cross-validation skipped: insufficient peers")
    agg: dict[str, _UrlAggregation] = {}
    for pid, rs in peer_results.items():
        for r in rs:
            a = agg.setdefault(r.url, _UrlAggregation(r.url, r.title, r.snippet))
            if pid not in a.peers:          # one vote per peer even if it repeats a URL
                a.peers.append(pid)
                a.scores.append(r.score)
                a.snippets.append(r.snippet)
    out: list[ValidatedResult] = []
    for a in agg.values():
        ratio, dev = len(a.peers) / n, _score_deviation(a.scores)
        if ratio >= AGREEMENT_THRESHOLD:
            if dev > SCORE_DEVIATION_RATIO:
                verdict, detail = VERDICT_SUSPICIOUS, f"score deviation={dev:.2f}"
            elif _has_snippet_mismatch(a.snippets):
                verdict, detail = VERDICT_SUSPICIOUS, "snippet similarity below threshold"
            else:
                verdict, detail = VERDICT_TRUSTED, "ok"
        elif len(a.peers) == 1:
            verdict, detail = VERDICT_FABRICATED, f"only 1/{n} peers returned this URL"
        else:
            verdict, detail = VERDICT_SUSPICIOUS, f"low agreement: {ratio:.0%}"
        out.append(ValidatedResult(a.url, a.title, a.snippet, _median(a.scores), verdict, round(ratio, 4), a.peers, round(dev, 4), detail))
    out.sort(key=lambda v: (v.agreement_ratio, v.score), reverse=True)
    sus, fab = sum(v.verdict == VERDICT_SUSPICIOUS for v in out), sum(v.verdict == VERDICT_FABRICATED for v in out)
    detail = f"{len(out)} URLs validated across {n} peers" + (f", {sus} suspicious" if sus else "") + (f", {fab} fabricated" if fab else "")
    return CrossValidationReport(query, n, out, sus, fab, detail)
