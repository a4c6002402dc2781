"""Result-quality helpers: freshness labels / grades, trust grades, citation extraction, claim cross-reference
(reference infomesh/data_quality.py:31-272)."""
from __future__ import annotations

import re
import time
from dataclasses import dataclass, field
from typing import Any

_MIN, _HOUR, _DAY, _WEEK, _MONTH, _QUARTER = 60, 3600, 86400, 604800, 2592000, 7776000


@dataclass(frozen=True)
class FreshnessIndicator:
    crawled_at: float
    age_seconds: float
    age_label: str
    freshness_grade: str


def _plural(n: int, unit: str) -> str:
    return f"{n} {unit}{'' if n == 1 else 's'} ago"


def compute_freshness_indicator(crawled_at: float, *, now: float | None = None) -> FreshnessIndicator:
    now = now or time.time()
    age = max(0.0, now - crawled_at)
    if age < _MIN:
        label = "just now"
    elif age < _HOUR:
        label = _plural(int(age / _MIN), "minute")
    elif age < _DAY:
        label = _plural(int(age / _HOUR), "hour")
    elif age < _WEEK:
        label = _plural(int(age / _DAY), "day")
    elif age < _MONTH:
        label = _plural(int(age / _WEEK), "week")
    else:
        label = _plural(int(age / _MONTH), "month")
    grade = "A" if age < _DAY else "B" if age < _WEEK else "C" if age < _MONTH else "D" if age < _QUARTER else "F"
    return FreshnessIndicator(crawled_at, age, label, grade)


@dataclass(frozen=True)
class TrustGrade:
    score: float
    grade: str
    label: str
    color: str


_TRUST_BANDS = ((0.9, "A+", "Highly Trusted", "green"), (0.8, "A", "Trusted", "green"), (0.65, "B", "Reliable", "blue"),
                (0.5, "C", "Moderate", "yellow"), (0.3, "D", "Low Trust", "orange"))


def compute_trust_grade(trust_score: float) -> TrustGrade:
    for floor, grade, label, color in _TRUST_BANDS:
        if trust_score >= floor:
            return TrustGrade(trust_score, grade, label, color)
    return TrustGrade(trust_score, "F", "Untrusted", "red")


@dataclass
class Citation:
    text: str
    citation_type: str  # url | doi | isbn | arxiv | rfc
    identifier: str
    context: str = ""


_CITATION_PATTERNS: tuple[tuple[re.Pattern[str], str], ...] = (
    (re.compile(r"\b(10\.\d{4,}/[^\s]+)\b"), "doi"),
    (re.compile(r"\b((?:978|979)[-\s]?\d[-\s]?\d{2,7}[-\s]?\d{1,7}[-\s]?\d)\b"), "isbn"),
    (re.compile(r"\b((?:arXiv:)?\d{4}\.\d{4,5}(?:v\d+)?)\b", re.I), "arxiv"),
    (re.compile(r"\b(RFC\s*\d{1,5})\b", re.I), "rfc"),
    (re.compile(r"(https?://[^\s<>\"')\]]+)"), "url"),
)


def extract_citations(text: str) -> list[Citation]:
    found: list[Citation] = []
    seen: set[str] = set()
    for pat, kind in _CITATION_PATTERNS:
        for m in pat.finditer(text):
            ident = m.group(1).strip()
            if ident in seen:
                continue
            seen.add(ident)
            ctx = text[max(0, m.start() - 50):min(len(text), m.end() + 50)].strip()
            found.append(Citation(m.group(0), kind, ident, ctx))
    return found


@dataclass
class FactCheckResult:
    claim: str
    supporting_sources: int
    contradicting_sources: int
    confidence: float
    sources: list[str] = field(default_factory=list)
    verdict: str = ""  # supported | disputed | unverified


def cross_reference_results(claim: str, results: list[Any], *, min_overlap: float = 0.3) -> FactCheckResult:
    """How many results' snippets share >= ``min_overlap`` of the claim's words."""
    import re

    tok = lambda t: set(re.findall(r"\w+", re.sub(r"</?(?:b|mark|em)>", "", t).lower()))  # noqa: E731 — ignores snippet highlight tags
    words = tok(claim)
    if not words:
        return FactCheckResult(claim, 0, 0, 0.0, verdict="unverified")
    support, urls = 0, []
    for r in results:
        have = tok(getattr(r, "snippet", "") or "")
        if len(words & have) / len(words) >= min_overlap:
            support += 1
            urls.append(r.url)
    total = len(results)
    if total == 0:
        conf, verdict = 0.0, "unverified"
    elif support >= total * 0.5:
        conf, verdict = min(support / total, 1.0), "supported"
    elif support == 0:
        conf, verdict = 0.1, "unverified"
    else:
        conf, verdict = support / total, "disputed"
    return FactCheckResult(claim, support, max(0, total - support), round(conf, 3), urls[:10], verdict)
