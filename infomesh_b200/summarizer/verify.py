"""Three-stage summary verification: (1) self-check — key-fact coverage >= 30 % and no number-contradiction,
(2) cross-validation — mean word-Jaccard >= 0.30 against peer summaries, (3) quality score / level
(reference infomesh/summarizer/verify.py:25-438)."""
from __future__ import annotations

import re
from dataclasses import dataclass, replace
from enum import StrEnum

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MIN_COVERAGE_RATIO = 0.30
MIN_KEY_FACTS = 3
MAX_KEY_FACTS = 20
MIN_CROSS_SIMILARITY = 0.30
_WORD = re.compile(r"\w+")
_NUM = re.compile(r"\b\d+(?:\.\d+)?\b")
_SENT_BREAK = re.compile(r"(?<=[.!?])\s+(?=[A-Z一-鿿])")
_INNER_CAP = re.compile(r"(?<!^)\b[A-Z][a-z]+")


class VerificationLevel(StrEnum):
    UNVERIFIED = "unverified"
    SELF_VERIFIED = "self_verified"
    CROSS_VALIDATED = "cross_validated"
    REPUTATION_TRUSTED = "reputation"


@dataclass(frozen=True)
class KeyFact:
    text: str
    source_offset: int
    found_in_summary: bool


@dataclass(frozen=True)
class SelfVerificationResult:
    key_facts: list[KeyFact]
    facts_found: int
    facts_total: int
    coverage_ratio: float
    has_contradiction: bool
    passed: bool
    detail: str


@dataclass(frozen=True)
class CrossValidationResult:
    peer_summaries: list[str]
    similarity_scores: list[float]
    avg_similarity: float
    passed: bool
    detail: str


@dataclass(frozen=True)
class VerificationReport:
    url: str
    content_hash: str
    summary: str
    level: VerificationLevel
    self_check: SelfVerificationResult | None
    cross_check: CrossValidationResult | None
    quality_score: float
    detail: str


def _split_sentences(text: str) -> list[tuple[str, int]]:
    out, start = [], 0
    for m in _SENT_BREAK.finditer(text):
        s = text[start:m.start() + 1].strip()
        if len(s) > 10:
            out.append((s, start))
        start = m.end()
    tail = text[start:].strip()
    if len(tail) > 10:
        out.append((tail, start))
    return out


def _fact_score(sentence: str) -> float:
    score = 0.0
    if re.search(r"\b\d+", sentence):
        score += 1.0
    if re.search(r"\d+%", sentence):
        score += 0.5
    score += min(1.0, 0.3 * len(_INNER_CAP.findall(sentence)))
    if 8 <= len(sentence.split()) <= 30:
        score += 0.5
    return score


def extract_key_facts(source_text: str, *, max_facts: int = MAX_KEY_FACTS) -> list[KeyFact]:
    ranked = sorted(((_fact_score(s), s, off) for s, off in _split_sentences(source_text)), key=lambda t: t[0], reverse=True)
    return [KeyFact(s.strip(), off, False) for sc, s, off in ranked[:max_facts] if sc > 0]


def check_key_facts(summary: str, key_facts: list[KeyFact]) -> list[KeyFact]:
    have = set(_WORD.findall(summary.lower()))
    out = []
    for f in key_facts:
        words = set(_WORD.findall(f.text.lower()))
        out.append(replace(f, found_in_summary=len(words & have) / len(words) >= 0.4) if words else f)
    return out


def detect_contradiction(source_text: str, summary: str) -> bool:
    """More than two numbers in the summary that never occur in the source."""
    return len(set(_NUM.findall(summary)) - set(_NUM.findall(source_text))) > 2


def self_verify(source_text: str, summary: str) -> SelfVerificationResult:
    facts = check_key_facts(summary, extract_key_facts(source_text))
    found, total = sum(f.found_in_summary for f in facts), len(facts)
    cov = found / total if total else 0.0
    contra = detect_contradiction(source_text, summary)
    low = total >= MIN_KEY_FACTS and cov < MIN_COVERAGE_RATIO
    notes = ([f"low coverage: {cov:.1%}"] if low else []) + (["contradiction detected"] if contra else [])
    return SelfVerificationResult(facts, found, total, round(cov, 4), contra, not low and not contra, "; ".join(notes) or "ok")


def compute_similarity(text_a: str, text_b: str) -> float:
    a, b = set(_WORD.findall(text_a.lower())), set(_WORD.findall(text_b.lower()))
    return len(a & b) / len(a | b) if a and b else 0.0


def cross_validate(our_summary: str, peer_summaries: list[str]) -> CrossValidationResult:
    if not peer_summaries:
        return CrossValidationResult([], [], 0.0, False, "no peer summaries available")
    scores = [compute_similarity(our_summary, p) for p in peer_summaries]
    avg = sum(scores) / len(scores)
    ok = avg >= MIN_CROSS_SIMILARITY
    return CrossValidationResult(list(peer_summaries), [round(s, 4) for s in scores], round(avg, 4), ok,
                                 f"avg_similarity={avg:.3f}" if ok else f"low similarity: {avg:.3f}")


def verify_summary(url: str, content_hash: str, source_text: str, summary: str, *,
                   peer_summaries: list[str] | None = None) -> VerificationReport:
    sc = self_verify(source_text, summary)
    cc = cross_validate(summary, peer_summaries) if peer_summaries else None
    q = (0.5 + 0.2 * sc.coverage_ratio if sc.passed else 0.0) + (0.3 * cc.avg_similarity if cc and cc.passed else 0.0)
    level = (VerificationLevel.CROSS_VALIDATED if cc and cc.passed and sc.passed
             else VerificationLevel.SELF_VERIFIED if sc.passed else VerificationLevel.UNVERIFIED)
    detail = "; ".join([f"self: {sc.detail}"] + ([f"cross: {cc.detail}"] if cc else []))
    logger.info("summary_verified", url=url, level=level.value, quality=round(min(1.0, q), 4))
    return VerificationReport(url, content_hash, summary, level, sc, cc, round(min(1.0, q), 4), detail)
