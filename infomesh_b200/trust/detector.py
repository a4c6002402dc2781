"""Deciding whether a peer should be cut off, from what the trust store and the farming detector know about it.

Contract (SURVEY §2.1 trust/ "detector"; reference infomesh/trust/detector.py): weak signals are >= 2 consecutive audit
failures, >= 1 farming anomaly, a trust score below 0.5, a farming block, and an exceeded rate limit.  Verdicts, first
match wins: already isolated -> ISOLATED; farming block -> HIGH; UNTRUSTED tier -> HIGH; two or more weak signals ->
MEDIUM; exactly one -> LOW; none -> NONE.  HIGH and MEDIUM recommend isolation (ISOLATED reports it as already done);
``assess_and_enforce`` carries the recommendation out.  A peer the trust store has never seen counts as score 0.5, tier
NORMAL, no failures.

Implementation: the observations about a peer are gathered once into a ``_Evidence`` record; weak signals are a table of
(predicate, label) probes over it; verdicts are an ordered rule table -- adding a signal or a rule is one more row."""
from __future__ import annotations

from dataclasses import dataclass
from enum import StrEnum
from typing import Callable

from infomesh_b200.credits.farming import FarmingDetector, FarmingVerdict
from infomesh_b200.trust.scoring import TrustStore, TrustTier
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

WEAK_SIGNAL_AUDIT_FAILURES: int = 2
WEAK_SIGNAL_ANOMALY_COUNT: int = 1
WEAK_SIGNAL_TRUST_THRESHOLD: float = 0.5
WEAK_SIGNAL_ISOLATION_COUNT: int = 2


class ThreatLevel(StrEnum):
    NONE = "none"
    LOW = "low"
    MEDIUM = "medium"
    HIGH = "high"
    ISOLATED = "isolated"


@dataclass(frozen=True)
class ThreatAssessment:
    peer_id: str
    threat_level: ThreatLevel
    trust_score: float
    trust_tier: TrustTier
    farming_verdict: FarmingVerdict
    consecutive_audit_failures: int
    anomaly_count: int
    weak_signals: list[str]
    should_isolate: bool
    detail: str


@dataclass(frozen=True)
class _Evidence:
    score: float
    tier: TrustTier
    audit_failures: int
    isolated: bool
    farming: FarmingVerdict
    anomalies: int
    rate_limited: bool

    @property
    def farming_blocked(self) -> bool:
        return self.farming == FarmingVerdict.BLOCKED


# (does the signal fire?, how it is reported)
_PROBES: tuple[tuple[Callable[[_Evidence], bool], Callable[[_Evidence], str]], ...] = (
    (lambda e: e.audit_failures >= WEAK_SIGNAL_AUDIT_FAILURES, lambda e: f"audit_failures={e.audit_failures}"),
    (lambda e: e.anomalies >= WEAK_SIGNAL_ANOMALY_COUNT, lambda e: f"anomalies={e.anomalies}"),
    (lambda e: e.score < WEAK_SIGNAL_TRUST_THRESHOLD, lambda e: f"low_trust={e.score:.3f}"),
    (lambda e: e.farming_blocked, lambda e: "farming_blocked"),
    (lambda e: e.rate_limited, lambda e: "rate_limited"),
)

# (applies?, level, isolate?, explanation) -- evaluated top-down
_RULES: tuple[tuple[Callable[[_Evidence, list[str]], bool], ThreatLevel, bool, Callable[[_Evidence, list[str]], str]], ...] = (
    (lambda e, w: e.farming_blocked, ThreatLevel.HIGH, True, lambda e, w: "blocked for credit farming"),
    (lambda e, w: e.tier == TrustTier.UNTRUSTED, ThreatLevel.HIGH, True, lambda e, w: f"untrusted peer (score={e.score:.3f})"),
    (lambda e, w: len(w) >= WEAK_SIGNAL_ISOLATION_COUNT, ThreatLevel.MEDIUM, True, lambda e, w: f"multiple weak signals: {', '.join(w)}"),
    (lambda e, w: len(w) == 1, ThreatLevel.LOW, False, lambda e, w: f"single weak signal: {w[0]}"),
    (lambda e, w: True, ThreatLevel.NONE, False, lambda e, w: "no threats detected"),
)


class MaliciousNodeDetector:
    def __init__(self, trust_store: TrustStore, farming_detector: FarmingDetector):
        self._trust = trust_store
        self._farming = farming_detector

    def _collect(self, peer_id: str, action: str) -> _Evidence:
        known = self._trust.get_trust(peer_id)
        farm = self._farming.check(peer_id, action)
        if known is None:
            score, tier, failures, isolated = 0.5, TrustTier.NORMAL, 0, False
        else:
            score, tier, failures, isolated = known.trust_score, known.tier, known.consecutive_audit_failures, bool(known.isolated)
        return _Evidence(score, tier, failures, isolated, farm.verdict, farm.anomaly_count, bool(farm.rate_limit_exceeded))

    def assess(self, peer_id: str, *, action: str = "crawl") -> ThreatAssessment:
        seen = self._collect(peer_id, action)

        def verdict(level: ThreatLevel, signals: list[str], isolate: bool, why: str) -> ThreatAssessment:
            return ThreatAssessment(peer_id=peer_id, threat_level=level, trust_score=seen.score, trust_tier=seen.tier,
                                    farming_verdict=seen.farming, consecutive_audit_failures=seen.audit_failures,
                                    anomaly_count=seen.anomalies, weak_signals=signals, should_isolate=isolate, detail=why)

        if seen.isolated:
            return verdict(ThreatLevel.ISOLATED, [], True, "already isolated")
        signals = [describe(seen) for fires, describe in _PROBES if fires(seen)]
        level, isolate, explain = next((lv, iso, ex) for applies, lv, iso, ex in _RULES if applies(seen, signals))
        result = verdict(level, signals, isolate, explain(seen, signals))
        if isolate:
            logger.warning("malicious_node_detected", peer_id=peer_id[:12], threat_level=level.value, detail=result.detail)
        return result

    def assess_and_enforce(self, peer_id: str, *, action: str = "crawl") -> ThreatAssessment:
        outcome = self.assess(peer_id, action=action)
        if outcome.should_isolate and outcome.threat_level is not ThreatLevel.ISOLATED:
            self._trust.isolate_peer(peer_id)
        return outcome
