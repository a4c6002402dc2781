"""Malicious-node detector combining trust and farming signals (reference infomesh/trust/detector.py:25-209):
farming-blocked or UNTRUSTED => HIGH; two or more weak signals => MEDIUM; both recommend isolation."""
from __future__ import annotations

from dataclasses import dataclass
from enum import StrEnum

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


class MaliciousNodeDetector:
    def __init__(self, trust_store: TrustStore, farming_detector: FarmingDetector):
        self._trust = trust_store
        self._farming = farming_detector

    def assess(self, peer_id: str, *, action: str = "crawl") -> ThreatAssessment:
        pt = self._trust.get_trust(peer_id)
        score = pt.trust_score if pt else 0.5
        tier = pt.tier if pt else TrustTier.NORMAL
        fails = pt.consecutive_audit_failures if pt else 0
        fc = self._farming.check(peer_id, action)
        mk = lambda level, signals, isolate, detail: ThreatAssessment(  # noqa: E731
            peer_id, level, score, tier, fc.verdict, fails, fc.anomaly_count, signals, isolate, detail)
        if pt is not None and pt.isolated:
            return mk(ThreatLevel.ISOLATED, [], True, "already isolated")
        weak = []
        if fails >= WEAK_SIGNAL_AUDIT_FAILURES:
            weak.append(f"audit_failures={fails}")
        if fc.anomaly_count >= WEAK_SIGNAL_ANOMALY_COUNT:
            weak.append(f"anomalies={fc.anomaly_count}")
        if score < WEAK_SIGNAL_TRUST_THRESHOLD:
            weak.append(f"low_trust={score:.3f}")
        if fc.verdict == FarmingVerdict.BLOCKED:
            weak.append("farming_blocked")
        if fc.rate_limit_exceeded:
            weak.append("rate_limited")
        if fc.verdict == FarmingVerdict.BLOCKED:
            out = mk(ThreatLevel.HIGH, weak, True, "blocked for credit farming")
        elif tier == TrustTier.UNTRUSTED:
            out = mk(ThreatLevel.HIGH, weak, True, f"untrusted peer (score={score:.3f})")
        elif len(weak) >= WEAK_SIGNAL_ISOLATION_COUNT:
            out = mk(ThreatLevel.MEDIUM, weak, True, f"multiple weak signals: {', '.join(weak)}")
        elif len(weak) == 1:
            out = mk(ThreatLevel.LOW, weak, False, f"single weak signal: {weak[0]}")
        else:
            out = mk(ThreatLevel.NONE, weak, False, "no threats detected")
        if out.should_isolate:
            logger.warning("malicious_node_detected", peer_id=peer_id[:12], threat_level=out.threat_level.value,
                           detail=out.detail)
        return out

    def assess_and_enforce(self, peer_id: str, *, action: str = "crawl") -> ThreatAssessment:
        a = self.assess(peer_id, action=action)
        if a.should_isolate and a.threat_level != ThreatLevel.ISOLATED:
            self._trust.isolate_peer(peer_id)
        return a
