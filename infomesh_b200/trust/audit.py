"""Random re-crawl audits: ~1 / hour, 3 independent auditors, majority of 2; auditors report the hashes they saw so
dishonest auditors are detectable; Merkle-proof audits avoid the re-crawl (reference infomesh/trust/audit.py:25-524)."""
from __future__ import annotations

import random
import time
from collections import Counter
from dataclasses import dataclass, field
from enum import StrEnum
from typing import Any

from infomesh_b200.hashing import content_hash, short_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

AUDITS_PER_HOUR: float = 1.0
AUDIT_NODES_PER_CHECK: int = 3
AUDIT_MAJORITY: int = 2
NEW_NODE_PROBATION_HOURS: float = 24.0
PROBATION_AUDIT_MULTIPLIER: float = 3.0
MIN_AUDITOR_AGE_HOURS: float = 24.0
AUDIT_TIMEOUT_SECONDS: float = 30.0


class AuditVerdict(StrEnum):
    PASS = "pass"
    FAIL = "fail"
    ERROR = "error"
    INCONCLUSIVE = "inconclusive"


@dataclass(frozen=True)
class AuditRequest:
    audit_id: str
    target_peer_id: str
    url: str
    expected_text_hash: str
    expected_raw_hash: str
    requested_at: float
    auditor_peer_ids: list[str] = field(default_factory=list)


@dataclass(frozen=True)
class AuditResult:
    audit_id: str
    auditor_peer_id: str
    target_peer_id: str
    url: str
    actual_text_hash: str | None
    actual_raw_hash: str | None
    verdict: AuditVerdict
    detail: str
    completed_at: float
    auditor_signature: bytes = b""


@dataclass(frozen=True)
class AuditSummary:
    audit_id: str
    target_peer_id: str
    url: str
    results: list[AuditResult]
    final_verdict: AuditVerdict
    pass_count: int
    fail_count: int
    error_count: int
    suspicious_auditors: list[str] = field(default_factory=list)


def _generate_audit_id(peer_id: str, url: str, timestamp: float) -> str:
    return short_hash(f"{peer_id}|{url}|{timestamp}".encode(), length=24)


def _cross_validate_auditor_hashes(results: list[AuditResult]) -> list[str]:
    """Auditors whose reported text hash differs from the majority hash (needs a strict majority to judge)."""
    seen = [(r.auditor_peer_id, r.actual_text_hash) for r in results
            if r.verdict != AuditVerdict.ERROR and r.actual_text_hash]
    if len(seen) < 2:
        return []
    top, votes = Counter(h for _, h in seen).most_common(1)[0]
    if votes <= len(seen) / 2:
        return []
    return [pid for pid, h in seen if h != top]


class AuditScheduler:
    def __init__(self):
        self._pending: dict[str, AuditRequest] = {}
        self._results: dict[str, list[AuditResult]] = {}
        self._completed: list[AuditSummary] = []
        self._last = 0.0

    def should_schedule(self, *, now: float | None = None, on_probation: bool = False) -> bool:
        rate = AUDITS_PER_HOUR * (PROBATION_AUDIT_MULTIPLIER if on_probation else 1.0)
        return (now or time.time()) - self._last >= 3600.0 / rate

    def create_audit(self, target_peer_id: str, url: str, expected_text_hash: str, expected_raw_hash: str,
                     available_auditors: list[str], *, now: float | None = None) -> AuditRequest | None:
        pool = [p for p in available_auditors if p != target_peer_id]
        if len(pool) < AUDIT_NODES_PER_CHECK:
            logger.warning("audit_insufficient_auditors", target=target_peer_id, available=len(pool))
            return None
        now = now or time.time()
        req = AuditRequest(_generate_audit_id(target_peer_id, url, now), target_peer_id, url, expected_text_hash,
                           expected_raw_hash, now, random.sample(pool, AUDIT_NODES_PER_CHECK))
        self._pending[req.audit_id] = req
        self._last = now
        return req

    def submit_result(self, result: AuditResult) -> AuditSummary | None:
        req = self._pending.get(result.audit_id)
        if req is None:
            logger.warning("audit_unknown", audit_id=result.audit_id)
            return None
        if result.auditor_peer_id not in req.auditor_peer_ids:
            logger.warning("audit_unassigned_auditor", audit_id=result.audit_id, auditor=result.auditor_peer_id[:12])
            return None
        got = self._results.setdefault(result.audit_id, [])
        if any(r.auditor_peer_id == result.auditor_peer_id for r in got):
            return None
        got.append(result)
        if len(got) < AUDIT_NODES_PER_CHECK:
            return None
        return self._finalize(result.audit_id)

    def _finalize(self, audit_id: str) -> AuditSummary:
        results = self._results.pop(audit_id, [])
        req = self._pending.pop(audit_id)
        n = Counter(r.verdict for r in results)
        if n[AuditVerdict.FAIL] >= AUDIT_MAJORITY:
            final = AuditVerdict.FAIL
        elif n[AuditVerdict.PASS] >= AUDIT_MAJORITY:
            final = AuditVerdict.PASS
        elif n[AuditVerdict.ERROR] >= AUDIT_MAJORITY:
            final = AuditVerdict.ERROR
        else:
            final = AuditVerdict.INCONCLUSIVE
        summary = AuditSummary(audit_id, req.target_peer_id, req.url, results, final, n[AuditVerdict.PASS],
                               n[AuditVerdict.FAIL], n[AuditVerdict.ERROR], _cross_validate_auditor_hashes(results))
        self._completed.append(summary)
        logger.info("audit_completed", audit_id=audit_id, verdict=final.value)
        return summary

    def expire_stale(self, *, now: float | None = None, timeout: float = AUDIT_TIMEOUT_SECONDS * 10) -> list[AuditSummary]:
        """Close audits whose auditors never answered (missing answers count as ERROR)."""
        now = now or time.time()
        out = []
        for aid, req in list(self._pending.items()):
            if now - req.requested_at < timeout:
                continue
            have = {r.auditor_peer_id for r in self._results.get(aid, [])}
            for pid in req.auditor_peer_ids:
                if pid not in have:
                    self._results.setdefault(aid, []).append(AuditResult(
                        aid, pid, req.target_peer_id, req.url, None, None, AuditVerdict.ERROR, "timeout", now))
            out.append(self._finalize(aid))
        return out

    @property
    def pending_count(self) -> int:
        return len(self._pending)

    @property
    def completed_audits(self) -> list[AuditSummary]:
        return list(self._completed)


def apply_audit_summary(trust_store: Any, summary: AuditSummary) -> None:
    """Feed an audit outcome into the trust store (target pass/fail; lying auditors fail too)."""
    if summary.final_verdict in (AuditVerdict.PASS, AuditVerdict.FAIL):
        trust_store.record_audit(summary.target_peer_id, passed=summary.final_verdict == AuditVerdict.PASS)
    for pid in summary.suspicious_auditors:
        trust_store.record_audit(pid, passed=False)


def perform_audit_check(url: str, expected_text_hash: str, expected_raw_hash: str, *,
                        actual_raw_body: bytes | None = None, actual_text: str | None = None,
                        auditor_peer_id: str = "", audit_id: str = "", target_peer_id: str = "") -> AuditResult:
    now = time.time()
    mk = lambda th, rh, v, d: AuditResult(audit_id, auditor_peer_id, target_peer_id, url, th, rh, v, d, now)  # noqa: E731
    if actual_text is None and actual_raw_body is None:
        return mk(None, None, AuditVerdict.ERROR, "no content available for verification")
    th = content_hash(actual_text) if actual_text else None
    rh = content_hash(actual_raw_body) if actual_raw_body else None
    bad = [n for n, ok in (("text_hash_mismatch", th is None or th == expected_text_hash),
                           ("raw_hash_mismatch", rh is None or rh == expected_raw_hash)) if not ok]
    return mk(th, rh, AuditVerdict.FAIL if bad else AuditVerdict.PASS, "; ".join(bad) or "content matches attestation")


def perform_merkle_audit(document_hash: str, proof: Any, expected_root_hash: str, *, auditor_peer_id: str = "",
                         audit_id: str = "", target_peer_id: str = "", url: str = "") -> AuditResult:
    from infomesh_b200.trust.merkle import MerkleTree

    now = time.time()
    mk = lambda th, v, d: AuditResult(audit_id, auditor_peer_id, target_peer_id, url, th, None, v, d, now)  # noqa: E731
    try:
        if proof.root_hash != expected_root_hash:
            return mk(None, AuditVerdict.FAIL, f"merkle_root_mismatch: proof_root={proof.root_hash[:16]}... "
                                               f"expected={expected_root_hash[:16]}...")
        if not MerkleTree.verify_document(document_hash, proof):
            return mk(document_hash, AuditVerdict.FAIL, "merkle_proof_invalid: document not in tree")
        return mk(document_hash, AuditVerdict.PASS, "merkle_proof_valid")
    except Exception as exc:  # noqa: BLE001
        return mk(None, AuditVerdict.ERROR, f"merkle_audit_error: {exc}")


def audit_result_canonical(result: AuditResult) -> bytes:
    """Bytes an auditor signs: id, both parties, URL, the evidence hashes, verdict and completion time — so a verdict cannot
    be changed after it was signed (reference infomesh/trust/audit.py:507-524)."""
    return "|".join([result.audit_id, result.auditor_peer_id, result.target_peer_id, result.url, result.actual_text_hash or "",
                     result.actual_raw_hash or "", result.verdict.value, f"{result.completed_at:.6f}"]).encode()


def sign_audit_result(result: AuditResult, key_pair: Any) -> AuditResult:
    from dataclasses import replace

    return replace(result, auditor_signature=key_pair.sign(audit_result_canonical(result)))


def verify_audit_result(result: AuditResult, public_key: bytes) -> bool:
    from infomesh_b200.p2p.keys import verify_with_public_key

    return bool(result.auditor_signature) and verify_with_public_key(public_key, audit_result_canonical(result), result.auditor_signature)
