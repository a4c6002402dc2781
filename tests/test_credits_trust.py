"""credits + trust planes."""
from __future__ import annotations

import time

import pytest

from infomesh_b200.credits import farming as F
from infomesh_b200.credits import scheduling as S
from infomesh_b200.credits import timezone_verify as TZ
from infomesh_b200.credits.github_identity import format_startup_message, is_valid_email
from infomesh_b200.credits.ledger import ActionType, ContributionTier, CreditLedger, CreditState, is_off_peak
from infomesh_b200.credits.sync import CreditSummary, CreditSyncManager, CreditSyncStore
from infomesh_b200.credits.verification import CreditProofBuilder
from infomesh_b200.p2p.keys import KeyPair, ensure_keys, export_public_key, rotate_keys, verify_revocation
from infomesh_b200.trust import attestation as AT
from infomesh_b200.trust import audit as AU
from infomesh_b200.trust.detector import MaliciousNodeDetector, ThreatLevel
from infomesh_b200.trust.dmca import TakedownManager, TakedownStatus, deserialize_notice, serialize_notice, takedown_dht_key
from infomesh_b200.trust.gdpr import DeletionBasis, DeletionManager, deletion_dht_key
from infomesh_b200.trust.merkle import MerkleTree, deserialize_proof, serialize_proof, verify_root_record
from infomesh_b200.trust.reputation import LLMReputationTracker, ReputationGrade
from infomesh_b200.trust.scoring import TrustStore, TrustTier, compute_trust_score, trust_tier


def test_ledger_weights_tiers_cap_and_debt(tmp_path):
    l = CreditLedger(tmp_path / "c.db", owner_email="a@b.co")
    assert l.record_action(ActionType.CRAWL, 10) == 10.0
    assert l.record_action(ActionType.LLM_SUMMARIZE_PEER, 2, off_peak=True) == 6.0     # 2.0 * 2 * 1.5
    assert l.record_action(ActionType.CRAWL, 1, off_peak=True) == 1.0                  # multiplier only for LLM
    assert l.balance() == 17.0 and l.tier() == ContributionTier.TIER_1 and l.search_cost() == 0.100
    with pytest.raises(ValueError):
        l.record_action(ActionType.CRAWL, 0)
    l.record_action(ActionType.LLM_SUMMARIZE_OWN, 1000)                                # 1500 LLM credits vs 11 non-LLM
    assert l.contribution_score() == pytest.approx(11 + 11 * 1.5)                      # LLM capped at 60 % of the total
    l.record_action(ActionType.GIT_FIX, 1)
    assert l.tier() == ContributionTier.TIER_3 and l.search_cost() == 0.033
    assert l.spend(1e9) and l.credit_state() == CreditState.GRACE                      # never blocked
    assert 71.9 < l.grace_remaining_hours() <= 72.0
    later = time.time() + 73 * 3600
    assert l.credit_state(now=later) == CreditState.DEBT
    assert l.search_allowance(now=later).search_cost == pytest.approx(0.066)
    l.record_action(ActionType.GIT_MAJOR, 20000)
    assert l.credit_state() == CreditState.NORMAL and l.stats().owner_email == "a@b.co"
    assert l.earnings_by_action()[0][0] == "git_major" and len(l.recent_entries(limit=3)) == 3
    assert is_off_peak(hour=23) and is_off_peak(hour=3) and not is_off_peak(hour=12)


def test_credit_proof_roundtrip_and_tamper():
    kp, l = KeyPair.generate(), CreditLedger()
    for i in range(25):
        l.record_action(ActionType.CRAWL, 1 + i, key_pair=kp, note=f"n{i}")
    l.record_action(ActionType.CRAWL, 1)                                               # unsigned entry is excluded
    proof = CreditProofBuilder(l, kp).build_proof(sample_size=7, request_id="r1")
    res = CreditProofBuilder.verify_proof(proof)
    assert res.verified and res.valid_signatures == 7 and res.valid_proofs == 7 and proof["entry_count"] == 25
    bad = dict(proof, sample_entries=[dict(proof["sample_entries"][0], credits=999.0)] + proof["sample_entries"][1:])
    assert not CreditProofBuilder.verify_proof(bad).verified
    assert not CreditProofBuilder.verify_proof(proof, known_public_key=KeyPair.generate().public_key_bytes()).verified
    assert CreditProofBuilder.verify_proof(CreditProofBuilder(CreditLedger(), kp).build_proof()).detail == "empty_ledger"


def test_farming_detector():
    d = F.FarmingDetector()
    now = 1_000_000.0
    assert d.check("new", "crawl", now=now).verdict == F.FarmingVerdict.PROBATION
    d.register_node("old", now=now - 48 * 3600)
    assert d.check("old", "crawl", now=now).verdict == F.FarmingVerdict.CLEAN and d.credit_multiplier("old", now=now) == 1.0
    for i in range(12):
        d.log_action("bot", "crawl", now=now - 600 + i * 30.0)                         # perfectly regular
    d.register_node("bot", now=now - 48 * 3600)
    assert d.check("bot", "crawl", now=now).verdict == F.FarmingVerdict.SUSPICIOUS
    for i in range(40):
        d.log_action("bot", "crawl", now=now - 10 + i * 0.1)                           # burst
    d.check("bot", "crawl", now=now)
    assert d.check("bot", "crawl", now=now).verdict == F.FarmingVerdict.BLOCKED and d.credit_multiplier("bot") == 0.0
    assert len(d.get_anomaly_history("bot")) >= 3
    d.unblock("bot")
    assert not d.is_blocked("bot")
    for i in range(130):
        d.log_action("fast", "crawl", now=now - 3500 + i * (20 + (i % 7)))
    d.register_node("fast", now=now - 48 * 3600)
    assert d.is_rate_limited("fast", "crawl", now=now)


def test_scheduling_and_timezone():
    assert S.is_off_peak_at(hour=6, start=23, end=7) and not S.is_off_peak_at(hour=7, start=23, end=7)
    assert S.parse_hhmm("23:00", 1) == 23 and S.parse_hhmm("bogus", 5) == 5
    nodes = [S.NodeScheduleInfo("a", 23, 7, "Asia/Seoul", True, 0.6), S.NodeScheduleInfo("b", 23, 7, "UTC", True, 0.9),
             S.NodeScheduleInfo("c", 23, 7, "UTC", False, 1.0)]
    sch = S.EnergyAwareScheduler()
    d = sch.schedule_llm_task(nodes, now_override_hour=2)
    assert d.target_peer_id == "b" and d.is_off_peak and d.credit_multiplier == 1.5
    d = sch.schedule_llm_task(nodes, now_override_hour=12)
    assert d.target_peer_id == "b" and not d.is_off_peak and d.credit_multiplier == 1.0
    assert sch.schedule_llm_task([nodes[2]]) is None
    assert [x.target_peer_id for x in sch.schedule_batch(nodes, 3, now_override_hour=2)] == ["b", "a", "b"]
    assert TZ.verify_timezone("p", "Asia/Seoul", "211.1.2.3").plausible
    assert not TZ.verify_timezone("p", "America/New_York", "211.1.2.3").plausible
    assert TZ.verify_timezone("p", "Asia/Seoul", "250.1.1.1").estimated_offset_hours is None
    tr = TZ.TimezoneConsistencyTracker()
    for i, tz in enumerate(["UTC", "Asia/Seoul", "UTC", "Asia/Tokyo"]):
        rec = tr.record_claim("p", tz, now=1000.0 + i)
    assert rec.suspicious and rec.changes_in_24h == 3


def test_credit_sync_same_owner():
    kp1, kp2 = KeyPair.generate(), KeyPair.generate()
    l1, l2 = CreditLedger(), CreditLedger()
    l1.record_action(ActionType.CRAWL, 10)
    l2.record_action(ActionType.CRAWL, 5)
    m1 = CreditSyncManager(l1, CreditSyncStore(), "Me@Example.com", kp1, kp1.peer_id)
    m2 = CreditSyncManager(l2, CreditSyncStore(), "me@example.com ", kp2, kp2.peer_id)
    other = CreditSyncManager(CreditLedger(), CreditSyncStore(), "x@y.zz", None, "o")
    s2 = m2.build_summary()
    assert m1.owner_email_hash == m2.owner_email_hash and m1.receive_summary(s2)
    assert not m1.receive_summary(m1.build_summary()) and not other.receive_summary(s2)
    agg = m1.aggregated_stats()
    assert agg.node_count == 2 and agg.total_earned == 15.0 and agg.balance == 15.0
    forged = CreditSummary.from_dict({**s2.to_dict(), "total_earned": 1e9})
    assert not m1.receive_summary(forged)                                              # signature no longer matches
    assert not m1.receive_summary(CreditSummary.from_dict({**s2.to_dict(), "timestamp": time.time() + 3600}))
    assert not m1.needs_sync(s2.peer_id) and m1.needs_sync("unknown")
    assert is_valid_email("a.b+c@d.io") and not is_valid_email("nope") and "not connected" in format_startup_message(None)


def test_keys_roundtrip_and_rotation(tmp_path):
    kp = ensure_keys(tmp_path)
    assert len(kp.peer_id) == 40 and ensure_keys(tmp_path).peer_id == kp.peer_id
    assert "BEGIN PUBLIC KEY" in export_public_key(tmp_path)
    assert oct((tmp_path / "keys" / "private.pem").stat().st_mode & 0o777) == "0o600"
    sig = kp.sign(b"msg")
    assert kp.verify(b"msg", sig) and not kp.verify(b"other", sig) and not kp.verify(b"msg", b"x" * 64)
    old, new, rec = rotate_keys(tmp_path)
    assert old.peer_id == kp.peer_id != new.peer_id and verify_revocation(rec)
    assert ensure_keys(tmp_path).peer_id == new.peer_id and list((tmp_path / "keys" / "revocations").iterdir())
    from dataclasses import replace
    assert not verify_revocation(replace(rec, new_peer_id="0" * 40))


def test_merkle_tree_proofs():
    hashes = [f"{i:064x}" for i in range(7)]
    t = MerkleTree()
    root = t.build(hashes)
    assert t.leaf_count == 7 and t.height == 4
    for i, h in enumerate(hashes):
        p = t.get_proof(i)
        assert MerkleTree.verify_proof(p) and MerkleTree.verify_document(h, p)
        assert MerkleTree.verify_proof(deserialize_proof(serialize_proof(p)))
    assert not MerkleTree.verify_document("f" * 64, t.get_proof(0))
    t2 = MerkleTree(); t2.build(hashes[:-1] + ["e" * 64])
    assert t2.root_hash != root
    with pytest.raises(ValueError):
        MerkleTree().build([])
    with pytest.raises(IndexError):
        t.get_proof(7)
    kp = KeyPair.generate()
    rec = t.create_root_record(kp.peer_id, kp)
    assert verify_root_record(rec, kp.public_key_bytes()) and not verify_root_record(rec, KeyPair.generate().public_key_bytes())


def test_attestation_and_audit_flow():
    kp = KeyPair.generate()
    att = AT.create_attestation("https://x/1", b"<html>raw</html>", "the text", kp)
    assert AT.verify_attestation(att, kp, raw_body=b"<html>raw</html>", extracted_text="the text").verified
    assert AT.verify_attestation(att, kp, extracted_text="changed").detail == "text_hash_mismatch"
    assert AT.verify_attestation(att, KeyPair.generate()).detail == "signature_invalid"
    assert AT.deserialize_attestation(AT.serialize_attestation(att)) == att
    assert AT.verify_attestation_with_key(att, kp.public_key_bytes())
    sch = AU.AuditScheduler()
    assert sch.should_schedule() and sch.create_audit("t", "u", "h", "r", ["a", "b"]) is None
    req = sch.create_audit("t", "https://x/1", att.text_hash, att.raw_hash, ["t", "a", "b", "c", "d"], now=100.0)
    assert len(req.auditor_peer_ids) == 3 and "t" not in req.auditor_peer_ids and not sch.should_schedule(now=200.0)
    a, b, c = req.auditor_peer_ids
    mk = lambda pid, text: AU.perform_audit_check(req.url, req.expected_text_hash, req.expected_raw_hash,
                                                  actual_text=text, auditor_peer_id=pid, audit_id=req.audit_id, target_peer_id="t")
    assert sch.submit_result(mk(a, "the text")) is None and sch.submit_result(mk(a, "the text")) is None   # duplicate ignored
    assert sch.submit_result(mk("stranger", "the text")) is None
    sch.submit_result(mk(b, "the text"))
    summ = sch.submit_result(mk(c, "something else"))
    assert summ.final_verdict == AU.AuditVerdict.PASS and summ.pass_count == 2 and summ.suspicious_auditors == [c]
    ts = TrustStore()
    AU.apply_audit_summary(ts, summ)
    assert ts.get_trust("t").audit_pass_rate == 1.0 and ts.get_trust(c).consecutive_audit_failures == 1
    assert AU.perform_audit_check("u", "h", "r").verdict == AU.AuditVerdict.ERROR
    t = MerkleTree(); t.build(["aa", "bb", "cc"])
    assert AU.perform_merkle_audit("bb", t.get_proof(1), t.root_hash).verdict == AU.AuditVerdict.PASS
    assert AU.perform_merkle_audit("zz", t.get_proof(1), t.root_hash).verdict == AU.AuditVerdict.FAIL
    assert AU.perform_merkle_audit("bb", t.get_proof(1), "0" * 64).detail.startswith("merkle_root_mismatch")


def test_trust_scoring_isolation_reputation_detector():
    assert compute_trust_score(720, 5000, 10, 10, 1.0, True) == pytest.approx(1.0)
    assert compute_trust_score(0, 0, 0, 0, 0.0) == pytest.approx(0.4 * 0.5 + 0.2 * 0.5)
    assert trust_tier(0.85) == TrustTier.TRUSTED and trust_tier(0.1) == TrustTier.UNTRUSTED
    rep = LLMReputationTracker()
    ts = TrustStore(reputation_tracker=rep)
    ts.update_uptime("p", 360); ts.update_contribution("p", 2500)
    for _ in range(6):
        ts.record_audit("p", passed=True); ts.record_summary_rating("p", 0.9)
    pt = ts.get_trust("p")
    assert pt.trust_score == pytest.approx(0.15 * 0.5 + 0.25 * 0.5 + 0.4 + 0.2 * 0.9) and pt.tier == TrustTier.NORMAL
    assert rep.get_reputation("p").grade in (ReputationGrade.GOOD, ReputationGrade.EXCELLENT) and rep.best_peers() == ["p"]
    assert rep.get_reputation("nobody") is None and rep.get_quality_score("nobody") == 0.5
    for _ in range(3):
        ts.record_audit("bad", passed=False)
    assert ts.is_isolated("bad") and [p.peer_id for p in ts.list_isolated()] == ["bad"] and ts.get_trust_score("unknown") == 0.5
    ts.unisolate("bad")
    det = MaliciousNodeDetector(ts, F.FarmingDetector())
    assert det.assess("p").threat_level in (ThreatLevel.NONE, ThreatLevel.LOW)
    ts.record_audit("meh", passed=False); ts.record_audit("meh", passed=False)
    a = det.assess_and_enforce("meh")
    assert a.should_isolate and ts.is_isolated("meh") and det.assess("meh").threat_level == ThreatLevel.ISOLATED


def test_dmca_and_gdpr_persist(tmp_path):
    kp = KeyPair.generate()
    tm = TakedownManager(str(tmp_path / "dmca.db"))
    n = tm.create_notice("https://x/bad", "copyright", kp, now=1000.0)
    assert tm.verify_notice(n, kp) and tm.is_taken_down("https://x/bad") and n.deadline == 1000.0 + 86400
    assert tm.check_compliance(n.notice_id, "me", now=1001.0) == TakedownStatus.PENDING
    assert tm.check_compliance(n.notice_id, "me", now=1e9) == TakedownStatus.EXPIRED
    tm.acknowledge(n.notice_id, "me"); tm.mark_complied(n.notice_id, "me"); tm.record_propagation(n.notice_id, "peerX")
    assert tm.check_compliance(n.notice_id, "me") == TakedownStatus.COMPLIED and tm.list_non_compliant("me") == []
    assert len(tm.list_non_compliant("other")) == 1 and deserialize_notice(serialize_notice(n)) == n
    tm.close()
    tm2 = TakedownManager(str(tmp_path / "dmca.db"))                                   # survives a restart
    assert tm2.is_taken_down("https://x/bad") and tm2.get_record(n.notice_id).propagated_to == ["peerX"]
    assert takedown_dht_key("https://x/bad").startswith("/infomesh/takedown/")
    with pytest.raises(ValueError):
        for i in range(11):
            tm2.create_notice(f"https://x/{i}", "r", kp, now=2000.0 + i)
    other = TakedownManager()
    assert other.receive_notice(n, kp) and not TakedownManager().receive_notice(n, KeyPair.generate())
    dm = DeletionManager(str(tmp_path / "gdpr.db"))
    r = dm.create_request("https://x/pii", DeletionBasis.RIGHT_TO_ERASURE, "remove me", kp, personal_data_fields=["name"])
    assert dm.is_blocked("https://x/pii") and dm.verify_request(r, kp) and dm.list_pending("me") == [r]
    dm.confirm_deletion(r.request_id, "me")
    assert dm.list_pending("me") == [] and deletion_dht_key("u").startswith("/infomesh/gdpr/")
    dm.close()
    dm2 = DeletionManager(str(tmp_path / "gdpr.db"))
    assert dm2.is_blocked("https://x/pii") and dm2.blocklist_size == 1 and dm2.get_request_for_url("https://x/pii") == r
    peer = DeletionManager()
    assert not peer.receive_request(r) and not peer.receive_request(r, KeyPair.generate()) and peer.receive_request(r, kp)
    assert dm2.unblock("https://x/pii", admin_key=kp) and not dm2.is_blocked("https://x/pii")
