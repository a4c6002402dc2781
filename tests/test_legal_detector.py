"""trust/dmca.py, trust/gdpr.py (signed notices, propagation, compliance, persistence) and trust/detector.py."""
import dataclasses

import pytest

from infomesh_b200.credits.farming import FarmingDetector, FarmingVerdict
from infomesh_b200.p2p.keys import KeyPair
from infomesh_b200.trust import dmca as DM
from infomesh_b200.trust import gdpr as GD
from infomesh_b200.trust.detector import MaliciousNodeDetector, ThreatLevel
from infomesh_b200.trust.scoring import TrustStore


# ------------------------------------------------------------------ DMCA
def test_takedown_create_verify_receive():
    kp, other = KeyPair.generate(), KeyPair.generate()
    m = DM.TakedownManager()
    n = m.create_notice("https://ex.org/p", "copyright " * 500, kp, contact_info="legal@ex.org", now=1000.0)
    assert len(n.reason) <= DM.MAX_NOTICE_LENGTH and n.deadline == 1000.0 + DM.COMPLIANCE_DEADLINE_HOURS * 3600
    assert m.verify_notice(n, kp) and not m.verify_notice(n, other) and m.is_taken_down("https://ex.org/p")
    peer = DM.TakedownManager()
    assert not peer.receive_notice(n, None) and not peer.receive_notice(n, other) and not peer.is_taken_down(n.url)
    assert peer.receive_notice(DM.deserialize_notice(DM.serialize_notice(n)), kp) and peer.get_notice_for_url(n.url).notice_id == n.notice_id
    tampered = dataclasses.replace(n, url="https://victim.org/")
    assert not peer.receive_notice(tampered, kp)


def test_takedown_compliance_lifecycle():
    kp = KeyPair.generate()
    m = DM.TakedownManager()
    n = m.create_notice("https://ex.org/p", "r", kp, now=0.0)
    assert m.check_compliance(n.notice_id, "peerA", now=10.0) == DM.TakedownStatus.PENDING
    assert m.check_compliance(n.notice_id, "peerA", now=n.deadline + 1) == DM.TakedownStatus.EXPIRED
    assert m.check_compliance("nope", "peerA") == DM.TakedownStatus.INVALID and m.acknowledge("nope", "x") is None
    m.acknowledge(n.notice_id, "peerA", now=20.0)
    assert m.check_compliance(n.notice_id, "peerA") == DM.TakedownStatus.ACKNOWLEDGED
    assert [x.notice_id for x in m.list_non_compliant("peerA")] == [n.notice_id]
    ack = m.mark_complied(n.notice_id, "peerA", now=30.0)
    assert ack.complied_at == 30.0 and m.check_compliance(n.notice_id, "peerA") == DM.TakedownStatus.COMPLIED
    assert m.list_non_compliant("peerA") == [] and len(m.list_non_compliant("peerB")) == 1
    m.record_propagation(n.notice_id, "peerB")
    m.record_propagation(n.notice_id, "peerB")
    assert m.get_record(n.notice_id).propagated_to == ["peerB"] and len(m.list_active()) == 1


def test_takedown_rate_limit_and_persistence(tmp_path):
    kp = KeyPair.generate()
    db = str(tmp_path / "dmca.db")
    m = DM.TakedownManager(db)
    for i in range(DM.TakedownManager.MAX_NOTICES_PER_HOUR):
        m.create_notice(f"https://ex.org/{i}", "r", kp, now=100.0 + i)
    with pytest.raises(ValueError):
        m.create_notice("https://ex.org/more", "r", kp, now=200.0)
    last = m.create_notice("https://ex.org/later", "r", kp, now=100.0 + 3700)
    m.mark_complied(last.notice_id, "me")
    m.close()
    again = DM.TakedownManager(db)
    assert again.is_taken_down("https://ex.org/3") and again.check_compliance(last.notice_id, "me") == DM.TakedownStatus.COMPLIED
    assert DM.takedown_dht_key("https://a") != DM.takedown_dht_key("https://b")
    again.close()


# ------------------------------------------------------------------ GDPR
def test_deletion_request_flow_and_blocklist():
    kp, other = KeyPair.generate(), KeyPair.generate()
    m = GD.DeletionManager()
    r = m.create_request("https://ex.org/me", GD.DeletionBasis.RIGHT_TO_ERASURE, "please " * 1000, kp, personal_data_fields=["name"])
    assert m.verify_request(r, kp) and not m.verify_request(r, other) and m.is_blocked(r.url) and m.blocklist_size == 1
    assert len(r.reason) <= GD.MAX_REASON_LENGTH
    peer = GD.DeletionManager()
    assert not peer.receive_request(r) and not peer.receive_request(r, other)
    assert peer.receive_request(GD.deserialize_request(GD.serialize_request(r)), kp) and peer.is_blocked(r.url)
    assert [x.request_id for x in peer.list_pending("peer")] == [r.request_id]
    conf = peer.confirm_deletion(r.request_id, "peer", now=5.0)
    assert conf.status == GD.DeletionStatus.DELETED and peer.list_pending("peer") == [] and len(peer.list_pending("other")) == 1
    assert peer.confirm_deletion("nope", "peer") is None
    peer.record_propagation(r.request_id, "p2")
    peer.record_propagation(r.request_id, "p2")
    assert peer.get_record(r.request_id).propagated_to == ["p2"] and len(peer.list_all()) == 1


def test_deletion_unblock_and_persistence(tmp_path):
    kp = KeyPair.generate()
    db = str(tmp_path / "gdpr.db")
    m = GD.DeletionManager(db)
    r1 = m.create_request("https://ex.org/a", GD.DeletionBasis.RIGHT_TO_ERASURE, "x", kp)
    m.create_request("https://ex.org/b", GD.DeletionBasis.RIGHT_TO_ERASURE, "y", kp)
    assert m.unblock("https://ex.org/a", admin_key=kp) and not m.unblock("https://ex.org/a", admin_key=kp)
    assert not m.is_blocked("https://ex.org/a") and m.is_blocked("https://ex.org/b")
    m.close()
    again = GD.DeletionManager(db)
    assert again.is_blocked("https://ex.org/b") and not again.is_blocked("https://ex.org/a") and again.get_record(r1.request_id) is not None
    assert GD.deletion_dht_key("https://a") != GD.deletion_dht_key("https://b")
    again.close()


# ------------------------------------------------------------------ malicious node detector
def test_detector_levels():
    ts, fd = TrustStore(), FarmingDetector()
    det = MaliciousNodeDetector(ts, fd)
    old = 1.0
    fd.register_node("clean", now=old)
    a = det.assess("clean")
    assert a.threat_level == ThreatLevel.NONE and not a.should_isolate and a.farming_verdict == FarmingVerdict.CLEAN
    fd.register_node("shaky", now=old)
    ts.update_uptime("shaky", 720)
    ts.update_contribution("shaky", 5000)
    for _ in range(8):
        ts.record_audit("shaky", passed=True)
    ts.record_audit("shaky", passed=False)
    ts.record_audit("shaky", passed=False)
    low = det.assess("shaky")
    assert low.threat_level == ThreatLevel.LOW and low.weak_signals == ["audit_failures=2"] and not low.should_isolate


def test_detector_enforces_isolation_for_blocked_and_untrusted_peers():
    ts, fd = TrustStore(), FarmingDetector()
    det = MaliciousNodeDetector(ts, fd)
    fd.register_node("farmer", now=1.0)
    for _ in range(3):
        fd.record_anomaly("farmer", "burst")
    a = det.assess_and_enforce("farmer")
    assert a.threat_level == ThreatLevel.HIGH and a.should_isolate and ts.is_isolated("farmer")
    assert det.assess("farmer").threat_level == ThreatLevel.ISOLATED
    fd.register_node("liar", now=1.0)
    for _ in range(2):
        ts.record_audit("liar", passed=False)
    ts.record_summary_rating("liar", 0.0)
    b = det.assess("liar")
    assert b.trust_score < 0.3 and b.threat_level == ThreatLevel.HIGH and "untrusted" in b.detail
