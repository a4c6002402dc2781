"""trust/scoring.py, trust/merkle.py, trust/attestation.py, trust/reputation.py — in-memory SQLite, real Ed25519 keys."""
import dataclasses

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.p2p.keys import KeyPair
from infomesh_b200.trust import attestation as AT
from infomesh_b200.trust import merkle as MK
from infomesh_b200.trust import reputation as RP
from infomesh_b200.trust import scoring as SC


# ------------------------------------------------------------------ scoring
def test_trust_score_formula_and_neutral_defaults():
    assert SC.compute_trust_score(0, 0, 0, 0, 0.0) == pytest.approx(0.40 * 0.5 + 0.20 * 0.5)
    full = SC.compute_trust_score(10_000, 1e9, 10, 10, 1.0, has_summary_data=True)
    assert full == pytest.approx(1.0)
    assert SC.compute_trust_score(360, 2500, 4, 3, 0.0, has_summary_data=True) == pytest.approx(0.15 * 0.5 + 0.25 * 0.5 + 0.40 * 0.75)
    assert SC.compute_trust_score(0, 0, 0, 0, 0.9, has_summary_data=False) == SC.compute_trust_score(0, 0, 0, 0, 0.0)


def test_trust_tiers():
    assert [SC.trust_tier(s) for s in (0.95, 0.8, 0.79, 0.5, 0.3, 0.29, -1)] == [
        SC.TrustTier.TRUSTED, SC.TrustTier.TRUSTED, SC.TrustTier.NORMAL, SC.TrustTier.NORMAL, SC.TrustTier.SUSPECT,
        SC.TrustTier.UNTRUSTED, SC.TrustTier.UNTRUSTED]


def test_trust_store_updates_and_events():
    ts = SC.TrustStore()
    assert ts.get_trust("p") is None and ts.get_trust_score("p") == 0.5
    ts.update_uptime("p", 720)
    ts.update_contribution("p", 5000)
    ts.record_audit("p", passed=True)
    ts.record_summary_rating("p", 1.7)                       # clamped to 1.0
    t = ts.get_trust("p")
    assert (t.uptime_score, t.contribution_score, t.audit_pass_rate, t.summary_quality) == (1.0, 1.0, 1.0, 1.0)
    assert t.trust_score == 1.0 and t.tier == SC.TrustTier.TRUSTED
    assert [e.field for e in ts.recent_events("p")] == ["summary", "audit", "contribution", "uptime"]


def test_three_consecutive_audit_failures_isolate_and_pass_resets():
    ts = SC.TrustStore()
    ts.record_audit("p", passed=False)
    ts.record_audit("p", passed=False)
    ts.record_audit("p", passed=True)
    ts.record_audit("p", passed=False)
    ts.record_audit("p", passed=False)
    assert not ts.is_isolated("p") and ts.get_trust("p").consecutive_audit_failures == 2
    ts.record_audit("p", passed=False)
    assert ts.is_isolated("p") and [t.peer_id for t in ts.list_isolated()] == ["p"] and ts.list_peers() == []
    assert [t.peer_id for t in ts.list_peers(include_isolated=True)] == ["p"]
    ts.unisolate("p")
    assert not ts.is_isolated("p") and ts.get_trust("p").consecutive_audit_failures == 0


def test_summary_ratings_forward_to_reputation_tracker():
    rep = RP.LLMReputationTracker()
    ts = SC.TrustStore(reputation_tracker=rep)
    ts.record_summary_rating("p", 0.8)
    assert rep.get_reputation("p").total_ratings == 1


# ------------------------------------------------------------------ merkle
def test_merkle_build_proofs_for_every_leaf_including_odd_counts():
    for n in (1, 2, 3, 5, 8, 13):
        docs = [content_hash(f"doc{i}") for i in range(n)]
        t = MK.MerkleTree()
        root = t.build(docs)
        assert t.leaf_count == n and len(root) == 64
        for i, d in enumerate(docs):
            pr = t.get_proof(i)
            assert MK.MerkleTree.verify_proof(pr) and MK.MerkleTree.verify_document(d, pr) and pr.root_hash == root
            assert not MK.MerkleTree.verify_document(content_hash("other"), pr)


def test_merkle_errors_and_tamper_detection():
    t = MK.MerkleTree()
    with pytest.raises(RuntimeError):
        t.get_proof(0)
    with pytest.raises(ValueError):
        t.build([])
    t.build([content_hash(str(i)) for i in range(4)])
    with pytest.raises(IndexError):
        t.get_proof(4)
    pr = t.get_proof(1)
    bad = dataclasses.replace(pr, proof_path=((pr.proof_path[0][0][::-1], pr.proof_path[0][1]),) + pr.proof_path[1:])
    assert not MK.MerkleTree.verify_proof(bad)
    assert MK.deserialize_proof(MK.serialize_proof(pr)) == pr
    assert MK.MerkleTree().root_hash == "" and t.height == 3


def test_merkle_order_matters_and_root_signature():
    a, b = MK.MerkleTree(), MK.MerkleTree()
    docs = [content_hash(str(i)) for i in range(4)]
    assert a.build(docs) != b.build(docs[::-1])
    kp, other = KeyPair.generate(), KeyPair.generate()
    rec = a.create_root_record(kp.peer_id, kp)
    assert MK.verify_root_record(rec, kp.public_key_bytes()) and not MK.verify_root_record(rec, other.public_key_bytes())
    rt = MK.deserialize_merkle_root(MK.serialize_merkle_root(rec))
    assert rt == rec and AT.verify_merkle_root(rt, kp) and not AT.verify_merkle_root(a.create_root_record("x"), kp)


# ------------------------------------------------------------------ attestation
def test_attestation_create_verify_and_mismatch_reporting():
    kp = KeyPair.generate()
    att = AT.create_attestation("https://u", b"<html>raw</html>", "clean text", kp, crawled_at=123.0)
    assert att.raw_hash == content_hash(b"<html>raw</html>") and att.content_length == 10 and att.peer_id == kp.peer_id
    assert AT.verify_attestation(att, kp).verified
    assert AT.verify_attestation(att, kp, raw_body=b"<html>raw</html>", extracted_text="clean text").detail == "ok"
    r = AT.verify_attestation(att, kp, raw_body=b"changed", extracted_text="clean text")
    assert not r.verified and not r.raw_match and r.text_match and r.detail == "raw_hash_mismatch"
    forged = dataclasses.replace(att, text_hash=content_hash("evil"))
    r = AT.verify_attestation(forged, kp, extracted_text="evil")
    assert r.text_match and not r.signature_valid and "signature_invalid" in r.detail


def test_attestation_codec_and_raw_key_verification():
    kp, other = KeyPair.generate(), KeyPair.generate()
    att = AT.create_attestation("https://u", b"r", "t", kp)
    back = AT.deserialize_attestation(AT.serialize_attestation(att))
    assert back == att and AT.verify_attestation_with_key(back, kp.public_key_bytes())
    assert not AT.verify_attestation_with_key(back, other.public_key_bytes())


# ------------------------------------------------------------------ reputation
def test_reputation_unknown_until_min_samples_then_graded_by_ema():
    r = RP.LLMReputationTracker()
    assert r.get_reputation("p") is None and r.get_quality_score("p") == 0.5
    for _ in range(RP.MIN_SAMPLES - 1):
        r.record_quality("p", 1.0)
    assert r.get_reputation("p").grade == RP.ReputationGrade.UNKNOWN
    r.record_quality("p", 1.0)
    rep = r.get_reputation("p")
    assert rep.grade == RP.ReputationGrade.EXCELLENT and rep.avg_quality == 1.0 and rep.recent_ratings == RP.MIN_SAMPLES
    ema = 0.5
    for _ in range(RP.MIN_SAMPLES):
        ema = RP.EMA_ALPHA * 1.0 + (1 - RP.EMA_ALPHA) * ema
    assert rep.ema_quality == round(ema, 4)


def test_reputation_ranking_filters_and_recent_window():
    r = RP.LLMReputationTracker()
    now = 10_000_000.0
    for _ in range(6):
        r.record_quality("good", 0.95, now=now)
        r.record_quality("bad", -3, now=now - 30 * 24 * 3600)       # clamped to 0, outside the 7-day window
    r.record_quality("new", 0.9, now=now)
    assert r.best_peers() == ["good", "bad"] and r.best_peers(limit=1) == ["good"]
    assert [p.peer_id for p in r.list_peers(grade=RP.ReputationGrade.UNRELIABLE)] == ["bad"]
    assert r.get_reputation("bad", now=now).recent_ratings == 0 and r.get_reputation("good", now=now).recent_avg == 0.95
    assert [p.peer_id for p in r.list_peers(min_ratings=1)][-1] == "bad" and r.prune_log(max_age_seconds=0) >= 13
