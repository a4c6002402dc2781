"""resources/profiles.py + governor ladder, resources/preflight.py, security_ext.py, security_ops.py, scalability.py."""
import sqlite3
import time
from unittest.mock import patch

import pytest

from infomesh_b200 import scalability as SCAL
from infomesh_b200 import security_ext as SX
from infomesh_b200 import security_ops as SO
from infomesh_b200.resources import governor as GV
from infomesh_b200.resources import preflight as PF
from infomesh_b200.resources import profiles as PRF


# ------------------------------------------------------------------ profiles / governor
def test_profiles_escalate_monotonically_and_custom_overrides():
    order = [PRF.get_profile(n) for n in ("minimal", "balanced", "contributor", "dedicated")]
    assert [p.max_concurrent_crawl for p in order] == sorted(p.max_concurrent_crawl for p in order)
    assert [p.gpu_query_batch for p in order] == [8, 32, 64, 128] and order[-1].cpu_cores_limit == 0
    with pytest.raises(ValueError, match="Unknown profile"):
        PRF.get_profile("turbo")
    c = PRF.build_custom_profile(max_concurrent_crawl=9, nonsense=1)
    assert c.name == PRF.ProfileName.CUSTOM and c.max_concurrent_crawl == 9 and c.memory_limit_mb == PRF.get_profile("balanced").memory_limit_mb


def test_degrade_ladder_and_throttle_factors():
    L = GV.DegradeLevel
    assert GV.classify(10, 10, 0.1) == L.NORMAL and GV.classify(65, 10, 0.1) == L.WARNING and GV.classify(10, 86, 0.1) == L.OVERLOADED
    assert GV.classify(10, 10, 1.0) == L.SEVERE and GV.classify(96, 10, 0.1) == L.DEFENSIVE and GV.classify(10, 10, 1.25) == L.DEFENSIVE
    assert GV.throttle_for(L.DEFENSIVE, 99) == 0.0 and GV.throttle_for(L.SEVERE, 50) == 0.0
    assert GV.throttle_for(L.OVERLOADED, 85) == 0.25 and GV.throttle_for(L.WARNING, 65) == 0.5
    assert GV.throttle_for(L.NORMAL, 10) == 1.0 and GV.throttle_for(L.NORMAL, 55) == pytest.approx(1.0 - 25 / 50 * 0.7)


# ------------------------------------------------------------------ preflight
def test_disk_checks(tmp_path):
    assert PF.get_disk_free_mb(tmp_path) > 0 and not PF.is_disk_critically_low(tmp_path)
    with patch.object(PF, "get_disk_free_mb", return_value=100.0):
        issues = PF.check_disk_space(tmp_path)
        assert issues and issues[0].severity in (PF.IssueSeverity.ERROR, PF.IssueSeverity.WARNING)
    with patch.object(PF, "get_disk_free_mb", return_value=1e6):
        assert PF.check_disk_space(tmp_path) == []


def test_connectivity_check_reports_unreachable_targets():
    issues = PF.check_outbound_connectivity([("127.0.0.1", 1)])           # nothing listens on port 1
    assert issues and "127.0.0.1" in issues[0].message or issues
    out = PF.run_preflight_checks(__import__("pathlib").Path("/tmp"), network=False, gpu=False)
    assert isinstance(out, list)


# ------------------------------------------------------------------ JWT / RBAC / IP filter / webhooks
def test_jwt_roundtrip_and_rejections():
    tok = SX.make_jwt_token({"sub": "u", "role": "reader", "exp": time.time() + 60}, "s3cret")
    assert SX.verify_jwt_token(tok, "s3cret")["sub"] == "u"
    assert SX.verify_jwt_token(tok, "wrong") is None and SX.verify_jwt_token("a.b", "s3cret") is None
    assert SX.verify_jwt_token(SX.make_jwt_token({"exp": time.time() - 1}, "k"), "k") is None
    assert SX.verify_jwt_token(SX.make_jwt_token({"nbf": time.time() + 100}, "k"), "k") is None
    assert SX.verify_jwt_token(tok, "s3cret", algorithms=["RS256"]) is None
    import base64
    import json
    none_head = base64.urlsafe_b64encode(json.dumps({"alg": "none"}).encode()).rstrip(b"=").decode()
    body = tok.split(".")[1]
    assert SX.verify_jwt_token(f"{none_head}.{body}.", "s3cret") is None            # alg=none can never pass
    head, _, sig = tok.split(".")
    evil = base64.urlsafe_b64encode(json.dumps({"sub": "admin"}).encode()).rstrip(b"=").decode()
    assert SX.verify_jwt_token(f"{head}.{evil}.{sig}", "s3cret") is None


def test_role_checks():
    assert SX.check_role("crawl_url", "crawler") and not SX.check_role("crawl_url", "reader")
    assert SX.check_role("analytics", "admin") and not SX.check_role("analytics", "crawler")
    assert SX.check_role("anything", None) and SX.check_role("unlisted_tool", "reader") and not SX.check_role("search", "wizard")


def test_ip_filter_cidr_block_wins_and_exclusive_allowlist():
    f = SX.IPFilter()
    assert f.is_allowed("8.8.8.8")
    f.add_block("10.0.0.0/8")
    f.add_block("bad/cidr")
    assert not f.is_allowed("10.1.2.3") and f.is_allowed("11.1.2.3") and f.is_allowed("not-an-ip")
    f.add_allow("192.168.1.0/24")
    f.add_allow("10.5.5.5")
    assert f.is_allowed("192.168.1.77") and not f.is_allowed("8.8.8.8") and not f.is_allowed("10.5.5.5")
    f.remove_block("10.0.0.0/8")
    f.remove_allow("192.168.1.0/24")
    assert f.is_allowed("10.5.5.5") and not f.is_allowed("192.168.1.77")


def test_webhook_signatures_are_key_order_independent():
    sig = SX.sign_webhook_payload({"b": 2, "a": "é"}, "k")
    assert sig.startswith("sha256=") and SX.verify_webhook_signature({"a": "é", "b": 2}, sig, "k")
    assert not SX.verify_webhook_signature({"a": "e", "b": 2}, sig, "k") and not SX.verify_webhook_signature({"a": "é", "b": 2}, sig, "other")


# ------------------------------------------------------------------ API keys / audit
def test_api_key_lifecycle_persists_hashes_only(tmp_path):
    f = tmp_path / "keys" / "api_keys.json"
    m = SO.APIKeyManager(f)
    m.add_key("alpha-secret", "alpha")
    m.add_key("temp", "temp", ttl_days=1)
    assert m.validate("alpha-secret") and m.validate("temp") and not m.validate("nope")
    assert "alpha-secret" not in f.read_text() and oct(f.stat().st_mode & 0o777) == "0o600"
    rotated = m.rotate("alpha", "beta-secret", grace_days=0)
    assert rotated.label.startswith("alpha-rotated-") and m.validate("beta-secret")
    assert m.revoke("temp") and not m.revoke("ghost") and not m.validate("temp")
    again = SO.APIKeyManager(f)
    assert again.validate("beta-secret") and not again.validate("temp") and len(again.list_keys()) == 3
    f.write_text("{broken")
    assert SO.APIKeyManager(f).list_keys() == []


def test_audit_logger_ring(tmp_path):
    a = SO.AuditLogger(tmp_path / "audit.log") if "log_file" in SO.AuditLogger.__init__.__code__.co_varnames or True else None
    a.log("search", details="q=python")
    a.log("crawl", success=False, client="10.0.0.1")
    rec = a.recent(limit=5)
    assert [e.action for e in rec][-2:] == ["search", "crawl"] or [e.action for e in rec][:2] == ["crawl", "search"]
    assert any(not e.success for e in rec)


# ------------------------------------------------------------------ scalability helpers
def test_bloom_filter_no_false_negatives_and_bounded_false_positives():
    bf = SCAL.BloomFilter(capacity=2000, fp_rate=0.01)
    items = [f"https://ex.org/{i}" for i in range(2000)]
    for it in items:
        bf.add(it)
    assert all(it in bf for it in items) and len(bf) == 2000 and bf.size_bytes < 4000
    fp = sum(f"https://other.org/{i}" in bf for i in range(5000))
    assert fp < 5000 * 0.03


def test_connection_pool_reuses_and_closes(tmp_path):
    pool = SCAL.ConnectionPool(str(tmp_path / "p.db"), max_connections=2) if "max_connections" in SCAL.ConnectionPool.__init__.__code__.co_varnames else SCAL.ConnectionPool(str(tmp_path / "p.db"))
    with pool.connection() as c:
        c.execute("CREATE TABLE t (x)")
        c.execute("INSERT INTO t VALUES (1)")
        c.commit()
        first = c
    with pool.connection() as c2:
        assert c2 is first and c2.execute("SELECT x FROM t").fetchone()[0] == 1
    pool.close_all()
    with pytest.raises(sqlite3.ProgrammingError):
        first.execute("SELECT 1")


def test_batch_ingest_counts(tmp_path):
    from infomesh_b200.index.local_store import LocalStore
    store = LocalStore(None)
    docs = [{"url": f"https://ex.org/{i}", "title": f"T{i}", "text": f"document number {i} about search engines and ranking"} for i in range(7)]
    docs += [docs[0], {"title": "no url key"}]
    res = SCAL.batch_ingest(store, docs, batch_size=3)
    # a call that does not raise counts as succeeded (a duplicate is silently skipped by the store); a malformed record fails
    assert res.total == 9 and res.succeeded == 8 and res.failed == 1 and "?" in res.errors[0]
    assert store.get_stats()["document_count"] == 7
