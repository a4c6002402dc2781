"""Platform layer: services/AppContext, resources, summariser (engine / verify / peer handler), runtime files,
errors, plugins, SLOs, security ops/ext, metrics, persistence (model: reference tests/test_services.py,
test_resources.py, test_summarizer*.py, test_runtime.py, test_security_ext.py, test_observability.py ...)."""
import asyncio
import json
import os
import time
from dataclasses import replace

import pytest

from infomesh_b200.config import Config


def _cfg(tmp_path, **node):
    base = Config()
    return replace(base, node=replace(base.node, data_dir=tmp_path, **node),
                   index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False))


def test_app_context_roles_and_index_flow(tmp_path):
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.services import AppContext, fetch_page, index_document, is_paywall_content, republish_local_index

    with AppContext(_cfg(tmp_path)) as ctx:
        assert ctx.worker is not None and ctx.link_graph is not None and ctx.ledger is not None and ctx.key_pair is not None
        page = ParsedPage(url="https://example.com/a", title="Alpha", text="tensor memory accumulators " * 10, language="en",
                          raw_html_hash="h1", text_hash="t1")
        doc_id = index_document(page, ctx.store)
        assert doc_id == 1 and index_document(page, ctx.store) is None          # duplicate
        got = fetch_page("https://example.com/a", store=ctx.store)
        assert got.success and got.is_cached and not got.is_stale and got.title == "Alpha"
        assert fetch_page("http://127.0.0.1/x", store=ctx.store).error.startswith("blocked")
        assert fetch_page("https://example.com/missing", store=ctx.store).error == "not_cached"

        class FakeDist:
            def __init__(self):
                self.batches = []

            async def publish_batch(self, docs):
                self.batches.append(docs)
                return 3 * len(docs)

        fd = FakeDist()
        assert asyncio.run(republish_local_index(ctx.store, distributed_index=fd)) == 3 and len(fd.batches) == 1
    assert is_paywall_content("Please Subscribe to continue reading") and not is_paywall_content("free text")
    search_only = AppContext(_cfg(tmp_path / "s", role="search"))
    assert search_only.worker is None and search_only.index_submit_receiver is not None
    search_only.close()
    crawler = AppContext(_cfg(tmp_path / "c", role="crawler"))
    assert crawler.worker is not None and crawler.link_graph is None
    crawler.close()


def test_profiles_governor_preflight(tmp_path):
    from infomesh_b200.resources import governor as G
    from infomesh_b200.resources import preflight as P
    from infomesh_b200.resources.profiles import ProfileName, build_custom_profile, get_profile

    assert get_profile("minimal").max_concurrent_crawl == 1 and get_profile("dedicated").cpu_cores_limit == 0
    cust = build_custom_profile(max_concurrent_crawl=7, bogus=1)
    assert cust.name == ProfileName.CUSTOM and cust.max_concurrent_crawl == 7 and cust.cpu_nice == 10
    with pytest.raises(ValueError):
        get_profile("turbo")
    assert G.classify(50, 50, 0.1) == G.DegradeLevel.NORMAL and G.classify(65, 10, 0) == G.DegradeLevel.WARNING
    assert G.classify(10, 86, 0) == G.DegradeLevel.OVERLOADED and G.classify(10, 10, 1.0) == G.DegradeLevel.SEVERE
    assert G.classify(96, 0, 0) == G.DegradeLevel.DEFENSIVE
    assert G.throttle_for(G.DegradeLevel.NORMAL, 20) == 1.0 and G.throttle_for(G.DegradeLevel.NORMAL, 55) == pytest.approx(0.65)
    assert G.throttle_for(G.DegradeLevel.SEVERE, 0) == 0.0
    gov = G.ResourceGovernor(get_profile("balanced"))
    st = gov.check_and_adjust()
    assert st.checks_performed == 1 and 0.0 <= st.throttle_factor <= 1.0 and gov.effective_max_concurrent >= 1
    assert P.check_disk_space(tmp_path) == [] or P.check_disk_space(tmp_path)[0].check == "disk_space"
    bad = P.check_outbound_connectivity([("127.0.0.1", 1)])
    assert bad and bad[0].severity == P.IssueSeverity.ERROR
    assert not P.is_disk_critically_low(tmp_path)


def test_port_check_helpers():
    from infomesh_b200.resources import port_check as PC

    with pytest.raises(ValueError):
        PC._validate_port(70000)
    with pytest.raises(ValueError):
        PC._validate_port(True)
    nsg = PC.NsgInfo.from_resource_id("/subscriptions/s/resourceGroups/rg1/providers/Microsoft.Network/networkSecurityGroups/nsgA", "NIC n")
    assert (nsg.name, nsg.resource_group) == ("nsgA", "rg1")
    assert "4001" in PC._get_manual_instructions(PC.CloudProvider.AWS, 4001) and "portproxy" in PC._get_wsl_manual_instructions(4001)
    assert PC.is_port_listening(1) is False


def test_summarizer_engine_verify_and_peer_handler():
    from infomesh_b200.summarizer import verify as V
    from infomesh_b200.summarizer.engine import LLMBackend, LLMRuntime, ModelInfo, SummarizationEngine, create_backend
    from infomesh_b200.summarizer.peer_handler import (PeerSummarizationHandler, RejectReason, RequestStatus, SummarizeRequest,
                                                       deserialize_response, serialize_response)

    class Echo(LLMBackend):
        async def generate(self, prompt, *, max_tokens=512):
            assert "Summary:" in prompt
            return "  Blackwell B200 has 148 SMs and 180 GB of HBM3e memory.  "

        async def is_available(self):
            return True

        async def model_info(self):
            return ModelInfo("echo", LLMRuntime.OLLAMA, "1B", None, True)

    src = ("The Blackwell B200 GPU has 148 SMs. It carries 180 GB of HBM3e memory on two dies. "
           "NVLink 5 gives every GPU 900 GB per second in each direction. The tensor cores accumulate in TMEM.")
    eng = SummarizationEngine(Echo())
    res = asyncio.run(eng.summarize("https://x.example", "B200", src))
    assert res.summary.startswith("Blackwell") and res.model == "echo" and res.token_count > 0 and len(res.content_hash) == 64
    rep = V.verify_summary("https://x.example", res.content_hash, src, res.summary, peer_summaries=["B200 has 148 SMs and 180 GB HBM3e memory"])
    assert rep.self_check.passed and rep.level == V.VerificationLevel.CROSS_VALIDATED and 0.5 < rep.quality_score <= 1.0
    bad = V.self_verify(src, "It costs 17 dollars, weighs 93 kg and ships in 2031.")
    assert bad.has_contradiction and not bad.passed
    assert V.compute_similarity("a b c", "b c d") == pytest.approx(0.5) and not V.cross_validate("x", []).passed
    for name, cls in (("ollama", "OllamaBackend"), ("llama_cpp", "LlamaCppBackend"), ("vllm", "VLLMBackend"), ("b200", "B200Backend")):
        assert type(create_backend(name, "m")).__name__ == cls
    with pytest.raises(ValueError):
        create_backend("gpt")
    assert asyncio.run(create_backend("ollama", base_url="http://127.0.0.1:9").is_available()) is False

    h = PeerSummarizationHandler(eng)
    req = SummarizeRequest("r1", "peerA", "https://x.example", "B200", src)

    async def flow():
        ok = await h.handle_request(req, requester_trust=0.9)
        cool = await h.handle_request(req, requester_trust=0.9)
        low = await h.handle_request(SummarizeRequest("r2", "peerB", "u", "t", src), requester_trust=0.1)
        big = await h.handle_request(SummarizeRequest("r3", "peerC", "u", "t", "x" * 20000), requester_trust=0.9)
        wire = await h.handle_payload({"request_id": "r4", "url": "u", "title": "t", "text": src}, "peerD")
        return ok, cool, low, big, wire

    ok, cool, low, big, wire = asyncio.run(flow())
    assert ok.status == RequestStatus.COMPLETED and ok.summary.startswith("Blackwell")
    assert cool.reject_reason == RejectReason.COOLDOWN and low.reject_reason == RejectReason.UNTRUSTED_PEER
    assert big.reject_reason == RejectReason.TEXT_TOO_LONG and wire["status"] == "completed"
    assert deserialize_response(serialize_response(low)) == low and h.total_served == 2 and h.total_rejected == 3


def test_runtime_pid_lock_and_status(tmp_path):
    from infomesh_b200 import runtime as R
    from infomesh_b200.resources.governor import GovernorState

    assert R.read_live_pid(tmp_path) is None
    R.write_pid_file(tmp_path, os.getpid())
    assert R.read_live_pid(tmp_path) == os.getpid()
    R.clear_pid_file(tmp_path, 1)                                   # not the owner: kept
    assert R.pid_path(tmp_path).exists()
    R.clear_pid_file(tmp_path, os.getpid())
    R.pid_path(tmp_path).write_text("999999999")
    assert R.read_live_pid(tmp_path) is None and not R.pid_path(tmp_path).exists()
    with R.StartupLock(tmp_path):
        second = R.StartupLock(tmp_path, timeout_seconds=0.1)
        assert second.acquire() is False
    third = R.StartupLock(tmp_path, timeout_seconds=0.1)
    assert third.acquire() is True
    third.release()
    st = R.build_runtime_status(pid=1, role="full", started_at=time.time() - 5, no_crawl=False, governor_state=GovernorState())
    R.write_runtime_status(tmp_path, st)
    assert R.read_runtime_status(tmp_path)["degrade_level"] == "NORMAL"
    st["updated_at"] -= 100
    R.write_runtime_status(tmp_path, st)
    assert R.read_runtime_status(tmp_path)["stale"] is True
    R.mark_runtime_stopped(tmp_path, 1)
    assert R.read_runtime_status(tmp_path, max_age_seconds=None)["status"] == "stopped"


def test_errors_plugins_slo_shutdown_dx():
    from infomesh_b200 import dx
    from infomesh_b200.errors import ERRORS, format_error, get_error
    from infomesh_b200.plugins import HookPoint, PluginRegistry, get_registry
    from infomesh_b200.shutdown import GracefulShutdown
    from infomesh_b200.slo import SLOTracker

    assert get_error("E001").http_status == 401 and ERRORS["E006"].code == "INFOMESH_E006" and "Unknown" in format_error("E999")
    assert get_error("E004").to_dict()["error"]["category"] == "SEARCH" and len([k for k in ERRORS if k < "E100"]) == 20
    reg = PluginRegistry()

    @reg.hook(HookPoint.PRE_INDEX)
    def drop_spam(doc):
        return None if "spam" in doc["text"] else doc

    reg.register_plugin("upper", "1.0", {HookPoint.PRE_INDEX: lambda d: {**d, "text": d["text"].upper()}})
    reg.register_plugin("boom", hooks={HookPoint.PRE_INDEX: lambda d: 1 / 0})
    assert reg.run_hook(HookPoint.PRE_INDEX, {"text": "ok"}) == {"text": "OK"} and reg.run_hook(HookPoint.PRE_INDEX, {"text": "spam"}) is None
    assert reg.hook_counts == {"pre_index": 3} and reg.unregister_plugin("boom") and get_registry() is get_registry()

    async def ahook(d):
        return d + 1

    reg.register_plugin("async", hooks={HookPoint.POST_RANK: ahook})
    assert asyncio.run(reg.run_hook_async(HookPoint.POST_RANK, 1)) == 2
    slo = SLOTracker()
    for ms in (10, 20, 3000):
        slo.record("search_latency_p99", ms)
    for ok in (True,) * 9 + (False,):
        slo.record_success("crawl_success_rate", ok)
    by = {d["name"]: d for d in slo.summary()["details"]}
    assert by["search_latency_p99"]["met"] is False and by["crawl_success_rate"]["met"] is True and by["node_uptime"]["met"] is True
    closed = []

    class Ctx:
        async def close_async(self):
            closed.append("ctx")

    sd = GracefulShutdown()
    sd._context = Ctx()
    sd.add_callback(lambda: closed.append("cb"))
    assert sd._try_set_shutting_down() and not sd._try_set_shutting_down()
    asyncio.run(sd.cleanup())
    assert closed == ["cb", "ctx"] and sd.is_shutting_down
    assert dx.get_tokenizer().tokenize("A big GPU!") == ["big", "gpu"]
    assert "## `search`" in dx.generate_tool_guide(format="markdown") and "crawl_url" in dx.generate_tool_guide()
    assert "### Breaking Changes" in dx.generate_changelog([dx.ChangelogEntry("1.0", "2026-01-01", ["a"], ["b"])])


def test_security_ops_and_ext(tmp_path):
    from infomesh_b200 import security_ext as X
    from infomesh_b200.security_ops import APIKeyManager, AuditLogger

    km = APIKeyManager(tmp_path / "keys.json")
    km.add_key("secret-1", "main")
    assert km.validate("secret-1") and not km.validate("nope")
    km.rotate("main", "secret-2", grace_days=1)
    assert km.validate("secret-1") and km.validate("secret-2")
    km.revoke("main")
    assert not APIKeyManager(tmp_path / "keys.json").validate("secret-1")
    assert "secret" not in (tmp_path / "keys.json").read_text()
    al = AuditLogger(tmp_path / "audit.log")
    al.log("search", client="1.2.3.4", details="q")
    assert al.recent()[0].action == "search"
    tok = X.make_jwt_token({"sub": "u", "exp": time.time() + 60}, "k")
    assert X.verify_jwt_token(tok, "k")["sub"] == "u" and X.verify_jwt_token(tok, "other") is None
    assert X.verify_jwt_token(X.make_jwt_token({"exp": time.time() - 1}, "k"), "k") is None
    import base64
    none_tok = base64.urlsafe_b64encode(b'{"alg":"none"}').decode().rstrip("=") + "." + tok.split(".")[1] + "."
    assert X.verify_jwt_token(none_tok, "k") is None
    assert X.check_role("crawl_url", "crawler") and not X.check_role("crawl_url", "reader") and not X.check_role("search", "root")
    f = X.IPFilter(allowlist={"10.0.0.0/8"}, blocklist={"10.1.1.1"})
    assert f.is_allowed("10.2.3.4") and not f.is_allowed("10.1.1.1") and not f.is_allowed("8.8.8.8")
    sig = X.sign_webhook_payload({"b": 1, "a": 2}, "s")
    assert sig.startswith("sha256=") and X.verify_webhook_signature({"a": 2, "b": 1}, sig, "s")
    assert X.TLSConfig(enabled=True).validate() == ["TLS cert_file is required", "TLS key_file is required"]
    with X.AuditLog(tmp_path / "a.db") as log:
        log.log("search", api_key="abc", arguments={"query": "x", "api_key": "abc"}, latency_ms=3)
        row = log.query(tool_name="search")[0]
        assert "abc" not in row["arguments_json"] and len(row["api_key_hash"]) == 16


def test_metrics_persistence_scalability_misc(tmp_path):
    from infomesh_b200 import scalability as S
    from infomesh_b200.benchmarks import BenchmarkSuite, benchmark
    from infomesh_b200.diagnostics import PartitionDetector, run_diagnostics
    from infomesh_b200.observability import metrics as M
    from infomesh_b200.persistence.store import PersistentStore
    from infomesh_b200.search.feedback import FeedbackStore
    from infomesh_b200.version_check import PeerVersionTracker, is_newer

    mc = M.MetricsCollector()
    mc.inc("search.total")
    mc.set_gauge("p2p_peers", 3)
    for v in (1.0, 2.0, 3.0):
        mc.observe("lat ms", v)
    text = mc.format_prometheus()
    assert "search_total 1.0" in text and "lat_ms_count 3" in text and 'lat_ms{quantile="0.5"} 2.000' in text
    assert mc.to_dict()["histograms"]["lat ms"]["avg"] == 2.0
    tr = M.QueryTrace("t", "q")
    tr.add_span(M.QuerySpan("s1", "p", "search", latency_ms=4.0))
    assert tr.to_dict()["total_latency_ms"] == 4.0 and len(M.generate_alert_rules()) >= 5
    assert len(M.generate_grafana_dashboard()["dashboard"]["panels"]) >= 6
    with PersistentStore(tmp_path / "p.db") as ps:
        ps.record_search(10)
        ps.record_search(30)
        ps.record_crawl()
        assert ps.get_analytics() == {"total_searches": 2, "total_crawls": 1, "total_fetches": 0, "avg_latency_ms": 20.0}
        ps.register_webhook("https://h.example/w")
        assert ps.get_webhooks() == ["https://h.example/w"] and ps.unregister_webhook("https://h.example/w")
        ps.save_session("s", "q", "r" * 5000)
        assert len(ps.get_session("s")["last_results"]) == 2000 and ps.expire_sessions(-1) == 1
        ps.add_history("q1", 3, 1.0)
        ps.add_history("q2", 1, 2.0)
        assert [h["query"] for h in ps.get_history()] == ["q2", "q1"] and ps.clear_history() == 2
        ps.save_preset("docs", {"language": "en"})
        assert ps.get_preset("docs") == {"language": "en"} and ps.list_presets() == ["docs"] and ps.delete_preset("docs")
    bf = S.BloomFilter(1000, 0.01)
    for i in range(500):
        bf.add(f"https://e.example/{i}")
    assert all(f"https://e.example/{i}" in bf for i in range(500))
    assert sum(f"https://other.example/{i}" in bf for i in range(2000)) < 60
    pool = S.ConnectionPool(str(tmp_path / "pool.db"), 2)
    with pool.connection() as c:
        c.execute("CREATE TABLE t (x)")
    pool.close_all()
    r = benchmark(lambda: sum(range(100)), iterations=20, name="sum")
    suite = BenchmarkSuite()
    suite.add(r)
    assert r.iterations == 20 and "sum:" in suite.report()
    pd = PartitionDetector()
    for n in (10, 10, 10, 10):
        assert pd.record(n) is None
    assert pd.record(1).severity == "critical"
    rep = run_diagnostics(tmp_path, p2p_port=1, admin_port=1)
    assert {c.name for c in rep.checks} >= {"data_dir", "key_pair", "index_db", "disk_space", "gpu"} and "ok" in rep.summary
    fb = FeedbackStore()
    fb.record_fetch("q", "https://a", 2)
    fb.record_citation("q", "https://a")
    fb.record_skip("q", ["https://b"])
    assert fb.get_boost("https://a") == pytest.approx(0.95 + 2.0) and fb.get_boost("https://b") == pytest.approx(-0.3)
    assert fb.is_reformulation("Q ") and fb.top_boosted_urls()[0].url == "https://a" and fb.signal_count() == 3
    assert is_newer("0.2.0", "0.1.9") and not is_newer("0.1.0rc1", "0.1.0")
    pv = PeerVersionTracker()
    pv.record("p", "9.9.9")
    assert pv.check_peer_update().latest == "9.9.9"
