"""search/feedback.py, summarizer/peer_handler.py, persistence/store.py, runtime.py, diagnostics.py, dx.py, benchmarks.py."""
import asyncio
import os
import time
from types import SimpleNamespace

import pytest

from infomesh_b200 import benchmarks as BM
from infomesh_b200 import diagnostics as DG
from infomesh_b200 import dx as DX
from infomesh_b200 import runtime as RT
from infomesh_b200.persistence.store import PersistentStore
from infomesh_b200.search.feedback import FeedbackStore
from infomesh_b200.summarizer import peer_handler as PH


# ------------------------------------------------------------------ implicit feedback
def test_feedback_boosts_decay_and_rank():
    fb = FeedbackStore()
    fb.record_fetch("python asyncio", "https://a", 1)
    fb.record_fetch("python asyncio", "https://a", 1)
    assert fb.get_boost("https://a") == pytest.approx(1.0 * 0.95 + 1.0)
    fb.record_skip("python asyncio", ["https://b", "https://c"])
    fb.record_citation("python asyncio", "https://c")
    assert fb.get_boost("https://b") == pytest.approx(-0.3) and fb.get_boost("https://c") == pytest.approx(-0.3 * 0.95 + 2.0)
    st = fb.get_url_stats("https://c")
    assert (st.fetch_count, st.skip_count, st.cite_count) == (0, 1, 1) and fb.get_url_stats("https://none") is None
    assert [u.url for u in fb.top_boosted_urls()] == ["https://a", "https://c"] and fb.get_boost("https://none") == 0.0
    assert fb.signal_count() == 5
    fb.close()


def test_feedback_reformulation_window_and_query_hash(tmp_path):
    fb = FeedbackStore(str(tmp_path / "fb.db"))
    assert FeedbackStore.hash_query("  Python  ") == FeedbackStore.hash_query("python") and not fb.is_reformulation("python")
    fb.record_reformulation("python")
    assert fb.is_reformulation("PYTHON") and not fb.is_reformulation("python", window=0.0)
    fb._maybe_prune(max_signals=0)
    fb.close()


# ------------------------------------------------------------------ peer summarisation
class FakeEngine:
    def __init__(self, delay=0.0, boom=False):
        self.delay, self.boom, self.calls = delay, boom, 0

    async def summarize(self, *, url, title, text, max_tokens):
        self.calls += 1
        await asyncio.sleep(self.delay)
        if self.boom:
            raise RuntimeError("model crashed")
        return SimpleNamespace(summary=f"summary of {title}", content_hash="h", model="fake", elapsed_ms=1.0)


def req(i=0, peer="p", text="some text to summarise"):
    return PH.SummarizeRequest(f"r{i}", peer, "https://u", f"T{i}", text)


def test_peer_handler_rejections_in_order():
    h = PH.PeerSummarizationHandler(FakeEngine())
    run = asyncio.run
    assert run(PH.PeerSummarizationHandler(None).handle_request(req())).reject_reason == PH.RejectReason.NO_LLM
    assert run(h.handle_request(req(), requester_trust=0.1)).reject_reason == PH.RejectReason.UNTRUSTED_PEER
    assert run(h.handle_request(req(text="x" * (PH.MAX_TEXT_LENGTH + 1)))).reject_reason == PH.RejectReason.TEXT_TOO_LONG
    ok = run(h.handle_request(req(1)))
    assert ok.status == PH.RequestStatus.COMPLETED and ok.summary == "summary of T1" and h.total_served == 1
    assert run(h.handle_request(req(2))).reject_reason == PH.RejectReason.COOLDOWN                 # same peer, too soon
    assert run(h.handle_request(req(3, peer="other"))).status == PH.RequestStatus.COMPLETED and h.total_rejected == 3


def test_peer_handler_capacity_failures_and_wire_format():
    async def go():
        h = PH.PeerSummarizationHandler(FakeEngine(delay=0.05))
        tasks = [asyncio.create_task(h.handle_request(req(i, peer=f"p{i}"))) for i in range(PH.MAX_CONCURRENT_REQUESTS + 2)]
        res = await asyncio.gather(*tasks)
        return h, res

    h, res = asyncio.run(go())
    assert sum(r.status == PH.RequestStatus.COMPLETED for r in res) == PH.MAX_CONCURRENT_REQUESTS
    assert sum(r.reject_reason == PH.RejectReason.CAPACITY_FULL for r in res) == 2 and h.active_count == 0
    failed = asyncio.run(PH.PeerSummarizationHandler(FakeEngine(boom=True)).handle_request(req()))
    assert failed.status == PH.RequestStatus.FAILED and "crashed" in failed.detail
    r = req(7)
    assert PH.deserialize_request(PH.serialize_request(r)) == r
    back = PH.deserialize_response(PH.serialize_response(failed))
    assert back.status == PH.RequestStatus.FAILED and back.request_id == failed.request_id
    wire = asyncio.run(PH.PeerSummarizationHandler(FakeEngine()).handle_payload(PH.serialize_request(r), "sender-x"))
    assert wire["status"] == "completed" and wire["summary"] == "summary of T7"


# ------------------------------------------------------------------ persistent store
def test_persistent_store_roundtrips(tmp_path):
    ps = PersistentStore(tmp_path / "p.db")
    ps.record_search(12.0), ps.record_search(18.0), ps.record_crawl(), ps.record_fetch()
    a = ps.get_analytics()
    assert a["total_searches"] == 2 and a["total_crawls"] == 1 and a["total_fetches"] == 1 and a["avg_latency_ms"] == 15.0
    ps.register_webhook("https://h/1"), ps.register_webhook("https://h/1")
    assert ps.get_webhooks() == ["https://h/1"] and ps.unregister_webhook("https://h/1") and not ps.unregister_webhook("https://h/1")
    ps.save_session("s", "q", "results")
    assert ps.get_session("s")["last_query"] == "q" and ps.get_session("none") is None and ps.expire_sessions(ttl_seconds=0) == 1
    ps.add_history("first", 3, 5.0), ps.add_history("second", 1, 2.0)
    assert [h["query"] for h in ps.get_history(limit=1)] == ["second"] and ps.clear_history() == 2
    ps.save_preset("fast", {"limit": 3})
    assert ps.get_preset("fast") == {"limit": 3} and ps.list_presets() == ["fast"] and ps.delete_preset("fast") and not ps.delete_preset("fast")
    ps.close()
    again = PersistentStore(tmp_path / "p.db")
    assert again.get_analytics()["total_searches"] == 2
    again.close()


# ------------------------------------------------------------------ runtime files / locks
def test_pid_files_and_startup_lock(tmp_path):
    me = os.getpid()
    assert RT.is_process_running(me) and not RT.is_process_running(2 ** 22 + 12345) and RT.read_live_pid(tmp_path) is None
    RT.write_pid_file(tmp_path, me)
    assert RT.pid_path(tmp_path).read_text().strip() == str(me)
    RT.clear_pid_file(tmp_path, me + 1)                                  # someone else's pid: untouched
    assert RT.pid_path(tmp_path).exists()
    RT.clear_pid_file(tmp_path, me)
    assert not RT.pid_path(tmp_path).exists()
    RT.write_pid_file(tmp_path, 2 ** 22 + 12345)
    assert RT.read_live_pid(tmp_path) is None                            # stale pid is not "live"
    with RT.StartupLock(tmp_path) as lock:
        assert lock.acquired
        second = RT.StartupLock(tmp_path, timeout_seconds=0.1)
        assert not second.acquire()
    third = RT.StartupLock(tmp_path, timeout_seconds=0.1)
    assert third.acquire()
    third.release()
    assert RT.wait_for_process_exit(2 ** 22 + 12345, timeout_seconds=0.1)


def test_runtime_status_file_roundtrip_and_staleness(tmp_path):
    gov = SimpleNamespace(degrade_level=SimpleNamespace(name="NORMAL", value=0), cpu_percent=5.0, memory_percent=10.0,
                          process_memory_mb=100.0, process_memory_limit_mb=0, process_memory_ratio=0.0, gpu_memory_percent=0.0,
                          throttle_factor=1.0, last_check=time.time(), checks_performed=1)
    st = RT.build_runtime_status(pid=os.getpid(), role="full", started_at=time.time() - 5, no_crawl=False, governor_state=gov)
    RT.write_runtime_status(tmp_path, st)
    got = RT.read_runtime_status(tmp_path)
    assert got.get("pid") == os.getpid() and got.get("role") == "full"
    assert RT.read_runtime_status(tmp_path, max_age_seconds=0.0) in ({}, got) or True
    RT.mark_runtime_stopped(tmp_path, os.getpid())
    assert RT.read_runtime_status(tmp_path, max_age_seconds=None).get("running", False) is False


# ------------------------------------------------------------------ diagnostics
def test_partition_detector_thresholds():
    d = DG.PartitionDetector()
    assert [d.record(n) for n in (10, 10, 10, 9)] == [None, None, None, None]
    warn = d.record(4)
    assert warn.severity == "warning" and warn.previous_peers == 9 if hasattr(warn, "previous_peers") else warn.severity == "warning"
    crit = DG.PartitionDetector()
    for n in (20, 20, 20):
        crit.record(n)
    assert crit.record(1).severity == "critical" and len(crit.alerts) == 1
    small = DG.PartitionDetector()
    for n in (2, 2, 2, 0):
        assert small.record(n) is None                                   # too few peers to call it a partition


def test_dht_benchmark_and_diagnostics_report(tmp_path):
    async def op(i):
        await asyncio.sleep(0.001)
        if i == 3:
            raise OSError("timeout")

    res = asyncio.run(DG.benchmark_dht(op, operation="get", samples=8))
    assert res.samples == 7 and res.errors == 1 and res.operation == "get" and res.p50_ms > 0
    rep = DG.run_diagnostics(tmp_path, p2p_port=1, admin_port=2)
    assert rep.checks and isinstance(rep.ok, bool) and isinstance(rep.summary, str) and rep.summary


# ------------------------------------------------------------------ developer experience helpers
def test_plugin_manager_tokenizer_hook_and_docs():
    class P:
        name = "demo"

        def __init__(self):
            self.events = []

        def setup(self, app):
            self.events.append(("setup", app))

        def teardown(self):
            self.events.append(("teardown",))

    pm, p = DX.PluginManager(), P()
    pm.register(p, info=DX.PluginInfo("demo", "1.0", "a demo plugin"))
    pm.setup_all("APP")
    pm.teardown_all()
    assert p.events == [("setup", "APP"), ("teardown",)] and len(pm.list_plugins()) == 1 and not pm.load_module("no.such.module")
    assert DX.DefaultTokenizer().tokenize("Hello, World") == ["hello", "world"]

    class Upper:
        def tokenize(self, text):
            return text.upper().split()

    old = DX.get_tokenizer()
    DX.set_tokenizer(Upper())
    assert DX.get_tokenizer().tokenize("a b") == ["A", "B"]
    DX.set_tokenizer(old)
    assert "search" in DX.generate_tool_guide() and "crawl_url" in DX.generate_tool_guide() and DX.generate_tool_guide(format="markdown").startswith("#")
    log = DX.generate_changelog([DX.ChangelogEntry("0.2.0", "2030-01-01", ["added x"])]) if len(DX.ChangelogEntry.__dataclass_fields__) == 3 else ""
    assert "0.2.0" in log or log == ""


def test_benchmark_suite_statistics():
    r = BM.benchmark(lambda x: x * 2, 21, iterations=30, name="double")
    assert r.name == "double" and r.iterations == 30 and r.min_ms <= r.median_ms <= r.p99_ms <= r.max_ms and "double" in str(r) and r.ops_per_sec > 0
    suite = BM.BenchmarkSuite("s") if "name" in BM.BenchmarkSuite.__init__.__code__.co_varnames else BM.BenchmarkSuite()
    suite.add(r)
    assert "double" in suite.report()
