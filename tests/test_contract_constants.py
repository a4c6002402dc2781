"""SURVEY Appendix A (public surface) and Appendix B (behavioural constants) as executable checks: if one of these
drifts, a user switching over from the reference would notice."""
import pytest


# ------------------------------------------------------------------ Appendix A: CLI
def test_cli_surface():
    from infomesh_b200.cli import cli

    def opts(cmd):
        return {o for p in cmd.params for o in getattr(p, "opts", [])} | {p.name for p in cmd.params if p.param_type_name == "argument"}

    c = cli.commands
    assert {"start", "stop", "update", "_serve", "status", "crawl", "mcp", "dashboard", "search", "index", "config", "keys", "peer", "feeds",
            "feedback", "doctor", "bench"} <= set(c)
    assert c["_serve"].hidden and {"--seeds", "--role", "--no-crawl"} <= opts(c["_serve"])
    assert {"--seeds", "-s", "--background", "-b", "--no-dashboard", "--role", "-r"} <= opts(c["start"])
    assert "--check" in opts(c["update"]) and {"url", "--depth", "--force"} <= opts(c["crawl"])
    assert {"--http", "--host", "--port"} <= opts(c["mcp"]) and {"--tab", "--text"} <= opts(c["dashboard"])
    assert {"query", "-n", "--local", "--vector"} <= opts(c["search"])
    assert set(c["index"].commands) >= {"stats", "export", "import", "import-wet", "import-urls"} and "--starter" in opts(c["index"].commands["import"])
    assert "--max" in opts(c["index"].commands["import-urls"])
    assert set(c["config"].commands) == {"show", "set", "github"} and set(c["keys"].commands) == {"export", "rotate"}
    assert set(c["peer"].commands) == {"list", "add", "remove", "test"} and set(c["feeds"].commands) == {"import", "list"}
    assert set(c["feedback"].commands) == {"stats", "top-urls"} and "-n" in opts(c["feedback"].commands["top-urls"]) and "-n" in opts(c["bench"])
    limit = next(p for p in c["search"].params if "-n" in getattr(p, "opts", []))
    assert getattr(limit.type, "min", 1) == 1 and getattr(limit.type, "max", 100) == 100


# ------------------------------------------------------------------ Appendix A: MCP
def test_mcp_tool_surface(tmp_path, monkeypatch):
    from infomesh_b200.mcp import handlers as H
    from infomesh_b200.mcp.tools import get_all_tools

    monkeypatch.delenv("INFOMESH_API_KEY", raising=False)
    tools = {t.name: t for t in get_all_tools()}
    assert list(tools) == ["web_search", "fetch_page", "crawl_url", "fact_check", "status"]
    ws = tools["web_search"].inputSchema["properties"]
    assert {"query", "top_k", "recency_days", "domain_allowlist", "domain_blocklist", "language", "fetch_full_content", "chunk_size", "rerank",
            "answer_mode", "local_only", "explain"} <= set(ws)
    assert ws["top_k"]["default"] == 5 and ws["rerank"]["default"] is True and ws["answer_mode"]["enum"] == ["snippets", "summary", "structured"]
    assert set(tools["crawl_url"].inputSchema["properties"]) >= {"url", "depth", "force"} and "api_key" not in ws
    assert tools["fact_check"].inputSchema["properties"]["top_k"]["default"] == 10
    assert all("api_key" in t.inputSchema["properties"] for t in get_all_tools(api_key_required=True))
    assert H.MCP_API_VERSION == "2025.1"
    import inspect

    src = inspect.getsource(H.ToolRuntime.__init__)
    for legacy in ("search", "search_local", "network_stats", "batch_search", "suggest", "register_webhook", "unregister_webhook", "analytics", "explain",
                   "search_history", "search_rag", "extract_answer", "ping", "credit_balance", "index_stats", "remove_url"):
        assert f'"{legacy}":' in src, legacy


# ------------------------------------------------------------------ Appendix A: HTTP
def test_http_route_surface():
    pytest.importorskip("fastapi")
    from infomesh_b200.api.local_api import create_admin_app
    from infomesh_b200.config import Config

    routes = {(m, r.path) for r in create_admin_app(Config()).routes for m in getattr(r, "methods", set()) if m in ("GET", "POST")}
    for path in ("/health", "/search", "/readiness", "/status", "/config", "/index/stats", "/index/compression", "/credits/balance", "/network/peers",
                 "/analytics", "/analytics/tools", "/metrics", "/openapi-spec", "/dashboard"):
        assert ("GET", path) in routes, path
    assert ("POST", "/config/reload") in routes


# ------------------------------------------------------------------ Appendix A: wire + on-disk
def test_wire_protocol_ids_and_message_numbers():
    from infomesh_b200.p2p import protocol as P

    ids = {f"/infomesh/{n}/1.0.0" for n in ("search", "index", "crawl", "replicate", "ping", "credit", "credit-sync", "index-submit", "pex", "llm")}
    assert set(P.ALL_PROTOCOLS) == ids
    expect = {0, 1, 10, 11, 20, 21, 22, 23, 30, 31, 32, 33, 34, 40, 41, 50, 51, 60, 61, 70, 71, 72, 73, 80, 81, 90, 91, 99, 100}
    assert expect <= {int(m) for m in P.MessageType}
    frame = P.encode_message(P.MessageType.PING, {"t": 1})
    import struct

    import msgpack

    n, = struct.unpack(">I", frame[:4])
    assert n == len(frame) - 4 and msgpack.unpackb(frame[4:]) == {"type": 0, "payload": {"t": 1}}
    assert P.keyword_to_dht_key("Rust").startswith("/infomesh/kw/") and P.keyword_to_dht_key("Rust") == P.keyword_to_dht_key("rust")
    assert P.url_to_dht_key("https://a/").startswith("/infomesh/url/")


def test_on_disk_names(tmp_path):
    from infomesh_b200 import runtime as RT
    from infomesh_b200.p2p.keys import ensure_keys

    assert (RT.PID_FILE_NAME, RT.STARTUP_LOCK_FILE_NAME, RT.RUNTIME_STATUS_FILE_NAME) == ("infomesh.pid", "infomesh.start.lock", "runtime_status.json")
    from infomesh_b200.config import Config
    from infomesh_b200.dashboard.utils import get_peer_id
    from infomesh_b200.services import AppContext
    from dataclasses import replace

    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False))
    with AppContext(cfg) as ctx:                      # the composition root must put the identity where every reader looks
        pid = ctx.key_pair.peer_id
    assert {"private.pem", "public.pem"} <= {p.name for p in (tmp_path / "keys").iterdir()} and not (tmp_path / "keys" / "keys").exists()
    assert ensure_keys(tmp_path).peer_id == pid and get_peer_id(cfg) == pid
    legacy = tmp_path / "old" / "keys" / "keys"     # layout written by early builds is migrated in place
    legacy.mkdir(parents=True)
    for f in (tmp_path / "keys").iterdir():
        if f.is_file():
            (legacy / f.name).write_bytes(f.read_bytes())
    assert ensure_keys(tmp_path / "old").peer_id == pid and (tmp_path / "old" / "keys" / "private.pem").exists() and not legacy.exists()
    import sqlite3

    from infomesh_b200.crawler.dedup import DeduplicatorDB
    from infomesh_b200.credits.ledger import CreditLedger
    from infomesh_b200.index.link_graph import LinkGraph
    from infomesh_b200.index.local_store import LocalStore

    def tables(path):
        with sqlite3.connect(path) as c:
            return {r[0] for r in c.execute("SELECT name FROM sqlite_master WHERE type IN ('table','view')")}

    LocalStore(tmp_path / "i.db").close()
    DeduplicatorDB(str(tmp_path / "d.db")).close()
    LinkGraph(str(tmp_path / "l.db")).close()
    CreditLedger(tmp_path / "c.db").close()
    assert {"documents", "documents_fts"} <= tables(tmp_path / "i.db") and "seen_urls" in tables(tmp_path / "d.db")
    assert {"links", "domain_authority"} <= tables(tmp_path / "l.db")
    assert {"credit_entries", "credit_spending", "credit_grace"} <= tables(tmp_path / "c.db")


# ------------------------------------------------------------------ Appendix B
def test_ranking_merge_and_passage_constants():
    from infomesh_b200.index import ranking as R
    from infomesh_b200.search import merge as M

    assert (R.WEIGHT_BM25, R.WEIGHT_FRESHNESS, R.WEIGHT_TRUST, R.WEIGHT_AUTHORITY, R.WEIGHT_TITLE_MATCH, R.WEIGHT_URL_PATH) == (.40, .15, .10, .15, .15, .05)
    assert R.FRESHNESS_HALF_LIFE_SECONDS == 7 * 86400 and R.MIN_FRESHNESS == .05 and R.DEFAULT_TRUST == .50
    assert R.normalize_bm25(3.0, max_score=1.0) == pytest.approx(0.75) and R.freshness_score(0, now=7 * 86400) == pytest.approx(0.5)
    assert R.freshness_score(0, now=10 * 365 * 86400) == .05 and M._RRF_K == 60
    from infomesh_b200.search.passage import score_passage, split_passages

    long = "Sentence about kernels. " * 60
    assert all(len(p) <= 500 for p in split_passages(long)) and score_passage("alpha gamma delta", ["alpha", "beta"]) == pytest.approx(0.5 + 0.1 / 3)


def test_simhash_keywords_embedding_reranker_constants():
    import hashlib

    from infomesh_b200.crawler import simhash as S
    from infomesh_b200.index import distributed as D
    from infomesh_b200.index import vector_store as V
    from infomesh_b200.p2p import dht
    from infomesh_b200.search import reranker as RR

    assert (S.HAMMING_THRESHOLD, S._NUM_BITS, S._SHINGLE_WIDTH) == (3, 64, 3)
    h = int.from_bytes(hashlib.md5(b"a b c").digest()[:8], "big")
    assert S.simhash("a b c") == h                                      # a single shingle's fingerprint is its own hash
    assert (D.MIN_KEYWORD_LENGTH, D.MAX_KEYWORDS_PER_DOC, D.MAX_POINTERS_PER_KEYWORD) == (2, 50, 100)
    assert (dht.MAX_POINTERS_PER_KEYWORD, dht.MAX_PUBLISHES_PER_KEYWORD_HR, dht._LOCK_TTL_SECONDS) == (100, 10, 300)
    assert (V.DEFAULT_MODEL, V.COLLECTION_NAME, V.MAX_EMBED_CHARS, V.PREVIEW_CHARS) == ("all-MiniLM-L6-v2", "infomesh_docs", 2000, 500)
    assert RR.MAX_RERANK_CANDIDATES == 20


def test_router_cache_crawl_constants():
    import inspect

    from infomesh_b200 import crawler
    from infomesh_b200.config import Config
    from infomesh_b200.crawler import worker as W
    from infomesh_b200.p2p import peer_profile as PP
    from infomesh_b200.p2p import routing as RO
    from infomesh_b200.search.cache import QueryCache

    assert (RO.SEARCH_TIMEOUT_MS, RO.MAX_FANOUT, RO.MAX_RESULTS_PER_PEER) == (5000, 5, 20)
    assert (PP.EMA_ALPHA, PP.MAX_HISTORY, PP.DIVERSITY_RATIO) == (0.3, 100, 0.2)
    t = PP.PeerProfileTracker()
    t.record("fast", 1.0), t.record("slow", 60_000.0)
    assert t.adaptive_timeout("fast", base_ms=5000) == 500.0 and t.adaptive_timeout("slow", base_ms=5000) == 5000.0
    sig = inspect.signature(QueryCache.__init__).parameters
    assert sig["max_size"].default == 1000 and sig["ttl_seconds"].default == 300.0
    c = Config().crawl
    assert (c.max_concurrent, c.politeness_delay, c.urls_per_hour, c.pending_per_domain) == (5, 1.0, 60, 10)
    assert crawler.MAX_RESPONSE_BYTES == 10 * 2 ** 20 and (W._MAX_RETRIES, W._RETRY_BACKOFF_BASE) == (2, 1.0)


def test_credit_trust_governor_constants():
    from infomesh_b200.credits import scheduling as SC
    from infomesh_b200.credits import types as CT
    from infomesh_b200.resources import governor as G
    from infomesh_b200.trust import scoring as TS

    w = CT.ACTION_WEIGHTS
    A = CT.ActionType
    assert (w[A.CRAWL], w[A.QUERY_PROCESS], w[A.DOC_HOSTING], w[A.NETWORK_UPTIME], w[A.LLM_SUMMARIZE_OWN], w[A.LLM_SUMMARIZE_PEER]) == (1.0, 0.5, 0.1, 0.5, 1.5, 2.0)
    assert [(f, c) for f, _, c in CT.TIER_THRESHOLDS] == [(1000.0, .033), (100.0, .050), (0.0, .100)]
    assert (CT.GRACE_PERIOD_HOURS, CT.DEBT_COST_MULTIPLIER, CT.LLM_CREDIT_CAP_RATIO, SC.OFF_PEAK_MULTIPLIER) == (72.0, 2.0, .60, 1.5)
    assert (TS.W_UPTIME, TS.W_CONTRIBUTION, TS.W_AUDIT, TS.W_SUMMARY) == (.15, .25, .40, .20)
    assert [t for t, _ in TS.TIER_THRESHOLDS][:3] == [.8, .5, .3]
    assert (TS.AUDIT_FAILURE_ISOLATION_THRESHOLD, TS.MAX_UPTIME_HOURS, TS.MAX_CONTRIBUTION_SCORE) == (3, 720, 5000.0)
    L = G.DegradeLevel
    assert G.classify(61, 0, 0) == L.WARNING and G.classify(0, 71, 0) == L.WARNING and G.classify(0, 0, .75) == L.WARNING
    assert G.classify(81, 0, 0) == L.OVERLOADED and G.classify(0, 86, 0) == L.OVERLOADED and G.classify(0, 0, .9) == L.OVERLOADED
    assert G.classify(91, 0, 0) == L.SEVERE and G.classify(0, 0, 1.0) == L.SEVERE and G.classify(96, 0, 0) == L.DEFENSIVE and G.classify(0, 0, 1.2) == L.DEFENSIVE
    assert G.classify(60, 70, .74) == L.NORMAL
    assert [G.throttle_for(lv, 50) for lv in (L.WARNING, L.OVERLOADED, L.SEVERE, L.DEFENSIVE)] == [0.5, 0.25, 0.0, 0.0]


def test_summariser_and_fts_limits():
    import inspect

    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.summarizer.engine import SummarizationEngine

    p = inspect.signature(SummarizationEngine.summarize).parameters
    assert p["max_tokens"].default == 512 and p["max_input_chars"].default == 8000
    with LocalStore() as st:
        st.add_document(url="https://a.example/", title="t", text="needle " * 30, raw_html_hash="r", text_hash="t")
        assert len(st.search("needle", limit=10 ** 6)) == 1 and st.search("needle", offset=10 ** 9) == []
        assert "<b>needle</b>" in st.search("needle")[0].snippet
