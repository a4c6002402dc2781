"""mcp/tools.py filters + schemas, mcp/session.py (sessions, analytics, webhooks), ToolRuntime edge cases."""
import asyncio
import json
import time
from dataclasses import replace

from infomesh_b200.config import Config
from infomesh_b200.mcp import tools as T
from infomesh_b200.mcp.session import AnalyticsTracker, SessionStore


def make_ctx(tmp_path):
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.services import AppContext, index_document

    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False))
    ctx = AppContext(cfg)
    for i in range(6):
        index_document(ParsedPage(url=f"https://ex.org/page{i}", title=f"Kernel tuning part {i}",
                                  text=f"Part {i}: tuning tensor core kernels requires profiling occupancy and memory bandwidth. " * 4,
                                  language="en", raw_html_hash=f"r{i}", text_hash=f"t{i}"), ctx.store)
    return ctx


# ------------------------------------------------------------------ filters / schemas
def test_extract_filters_both_spellings():
    now = time.time()
    f = T.extract_filters({"language": "ko", "recency_days": "3", "domain_allowlist": ["a.com"], "domain_blocklist": ["b.org"], "date_to": 5})
    assert f["language"] == "ko" and abs(f["date_from"] - (now - 3 * 86400)) < 5 and f["include_domains"] == ["a.com"]
    assert f["exclude_domains"] == ["b.org"] and f["date_to"] == 5.0
    legacy = T.extract_filters({"date_from": 10, "include_domains": ["x.io"], "exclude_domains": [], "language": ""})
    assert legacy == {"date_from": 10.0, "include_domains": ["x.io"]}
    assert T.extract_filters({"recency_days": "soon"}) == {} and T.extract_filters({"recency_days": -1}) == {}


def test_tool_schemas_cover_the_public_surface():
    names = {t["name"] for t in T.tool_schemas()}
    assert {"web_search", "fetch_page", "crawl_url", "fact_check", "status"} <= names
    for t in T.tool_schemas(api_key_required=True):
        assert t["inputSchema"]["type"] == "object" and "api_key" in t["inputSchema"]["properties"]
        assert t["description"]
    ws = next(t for t in T.tool_schemas() if t["name"] == "web_search")
    assert "query" in ws["inputSchema"]["required"] and "api_key" not in ws["inputSchema"]["properties"]


# ------------------------------------------------------------------ sessions / analytics
def test_session_ttl_expiry_and_capacity():
    ss = SessionStore(max_size=3, ttl_seconds=0.05)
    a = ss.get_or_create("a")
    a.last_query = "q"
    assert ss.get_or_create("a") is a
    time.sleep(0.06)
    fresh = ss.get_or_create("a")
    assert fresh is not a and fresh.last_query == ""
    for k in "bcd":
        ss.get_or_create(k)
    assert len(ss) <= 3 and "d" in ss


def test_analytics_tracker_is_consistent_under_concurrency():
    tr = AnalyticsTracker()

    async def go():
        await asyncio.gather(*(tr.record_search(10.0 + i) for i in range(50)), *(tr.record_crawl() for _ in range(7)), tr.record_fetch())

    asyncio.run(go())
    tr.record_tool("web_search")
    tr.record_tool("web_search")
    d = tr.to_dict()
    assert d["total_searches"] == 50 and d["total_crawls"] == 7 and d["total_fetches"] == 1 and d["avg_latency_ms"] == 34.5
    assert tr.tool_calls == {"web_search": 2}


# ------------------------------------------------------------------ runtime edge cases
def test_runtime_clamps_params_caches_and_pages(tmp_path):
    from infomesh_b200.mcp.handlers import ToolRuntime

    ctx = make_ctx(tmp_path)
    rt = ToolRuntime(ctx)

    async def flow():
        big = json.loads(await rt.call("search_local", {"query": "kernel tuning", "limit": 9999, "format": "json"}))
        one = json.loads(await rt.call("search_local", {"query": "kernel tuning", "limit": "1", "format": "json"}))
        junk = json.loads(await rt.call("search_local", {"query": "kernel tuning", "limit": "many", "format": "json"}))
        page2 = json.loads(await rt.call("search_local", {"query": "kernel tuning", "limit": 2, "offset": 2, "format": "json"}))
        first = await rt.call("search_local", {"query": "Kernel Tuning ", "limit": 3})
        again = await rt.call("search_local", {"query": "kernel tuning", "limit": 3})
        return big, one, junk, page2, first, again

    big, one, junk, page2, first, again = asyncio.run(flow())
    assert len(big["results"]) == 6 and len(one["results"]) == 1 and 1 <= len(junk["results"]) <= 10
    assert len(page2["results"]) == 2 and {r["url"] for r in page2["results"]}.isdisjoint({r["url"] for r in big["results"][:2]})
    assert first == again and rt.query_cache.stats.hits >= 1                                   # normalised query -> cache hit
    ctx.close()


def test_runtime_history_explain_rag_answers(tmp_path):
    from infomesh_b200.mcp.handlers import ToolRuntime

    ctx = make_ctx(tmp_path)
    rt = ToolRuntime(ctx)

    async def flow():
        await rt.call("search_local", {"query": "occupancy profiling", "session_id": "sess"})
        hist = rt.sessions.get_or_create("sess").last_query
        nohist = await rt.call("search_history", {})
        expl = json.loads(await rt.call("explain", {"query": "memory bandwidth"}))
        rag = json.loads(await rt.call("search_rag", {"query": "tensor core kernels", "max_chunks": 2}))
        ans = json.loads(await rt.call("extract_answer", {"query": "tuning tensor core kernels"}))
        bad = await rt.call("extract_answer", {"query": ""})
        return hist, nohist, expl, rag, ans, bad

    hist, nohist, expl, rag, ans, bad = asyncio.run(flow())
    assert hist == "occupancy profiling" and json.loads(nohist) == {"history": []}            # no persistent store attached
    assert expl["results"] and "bm25" in json.dumps(expl) and len(rag["chunks"]) <= 2 and rag["query"]
    assert ans["answers"] and bad.startswith("Error [INVALID_PARAM]")
    ctx.close()


def test_runtime_auth_and_unknown_tool_render_errors(tmp_path):
    from infomesh_b200.mcp.handlers import ErrorCode, ToolError, ToolRuntime

    e = ToolError(ErrorCode.RATE_LIMITED, "slow down", hint="wait 60 s")
    assert e.render() == "Error [RATE_LIMITED]: slow down\nHint: wait 60 s" and ToolError("X", "m").render() == "Error [X]: m"
    ctx = make_ctx(tmp_path)
    rt = ToolRuntime(ctx)

    async def boom(args):
        raise RuntimeError("handler exploded")

    rt._handlers["boom"] = boom
    out = asyncio.run(rt.call("boom", {}))
    assert out.startswith("Error [INTERNAL]") and "handler exploded" not in out                # internals are not leaked
    assert "web_search" in rt.tool_names and asyncio.run(rt.call("nope")).startswith("Error [NOT_FOUND]")
    ctx.close()


def test_functional_handlers_wrap_the_same_runtime(tmp_path):
    """Reference-style free functions (explicit collaborators as keyword arguments) return MCP text content."""
    import asyncio
    import json

    from infomesh_b200.credits.ledger import ActionType, CreditLedger
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.mcp import handlers as H
    from infomesh_b200.mcp.session import AnalyticsTracker
    from infomesh_b200.services import index_document

    with LocalStore(tmp_path / "i.db") as store:
        index_document(ParsedPage(url="https://e.org/tmem", title="Tensor memory", text="Tensor memory holds accumulators for tcgen05 MMA. " * 6, language="en",
                                  raw_html_hash="r", text_hash="t"), store)
        led = CreditLedger()
        led.record_action(ActionType.CRAWL, 5)
        an = AnalyticsTracker()
        out = asyncio.run(H.handle_search("search_local", {"query": "tensor memory", "format": "json"}, store=store, link_graph=None, ledger=led, analytics=an))
        data = json.loads(out[0].text)
        assert out[0].type == "text" and data["results"][0]["url"] == "https://e.org/tmem" and an.tool_calls["search_local"] == 1
        assert led.stats().total_spent > 0                                   # the search was charged
        web = asyncio.run(H.handle_web_search({"query": "tensor memory", "local_only": True}, store=store, ledger=None))
        assert "e.org/tmem" in web[0].text
        assert "e.org/tmem" in H.handle_suggest({"prefix": "Tens"}, store=store)[0].text or "Tensor" in H.handle_suggest({"prefix": "Tens"}, store=store)[0].text
        assert json.loads(H.handle_ping()[0].text)["api_version"] == H.MCP_API_VERSION
        assert json.loads(H.handle_index_stats({"format": "json"}, store=store, vector_store=None)[0].text)["document_count"] == 1
        bal = H.handle_credit_balance({"format": "json"}, ledger=led, credit_sync_manager=None)[0].text
        assert "balance" in bal
        status = H.handle_status({}, store=store, vector_store=None, link_graph=None, ledger=led, scheduler=None, p2p_node=None, distributed_index=None,
                                 analytics=an)[0].text
        assert json.loads(status)["documents_indexed"] == 1
        assert "Documents: 1" in H.handle_index_stats({"format": "text"}, store=store, vector_store=None)[0].text
        assert H.handle_credit_balance({"format": "text"}, ledger=None)[0].text.splitlines()[2:] == ["Balance: 0", "State: normal", "Search cost: 0.1"]
        expl = json.loads(asyncio.run(H.handle_explain({"query": "tensor"}, store=store, link_graph=None))[0].text)
        assert expl["results"] and "weights" in expl["results"][0]
        cached = asyncio.run(H.handle_fetch({"url": "https://e.org/tmem"}, store=store, worker=None, vector_store=None))[0].text
        assert "Tensor memory" in cached
        assert "Removed from index" in H.handle_remove_url({"url": "https://e.org/tmem"}, store=store)[0].text
        assert "Error [" in H.handle_remove_url({"url": "https://e.org/tmem"}, store=store)[0].text

        class Broke:
            def search_allowance(self):
                raise RuntimeError("db locked")

        H.deduct_search_cost(Broke()), H.deduct_search_cost(None)          # never raises

        async def inside_loop():
            return H.handle_ping(), H.handle_index_stats({}, store=store, vector_store=None)

        assert len(asyncio.run(inside_loop())) == 2
