"""MCP tool runtime, tool schemas, sessions/webhooks (model: reference tests/test_mcp*.py)."""
import asyncio
import json
from dataclasses import replace

import pytest

from infomesh_b200.config import Config


def _ctx(tmp_path):
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.services import AppContext, index_document

    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False))
    ctx = AppContext(cfg)
    docs = [("https://docs.python.org/3/library/asyncio.html", "asyncio — Asynchronous I/O",
             "asyncio is a library to write concurrent code using the async await syntax. The event loop runs asynchronous tasks and callbacks. " * 3),
            ("https://doc.rust-lang.org/book/ownership.html", "Understanding Ownership",
             "Ownership is a set of rules that govern how a Rust program manages memory. Borrowing and lifetimes keep references valid. " * 3),
            ("https://example.org/python-history", "History of Python",
             "Python was created by Guido van Rossum and first released in 1991. Python emphasises code readability. " * 3)]
    for i, (u, t, x) in enumerate(docs):
        index_document(ParsedPage(url=u, title=t, text=x, language="en", raw_html_hash=f"r{i}", text_hash=f"t{i}"), ctx.store)
    return ctx


def test_tool_schemas_and_filters():
    from infomesh_b200.mcp import tools as T

    names = [s["name"] for s in T.tool_schemas()]
    assert names == ["web_search", "fetch_page", "crawl_url", "fact_check", "status"]
    ws = T.tool_schemas(api_key_required=True)[0]["inputSchema"]
    assert ws["required"] == ["query"] and "api_key" in ws["properties"] and ws["properties"]["answer_mode"]["enum"] == ["snippets", "summary", "structured"]
    tools = T.get_all_tools()
    assert tools[0].name == "web_search" and tools[0].annotations.readOnlyHint is True
    f = T.extract_filters({"recency_days": 2, "domain_allowlist": ["a.com"], "exclude_domains": ["b.com"], "language": "en"})
    assert f["include_domains"] == ["a.com"] and f["exclude_domains"] == ["b.com"] and f["language"] == "en" and f["date_from"] > 0
    assert T.check_api_key({"api_key": "k"}, "k") is None and T.check_api_key({}, "k").startswith("Error") and T.check_api_key({}, None) is None


def test_tool_runtime_end_to_end(tmp_path):
    from infomesh_b200.mcp.handlers import MCP_API_VERSION, ToolRuntime
    from infomesh_b200.persistence.store import PersistentStore

    ctx = _ctx(tmp_path)
    rt = ToolRuntime(ctx, pstore=PersistentStore(tmp_path / "p.db"))

    async def flow():
        out = {}
        out["ws"] = json.loads(await rt.call("web_search", {"query": "asyncio event loop", "top_k": 3}))
        out["ws_full"] = json.loads(await rt.call("web_search", {"query": "rust ownership", "fetch_full_content": True}))
        out["explain"] = json.loads(await rt.call("web_search", {"query": "asyncio", "explain": True}))
        out["rag"] = json.loads(await rt.call("web_search", {"query": "asyncio", "chunk_size": 80}))
        out["answer"] = json.loads(await rt.call("web_search", {"query": "python created", "answer_mode": "summary"}))
        out["text"] = await rt.call("search_local", {"query": "ownership rules", "session_id": "s1"})
        out["none"] = await rt.call("search_local", {"query": "zzzqqqxxx"})
        out["bad"] = await rt.call("web_search", {"query": "  "})
        out["fact"] = json.loads(await rt.call("fact_check", {"claim": "Python was created by Guido van Rossum in 1991"}))
        out["fetch"] = await rt.call("fetch_page", {"url": "https://example.org/python-history"})
        out["ssrf"] = await rt.call("fetch_page", {"url": "http://169.254.169.254/latest"})
        out["status"] = json.loads(await rt.call("status", {}))
        out["batch"] = json.loads(await rt.call("batch_search", {"queries": ["asyncio", "ownership", ""], "limit": 2, "format": "json"}))
        out["batch_text"] = await rt.call("batch_search", {"queries": ["asyncio", "ownership"], "limit": 2})
        out["eleven"] = await rt.call("batch_search", {"queries": ["asyncio"] * 11})
        out["too_many"] = await rt.call("batch_search", {"queries": []})
        out["suggest"] = json.loads(await rt.call("suggest", {"prefix": "own"}))
        out["hook_bad"] = await rt.call("register_webhook", {"url": "http://127.0.0.1/x"})
        out["ping"] = json.loads(await rt.call("ping", {}))
        out["credits"] = json.loads(await rt.call("credit_balance", {}))
        out["istats"] = json.loads(await rt.call("index_stats", {}))
        out["rm"] = await rt.call("remove_url", {"url": "https://example.org/python-history"})
        out["rm2"] = await rt.call("remove_url", {"url": "https://example.org/python-history"})
        out["unknown"] = await rt.call("nope", {})
        out["analytics"] = json.loads(await rt.call("analytics", {}))
        return out

    o = asyncio.run(flow())
    assert o["ws"]["results"][0]["url"].endswith("asyncio.html") and o["ws"]["api_version"] == MCP_API_VERSION and "quota" in o["ws"]
    assert "full_text" in o["ws_full"]["results"][0]
    assert o["explain"]["results"][0]["weights"]["bm25"] == 0.4 and o["rag"]["chunks"] and o["rag"]["context_window"]
    assert o["answer"]["answers"] and o["answer"]["sources"]
    assert "Understanding Ownership" in o["text"] and "Attribution:" in o["text"] and "s1" in rt.sessions
    assert o["none"].startswith("No results found") and o["bad"].startswith("Error [INVALID_PARAM]")
    assert o["fact"]["verdict"] == "supported" and o["fact"]["supporting"] >= 1
    assert "History of Python" in o["fetch"] and "COPYRIGHT NOTICE" in o["fetch"] and o["ssrf"].startswith("Error [SSRF_BLOCKED]")
    assert o["status"]["documents_indexed"] == 3 and o["status"]["status"] == "ok" and o["status"]["gpu"] == {"enabled": False}
    assert len(o["batch"]["batch_results"]) == 3 and o["batch"]["batch_results"][2] == {"query": "", "error": "invalid"}
    assert o["batch"]["batch_results"][0]["query"] == "asyncio" and o["too_many"].startswith("Error [INVALID_PARAM]")
    assert o["batch_text"].startswith("--- Query 1: asyncio ---\nFound ") and "--- Query 2: ownership ---" in o["batch_text"]
    assert o["eleven"].count("--- Query ") == 10                                   # extra queries are ignored, not an error
    assert any("Ownership" in s for s in o["suggest"]["suggestions"]) and o["hook_bad"].startswith("Error [SSRF_BLOCKED]")
    assert o["ping"]["status"] == "ok" and "balance" in o["credits"] and o["istats"]["document_count"] == 3
    assert o["rm"].startswith("Removed") and o["rm2"].startswith("Error [NOT_FOUND]") and o["unknown"].startswith("Error [NOT_FOUND]")
    assert o["analytics"]["total_searches"] >= 4 and o["analytics"]["persistent"]["total_fetches"] == 1 and o["analytics"]["tools"]["web_search"] == 6
    ctx.close()


def test_sessions_and_webhook_registry():
    from infomesh_b200.mcp.session import SessionStore, WebhookRegistry

    ss = SessionStore(max_size=2, ttl_seconds=100)
    a = ss.get_or_create("a")
    a.updated_at -= 1
    ss.get_or_create("b")
    ss.get_or_create("c")
    assert len(ss) == 2 and "a" not in ss and ss.get_or_create("b") is ss.get_or_create("b")
    wr = WebhookRegistry(max_registrations=1)
    assert wr.register("https://hooks.example.com/x") is None and wr.register("https://hooks.example.com/x") is None
    assert "Max webhooks" in wr.register("https://hooks.example.com/y") and "blocked" in WebhookRegistry().register("http://10.0.0.1/")
    assert wr.unregister("https://hooks.example.com/x") and not wr.unregister("nope") and asyncio.run(wr.notify("e", {})) == 0


def test_mcp_server_lists_and_calls_tools(tmp_path, monkeypatch):
    from mcp.types import CallToolRequest, CallToolRequestParams, ListToolsRequest

    from infomesh_b200.mcp import server as S

    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp_path), index=replace(base.index, db_path=tmp_path / "index.db", vector_search=False))
    app, ctx, pstore = S._create_app(cfg, api_key="sekrit")

    async def go():
        tools = await app.request_handlers[ListToolsRequest](ListToolsRequest(method="tools/list"))
        denied = await app.request_handlers[CallToolRequest](CallToolRequest(method="tools/call", params=CallToolRequestParams(name="status", arguments={})))
        ok = await app.request_handlers[CallToolRequest](CallToolRequest(method="tools/call", params=CallToolRequestParams(name="status", arguments={"api_key": "sekrit"})))
        return tools, denied, ok

    tools, denied, ok = asyncio.run(go())
    assert [t.name for t in tools.root.tools][:2] == ["web_search", "fetch_page"] and "api_key" in tools.root.tools[0].inputSchema["properties"]
    assert "invalid or missing api_key" in denied.root.content[0].text and json.loads(ok.root.content[0].text)["status"] == "ok"
    pstore.close()
    ctx.close()


def test_admin_api_routes(tmp_path, monkeypatch):
    from fastapi.testclient import TestClient

    from infomesh_b200.api import extensions as X
    from infomesh_b200.api.local_api import _format_duration, create_admin_app
    from infomesh_b200.mcp.handlers import ToolRuntime

    ctx = _ctx(tmp_path)
    app = create_admin_app(ctx.config, runtime=ToolRuntime(ctx))
    c = TestClient(app)
    assert c.get("/health").json() == {"status": "ok"} and c.get("/health?detail=1").json()["db"] == "ok"
    assert c.get("/readiness").status_code == 200
    s = c.get("/search", params={"q": "asyncio event loop", "limit": 50}).json()
    assert s["results"] and s["results"][0]["url"].endswith("asyncio.html") and len(s["results"][0]["snippet"]) <= 300
    assert c.get("/search").json()["error"] == "query required"
    assert c.get("/status").json()["index"]["document_count"] == 3 and c.get("/index/compression").json()["documents"] == 3
    cfg = c.get("/config").json()
    assert isinstance(cfg["node"]["data_dir"], str) and "crawl" in cfg
    assert c.post("/config/reload").json() == {"status": "reloaded"}
    import time as _t
    _t.sleep(1.05)                       # stay under the 10 requests / second limiter
    assert "balance" in c.get("/credits/balance").json() and c.get("/network/peers").json()["connected"] == 0
    assert c.get("/analytics").json()["total_searches"] == 1 and c.get("/analytics/tools").json()["tool_usage"]["web_search"] == 1
    m = c.get("/metrics")
    assert "infomesh_search_total 1.0" in m.text and m.headers["x-frame-options"] == "DENY"
    assert c.get("/gpu/stats").json() == {"enabled": False} and c.post("/index/submit", content=b"x").status_code == 404
    spec = c.get("/openapi-spec").json()
    assert spec["openapi"] == "3.1.0" and "/search" in spec["paths"] and "web_search" in spec["components"]["schemas"]
    _t.sleep(1.05)
    assert "<title>InfoMesh node</title>" in c.get("/dashboard").text
    monkeypatch.setenv("INFOMESH_API_KEY", "k1")
    assert c.get("/health").status_code == 401 and c.get("/health", headers={"x-api-key": "k1"}).status_code == 200
    monkeypatch.delenv("INFOMESH_API_KEY")
    _t.sleep(1.05)
    codes = [c.get("/health").status_code for _ in range(15)]
    assert 429 in codes
    assert TestClient(app, client=("8.8.8.8", 1234)).get("/health").status_code == 403
    assert _format_duration(59) == "59s" and _format_duration(3700).startswith("1h")
    rl = X.RateLimiter(X.RateLimitConfig(requests_per_minute=2))
    assert rl.check("a") and rl.check("a") and not rl.check("a") and rl.remaining("a") == 0 and rl.remaining("b") == 2
    km = X.APIKeyManager()
    k = km.create_key("ci", expires_in_days=1, permissions=["search"])
    assert k.key.startswith("im_") and len(k.key) == 51 and km.validate(k.key) is k
    k2 = km.rotate(k.key)
    assert km.validate(k.key) is None and km.validate(k2.key).permissions == ["search"] and len(km.list_keys()) == 2
    assert "complete -F _infomesh_complete infomesh" in X.generate_bash_completion() and "#compdef infomesh" in X.generate_zsh_completion()
    ctx.close()
