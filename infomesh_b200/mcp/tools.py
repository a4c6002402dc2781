"""Tool schemas: five intent-level tools (web_search, fetch_page, crawl_url, fact_check, status); ``api_key`` is added
to every schema when the server runs with INFOMESH_API_KEY (reference infomesh/mcp/tools.py:26-334).  Schemas are plain
dicts; :func:`get_all_tools` wraps them in ``mcp.types.Tool`` when the ``mcp`` package is importable."""
from __future__ import annotations

import hmac
import time
from typing import Any


def _prop(kind: str, desc: str, **extra: Any) -> dict[str, Any]:
    return {"type": kind, "description": desc, **extra}


TOOL_SPECS: list[dict[str, Any]] = [
    {"name": "web_search", "title": "Web Search", "read_only": True, "open_world": True,
     "description": ("Search the web via InfoMesh P2P search engine. Returns ranked results with optional full content, RAG "
                     "chunking, and score explanations. Example: web_search(query='python asyncio tutorial')"),
     "properties": {
         "query": _prop("string", "Search query string"),
         "top_k": _prop("integer", "Number of results to return", default=5),
         "recency_days": _prop("integer", "Filter results published within the last N days"),
         "domain_allowlist": _prop("array", "Only include results from these domains", items={"type": "string"}),
         "domain_blocklist": _prop("array", "Exclude results from these domains", items={"type": "string"}),
         "language": _prop("string", "ISO 639-1 language code (e.g. 'en', 'ko', 'ja')"),
         "fetch_full_content": _prop("boolean", "Fetch and return full article text for each result", default=False),
         "chunk_size": _prop("integer", "Chunk size for RAG context splitting. When set, returns source-attributed chunks."),
         "rerank": _prop("boolean", "Apply semantic re-ranking (cross-encoder on GPU nodes, local LLM otherwise)", default=True),
         "answer_mode": _prop("string", "Response mode: 'snippets' (ranked results), 'summary' (answer extraction), 'structured' "
                                        "(JSON with scores and metadata)", enum=["snippets", "summary", "structured"], default="snippets"),
         "local_only": _prop("boolean", "Search local index only (offline, <10ms)", default=False),
         "explain": _prop("boolean", "Include score breakdown (BM25, freshness, trust, authority) per result", default=False)},
     "required": ["query"]},
    {"name": "fetch_page", "title": "Fetch Page", "read_only": True, "open_world": True,
     "description": "Fetch full text of a specific URL. Returns cached content or crawls live. Max 100KB. Example: fetch_page(url='https://...')",
     "properties": {"url": _prop("string", "URL to fetch")}, "required": ["url"]},
    {"name": "crawl_url", "title": "Crawl URL", "read_only": False,
     "description": "Add a URL to the crawl queue and index it. Rate limited to 60/hour. Example: crawl_url(url='https://example.com', depth=1)",
     "properties": {"url": _prop("string", "URL to crawl"),
                    "depth": _prop("integer", "Link-follow depth (0=this page only). Stays within same domain.", default=0),
                    "force": _prop("boolean", "Force re-crawl even if previously crawled (bypasses all dedup checks)", default=False)},
     "required": ["url"]},
    {"name": "fact_check", "title": "Fact Check", "read_only": True, "open_world": True,
     "description": ("Cross-reference a claim against indexed web content. Returns verdict with supporting/contradicting sources. "
                     "Example: fact_check(claim='Python was created in 1991')"),
     "properties": {"claim": _prop("string", "Claim to verify"), "top_k": _prop("integer", "Max sources to check", default=10)},
     "required": ["claim"]},
    {"name": "status", "title": "Node Status", "read_only": True, "idempotent": True,
     "description": "Node status: index size, peer count, credit balance, search quota, and analytics. Example: status()",
     "properties": {}, "required": []},
]
TOOL_NAMES = tuple(t["name"] for t in TOOL_SPECS)
LEGACY_TOOL_NAMES = ("search", "search_local", "network_stats", "batch_search", "suggest", "register_webhook", "unregister_webhook",
                     "analytics", "explain", "search_history", "search_rag", "extract_answer", "ping", "credit_balance", "index_stats",
                     "remove_url")


def tool_schemas(*, api_key_required: bool = False) -> list[dict[str, Any]]:
    """JSON-schema view (used by the HTTP ``/openapi-spec`` route and the SDK)."""
    out = []
    for spec in TOOL_SPECS:
        props = {k: dict(v) for k, v in spec["properties"].items()}
        if api_key_required:
            props["api_key"] = _prop("string", "API key (required when INFOMESH_API_KEY is set)")
        schema: dict[str, Any] = {"type": "object", "properties": props}
        if spec["required"]:
            schema["required"] = list(spec["required"])
        out.append({"name": spec["name"], "description": spec["description"], "inputSchema": schema,
                    "annotations": {"title": spec["title"], "readOnlyHint": spec.get("read_only", False),
                                    "openWorldHint": spec.get("open_world", False), "idempotentHint": spec.get("idempotent", False)}})
    return out


def get_all_tools(*, api_key_required: bool = False) -> list[Any]:
    from mcp.types import Tool, ToolAnnotations

    return [Tool(name=s["name"], description=s["description"], inputSchema=s["inputSchema"],
                 annotations=ToolAnnotations(**{k: v for k, v in s["annotations"].items() if v or k == "title"}))
            for s in tool_schemas(api_key_required=api_key_required)]


def extract_filters(args: dict[str, Any]) -> dict[str, Any]:
    """Accepts both spellings: recency_days / domain_allowlist / domain_blocklist and the legacy
    date_from / date_to / include_domains / exclude_domains."""
    f: dict[str, Any] = {}
    if isinstance(args.get("language"), str) and args["language"]:
        f["language"] = args["language"]
    if args.get("recency_days") is not None:
        try:
            days = int(args["recency_days"])
            if days > 0:
                f["date_from"] = time.time() - days * 86400
        except (TypeError, ValueError):
            pass
    elif args.get("date_from") is not None:
        f["date_from"] = float(args["date_from"])
    if args.get("date_to") is not None:
        f["date_to"] = float(args["date_to"])
    inc = args.get("domain_allowlist") or args.get("include_domains")
    exc = args.get("domain_blocklist") or args.get("exclude_domains")
    if isinstance(inc, list) and inc:
        f["include_domains"] = inc
    if isinstance(exc, list) and exc:
        f["exclude_domains"] = exc
    return f


def check_api_key(arguments: dict[str, Any], expected_key: str | None) -> str | None:
    if expected_key is None:
        return None
    got = arguments.get("api_key")
    if not isinstance(got, str) or not hmac.compare_digest(got.encode(), expected_key.encode()):
        return "Error: invalid or missing api_key"
    return None
