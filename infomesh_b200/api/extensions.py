"""API extensions: hand-written OpenAPI 3.1 document for the admin + MCP-HTTP surface, sliding-window rate limiter,
in-memory API keys (``im_<48 hex>``, expiry, permissions, rotate), bash/zsh completion scripts
(reference infomesh/api/extensions.py:27-420)."""
from __future__ import annotations

import secrets
import time
from collections import defaultdict, deque
from dataclasses import dataclass, field
from typing import Any

from infomesh_b200 import __version__
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_ROUTES: tuple[tuple[str, str, str, tuple[tuple[str, str, str], ...]], ...] = (
    ("get", "/health", "Liveness probe (detail=1 for disk / memory / runtime checks)", (("detail", "string", "any value enables detail"),)),
    ("get", "/readiness", "Readiness probe — index database accessible", ()),
    ("get", "/search", "Search the local index (GPU hybrid pipeline when attached)", (("q", "string", "query"), ("limit", "integer", "max results (<= 20)"))),
    ("get", "/status", "Node status overview", ()),
    ("get", "/config", "Current configuration with secrets redacted", ()),
    ("post", "/config/reload", "Reload configuration from disk", ()),
    ("get", "/index/stats", "Index statistics", ()),
    ("get", "/index/compression", "Index compression statistics", ()),
    ("get", "/credits/balance", "Local credit balance and earnings", ()),
    ("get", "/network/peers", "Connected peers summary", ()),
    ("get", "/analytics", "Search / crawl / fetch counters", ()),
    ("get", "/analytics/tools", "MCP tool usage breakdown", ()),
    ("get", "/metrics", "Prometheus exposition text", ()),
    ("get", "/gpu/stats", "GPU index residency, batch size, graph state", ()),
    ("post", "/index/submit", "Accept a msgpack IndexSubmit frame from a crawler node", ()),
    ("get", "/dashboard", "HTML dashboard", ()),
)


def generate_openapi_spec() -> dict[str, Any]:
    paths: dict[str, Any] = {}
    for method, path, summary, params in _ROUTES:
        op: dict[str, Any] = {"summary": summary, "responses": {"200": {"description": "OK"}}}
        if params:
            op["parameters"] = [{"name": n, "in": "query", "required": n == "q", "schema": {"type": t}, "description": d} for n, t, d in params]
        paths.setdefault(path, {})[method] = op
    try:
        from infomesh_b200.mcp.tools import tool_schemas

        tools = {t["name"]: t["inputSchema"] for t in tool_schemas()}
    except Exception:  # noqa: BLE001
        tools = {}
    # the MCP endpoint and the two generic schemas that clients generated from the reference's spec expect
    paths["/mcp"] = {"post": {"summary": "MCP endpoint (streamable HTTP transport, port 8081)",
                              "requestBody": {"required": True, "content": {"application/json": {"schema": {"type": "object"}}}},
                              "responses": {"200": {"description": "MCP response"}}}}
    tools = {**tools,
             "SearchRequest": {"type": "object", "required": ["query"],
                               "properties": {"query": {"type": "string", "description": "Search query"},
                                              "limit": {"type": "integer", "default": 10, "maximum": 50},
                                              "format": {"type": "string", "enum": ["text", "json"], "default": "text"},
                                              "language": {"type": "string", "description": "ISO 639-1 code"}}},
             "SearchResult": {"type": "object",
                              "properties": {"url": {"type": "string"}, "title": {"type": "string"}, "snippet": {"type": "string"},
                                             "score": {"type": "number"}, "domain": {"type": "string"}, "crawled_at": {"type": "number"}}}}
    return {"openapi": "3.1.0",
            "info": {"title": "InfoMesh API", "version": __version__, "description": "Local admin API and MCP tool schemas of an InfoMesh node."},
            "servers": [{"url": "http://127.0.0.1:8080", "description": "admin API"}, {"url": "http://127.0.0.1:8081", "description": "MCP streamable HTTP (/mcp)"}],
            "paths": paths,
            "components": {"securitySchemes": {"ApiKeyAuth": {"type": "apiKey", "in": "header", "name": "x-api-key"}}, "schemas": tools},
            "security": [{"ApiKeyAuth": []}]}


@dataclass
class RateLimitConfig:
    requests_per_minute: int = 60
    burst_size: int = 10
    window_seconds: float = 60.0


class RateLimiter:
    def __init__(self, config: RateLimitConfig | None = None):
        self._cfg = config or RateLimitConfig()
        self._hits: dict[str, deque[float]] = defaultdict(deque)

    def _live(self, key: str, now: float) -> deque[float]:
        q = self._hits[key]
        while q and now - q[0] >= self._cfg.window_seconds:
            q.popleft()
        return q

    def check(self, key: str = "global") -> bool:
        now = time.time()
        q = self._live(key, now)
        if len(q) >= self._cfg.requests_per_minute:
            return False
        q.append(now)
        return True

    def remaining(self, key: str = "global") -> int:
        return max(0, self._cfg.requests_per_minute - len(self._live(key, time.time()))) if key in self._hits else self._cfg.requests_per_minute

    def reset(self, key: str = "global") -> None:
        self._hits.pop(key, None)


@dataclass
class APIKey:
    key: str
    name: str
    created_at: float
    expires_at: float | None = None
    permissions: list[str] = field(default_factory=list)
    active: bool = True

    def is_valid(self, *, now: float | None = None) -> bool:
        return self.active and (self.expires_at is None or (now or time.time()) <= self.expires_at)


class APIKeyManager:
    def __init__(self):
        self._keys: dict[str, APIKey] = {}

    def create_key(self, name: str, *, expires_in_days: int | None = None, permissions: list[str] | None = None) -> APIKey:
        now = time.time()
        k = APIKey(f"im_{secrets.token_hex(24)}", name, now, now + expires_in_days * 86400 if expires_in_days is not None else None,
                   list(permissions or []))
        self._keys[k.key] = k
        logger.info("api_key_created", name=name)
        return k

    def validate(self, key: str) -> APIKey | None:
        k = self._keys.get(key)
        return k if k is not None and k.is_valid() else None

    def revoke(self, key: str) -> bool:
        k = self._keys.get(key)
        if k is None:
            return False
        k.active = False
        return True

    def list_keys(self) -> list[APIKey]:
        return list(self._keys.values())

    def rotate(self, old_key: str) -> APIKey | None:
        old = self._keys.get(old_key)
        if old is None:
            return None
        self.revoke(old_key)
        return self.create_key(old.name, permissions=old.permissions)


_COMMANDS = ["start", "stop", "update", "status", "crawl", "mcp", "dashboard", "search", "index", "config", "keys", "peer", "feeds",
             "feedback", "doctor", "bench"]
_SUBCOMMANDS = {"index": ["stats", "export", "import", "import-wet", "import-urls", "gpu-build"], "config": ["show", "set", "github"],
                "keys": ["export", "rotate"], "peer": ["list", "add", "remove", "test"], "feeds": ["import", "list"], "feedback": ["stats", "top-urls"]}


def get_completion_commands() -> list[str]:
    return list(_COMMANDS)


def generate_bash_completion() -> str:
    cases = "\n".join(f'        {cmd}) COMPREPLY=( $(compgen -W "{" ".join(subs)}" -- "$cur") ) ;;' for cmd, subs in _SUBCOMMANDS.items())
    return f'''# bash completion for infomesh
_infomesh_complete() {{
    local cur prev
    cur="${{COMP_WORDS[COMP_CWORD]}}"
    prev="${{COMP_WORDS[COMP_CWORD-1]}}"
    if [ "$COMP_CWORD" -eq 1 ]; then
        COMPREPLY=( $(compgen -W "{" ".join(_COMMANDS)}" -- "$cur") )
        return 0
    fi
    case "$prev" in
{cases}
        *) COMPREPLY=() ;;
    esac
}}
complete -F _infomesh_complete infomesh
'''


def generate_zsh_completion() -> str:
    subs = "\n".join(f"        {cmd}) _values 'subcommand' {' '.join(s)} ;;" for cmd, s in _SUBCOMMANDS.items())
    return f'''#compdef infomesh
_infomesh() {{
    local -a commands
    commands=({" ".join(_COMMANDS)})
    if (( CURRENT == 2 )); then
        _describe 'command' commands
        return
    fi
    case "$words[2]" in
{subs}
    esac
}}
_infomesh "$@"
'''
