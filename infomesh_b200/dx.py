"""Developer-experience helpers: object plugins with setup/teardown, the pluggable search tokenizer, the MCP tool
reference generator and changelog rendering (reference infomesh/dx.py:22-334)."""
from __future__ import annotations

import importlib
import re
from dataclasses import dataclass, field
from typing import Any, Protocol

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class PluginProtocol(Protocol):
    name: str

    def setup(self, app: Any) -> None: ...

    def teardown(self) -> None: ...


@dataclass
class PluginInfo:
    name: str
    version: str = "0.0.0"
    description: str = ""
    author: str = ""
    module_path: str = ""
    enabled: bool = True


class PluginManager:
    def __init__(self):
        self._plugins: dict[str, PluginProtocol] = {}
        self._info: dict[str, PluginInfo] = {}

    def register(self, plugin: PluginProtocol, *, info: PluginInfo | None = None) -> None:
        self._plugins[plugin.name] = plugin
        self._info[plugin.name] = info or PluginInfo(name=plugin.name)

    def load_module(self, module_path: str) -> bool:
        """The module must expose a ``plugin`` object."""
        try:
            plugin = getattr(importlib.import_module(module_path), "plugin", None)
        except Exception as exc:  # noqa: BLE001
            logger.error("plugin_load_error", module=module_path, error=str(exc))
            return False
        if plugin is None:
            logger.warning("plugin_missing_plugin_var", module=module_path)
            return False
        self.register(plugin, info=PluginInfo(name=plugin.name, module_path=module_path))
        return True

    def _each(self, what: str, call) -> None:
        for name, plugin in self._plugins.items():
            if not self._info[name].enabled:
                continue
            try:
                call(plugin)
            except Exception as exc:  # noqa: BLE001
                logger.error(f"plugin_{what}_error", name=name, error=str(exc))

    def setup_all(self, app: Any) -> None:
        self._each("setup", lambda p: p.setup(app))

    def teardown_all(self) -> None:
        self._each("teardown", lambda p: p.teardown())

    def list_plugins(self) -> list[PluginInfo]:
        return list(self._info.values())


class TokenizerHook(Protocol):
    def tokenize(self, text: str) -> list[str]: ...


class DefaultTokenizer:
    _word = re.compile(r"\w+")

    def tokenize(self, text: str) -> list[str]:
        return [w for w in self._word.findall(text.lower()) if len(w) >= 2]


_active_tokenizer: TokenizerHook = DefaultTokenizer()


def set_tokenizer(tokenizer: TokenizerHook) -> None:
    global _active_tokenizer
    _active_tokenizer = tokenizer
    logger.info("custom_tokenizer_set", type=type(tokenizer).__name__)


def get_tokenizer() -> TokenizerHook:
    return _active_tokenizer


_SEARCH_PARAMS = ("query (str), limit (int), format (str), language (str), date_from (float), date_to (float), "
                  "include_domains (list), exclude_domains (list), offset (int), snippet_length (int), session_id (str)")
MCP_TOOLS_GUIDE: list[dict[str, str]] = [
    {"name": "search", "description": "Full network search", "params": _SEARCH_PARAMS,
     "example": '{"query": "python asyncio", "limit": 5, "format": "json"}'},
    {"name": "search_local", "description": "Local-only search (offline capable)", "params": "Same as search",
     "example": '{"query": "docker guide", "limit": 3}'},
    {"name": "fetch_page", "description": "Fetch full text of a URL", "params": "url (str), format (str)",
     "example": '{"url": "https://example.com"}'},
    {"name": "crawl_url", "description": "Crawl and index a URL", "params": "url (str), depth (int), force (bool), webhook_url (str)",
     "example": '{"url": "https://docs.python.org", "depth": 1}'},
    {"name": "network_stats", "description": "Network status and statistics", "params": "format (str)",
     "example": '{"format": "json"}'},
    {"name": "batch_search", "description": "Multiple searches in one call", "params": "queries (list[str]), limit (int), format (str)",
     "example": '{"queries": ["python", "rust"], "limit": 3}'},
    {"name": "suggest", "description": "Search autocomplete suggestions", "params": "prefix (str), limit (int)",
     "example": '{"prefix": "pyth", "limit": 5}'},
    {"name": "register_webhook", "description": "Register crawl completion webhook", "params": "url (str)",
     "example": '{"url": "https://example.com/webhook"}'},
    {"name": "analytics", "description": "Search and crawl analytics", "params": "format (str)", "example": '{"format": "json"}'},
]


def generate_tool_guide(*, format: str = "text") -> str:
    if format == "markdown":
        out = ["# InfoMesh MCP Tools Reference\n"]
        for t in MCP_TOOLS_GUIDE:
            out += [f"## `{t['name']}`\n", f"{t['description']}\n", f"**Parameters**: {t['params']}\n", f"**Example**: `{t['example']}`\n"]
        return "\n".join(out)
    out = ["InfoMesh MCP Tools Reference", "=" * 30, ""]
    for t in MCP_TOOLS_GUIDE:
        out += [f"  {t['name']}", f"    {t['description']}", f"    Params: {t['params']}", f"    Example: {t['example']}", ""]
    return "\n".join(out)


@dataclass
class ChangelogEntry:
    version: str
    date: str
    changes: list[str] = field(default_factory=list)
    breaking: list[str] = field(default_factory=list)

    def to_markdown(self) -> str:
        out = [f"## [{self.version}] - {self.date}\n"]
        if self.breaking:
            out += ["### Breaking Changes\n", *[f"- {c}" for c in self.breaking], ""]
        if self.changes:
            out += ["### Changes\n", *[f"- {c}" for c in self.changes]]
        return "\n".join(out)


def generate_changelog(entries: list[ChangelogEntry]) -> str:
    head = ("# Changelog\n\nAll notable changes to InfoMesh will be documented in this file.\n\n"
            "The format is based on [Keep a Changelog](https://keepachangelog.com/).\n")
    return f"{head}\n" + "\n\n".join(e.to_markdown() for e in entries)
