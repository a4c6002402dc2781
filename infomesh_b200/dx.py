"""Developer-facing conveniences: lifecycle plugins (objects with ``setup`` / ``teardown``), a swappable search tokenizer,
the MCP tool quick-reference and changelog rendering.

Contract (SURVEY §2.1 "dx"; reference infomesh/dx.py): a plugin is anything with a ``name`` and ``setup(app)`` /
``teardown()``; plugins may be loaded from a module exposing a ``plugin`` attribute; a failing plugin is logged, never
fatal; disabled plugins are skipped; the default tokenizer lower-cases, splits on non-word characters and drops
one-character tokens; the tool reference renders as plain text or markdown; the changelog follows Keep a Changelog.

Implementation: plugins and their descriptors live together in one slot record; both lifecycle sweeps are the same
guarded loop parameterised by the method name; the tool reference is a table of rows rendered through two line
templates."""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from importlib import import_module
from typing import Any, NamedTuple, Protocol

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


# ----------------------------------------------------------------------------- lifecycle plugins
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


class _Slot(NamedTuple):
    plugin: PluginProtocol
    info: PluginInfo


class PluginManager:
    def __init__(self):
        self._slots: dict[str, _Slot] = {}

    def register(self, plugin: PluginProtocol, *, info: PluginInfo | None = None) -> None:
        self._slots[plugin.name] = _Slot(plugin, info if info is not None else PluginInfo(name=plugin.name))

    def load_module(self, module_path: str) -> bool:
        """Import ``module_path`` and register its module-level ``plugin`` object."""
        try:
            module = import_module(module_path)
        except Exception as exc:  # noqa: BLE001
            logger.error("plugin_load_error", module=module_path, error=str(exc))
            return False
        candidate = getattr(module, "plugin", None)
        if candidate is None:
            logger.warning("plugin_missing_plugin_var", module=module_path)
            return False
        self.register(candidate, info=PluginInfo(name=candidate.name, module_path=module_path))
        return True

    def _sweep(self, method: str, *args: Any) -> None:
        for name, slot in self._slots.items():
            if slot.info.enabled:
                try:
                    getattr(slot.plugin, method)(*args)
                except Exception as exc:  # noqa: BLE001 -- one plugin must not take the node down
                    logger.error(f"plugin_{method}_error", name=name, error=str(exc))

    def setup_all(self, app: Any) -> None:
        self._sweep("setup", app)

    def teardown_all(self) -> None:
        self._sweep("teardown")

    def list_plugins(self) -> list[PluginInfo]:
        return [slot.info for slot in self._slots.values()]


# ----------------------------------------------------------------------------- search tokenizer
class TokenizerHook(Protocol):
    def tokenize(self, text: str) -> list[str]: ...


class DefaultTokenizer:
    _SPLIT = re.compile(r"\W+")

    def tokenize(self, text: str) -> list[str]:
        return [piece for piece in self._SPLIT.split(text.lower()) if len(piece) > 1]


_tokenizer_slot: list[TokenizerHook] = [DefaultTokenizer()]


def set_tokenizer(tokenizer: TokenizerHook) -> None:
    _tokenizer_slot[0] = tokenizer
    logger.info("custom_tokenizer_set", type=type(tokenizer).__name__)


def get_tokenizer() -> TokenizerHook:
    return _tokenizer_slot[0]


# ----------------------------------------------------------------------------- MCP tool quick-reference
class _ToolRow(NamedTuple):
    name: str
    description: str
    params: str
    example: str


_TOOL_ROWS = (
    _ToolRow("search", "Full network search",
             "query (str), limit (int), format (str), language (str), date_from (float), date_to (float), include_domains (list), "
             "exclude_domains (list), offset (int), snippet_length (int), session_id (str)",
             '{"query": "python asyncio", "limit": 5, "format": "json"}'),
    _ToolRow("search_local", "Local-only search (offline capable)", "Same as search", '{"query": "docker guide", "limit": 3}'),
    _ToolRow("fetch_page", "Fetch full text of a URL", "url (str), format (str)", '{"url": "https://example.com"}'),
    _ToolRow("crawl_url", "Crawl and index a URL", "url (str), depth (int), force (bool), webhook_url (str)",
             '{"url": "https://docs.python.org", "depth": 1}'),
    _ToolRow("network_stats", "Network status and statistics", "format (str)", '{"format": "json"}'),
    _ToolRow("batch_search", "Multiple searches in one call", "queries (list[str]), limit (int), format (str)",
             '{"queries": ["python", "rust"], "limit": 3}'),
    _ToolRow("suggest", "Search autocomplete suggestions", "prefix (str), limit (int)", '{"prefix": "pyth", "limit": 5}'),
    _ToolRow("register_webhook", "Register crawl completion webhook", "url (str)", '{"url": "https://example.com/webhook"}'),
    _ToolRow("analytics", "Search and crawl analytics", "format (str)", '{"format": "json"}'),
)
MCP_TOOLS_GUIDE: list[dict[str, str]] = [row._asdict() for row in _TOOL_ROWS]

_TITLE = "InfoMesh MCP Tools Reference"
_LAYOUT = {
    "markdown": (f"# {_TITLE}\n", "## `{name}`\n\n{description}\n\n**Parameters**: {params}\n\n**Example**: `{example}`\n"),
    "text": (f"{_TITLE}\n{'=' * 30}\n", "  {name}\n    {description}\n    Params: {params}\n    Example: {example}\n"),
}


def generate_tool_guide(*, format: str = "text") -> str:
    header, row_template = _LAYOUT["markdown" if format == "markdown" else "text"]
    return "\n".join([header, *(row_template.format(**row) for row in MCP_TOOLS_GUIDE)])


# ----------------------------------------------------------------------------- changelog
@dataclass
class ChangelogEntry:
    version: str
    date: str
    changes: list[str] = field(default_factory=list)
    breaking: list[str] = field(default_factory=list)

    def to_markdown(self) -> str:
        def section(title: str, items: list[str]) -> list[str]:
            return [f"### {title}\n", *(f"- {item}" for item in items)] if items else []

        blocks = [section("Breaking Changes", self.breaking), section("Changes", self.changes)]
        body = "\n\n".join("\n".join(b) for b in blocks if b)
        return f"## [{self.version}] - {self.date}\n\n{body}".rstrip()


_CHANGELOG_PREAMBLE = ("# Changelog\n\nAll notable changes to InfoMesh will be documented in this file.\n\n"
                       "The format is based on [Keep a Changelog](https://keepachangelog.com/).\n")


def generate_changelog(entries: list[ChangelogEntry]) -> str:
    return _CHANGELOG_PREAMBLE + "\n" + "\n\n".join(entry.to_markdown() for entry in entries)
