"""Hook-based plugin registry: functions registered at ten pipeline points transform (or, by returning None, drop)
the item flowing through (reference infomesh/plugins.py:34-152)."""
from __future__ import annotations

import asyncio
from collections import defaultdict
from enum import StrEnum
from typing import Any, Callable

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class HookPoint(StrEnum):
    PRE_CRAWL = "pre_crawl"
    POST_CRAWL = "post_crawl"
    PRE_INDEX = "pre_index"
    POST_INDEX = "post_index"
    PRE_SEARCH = "pre_search"
    POST_SEARCH = "post_search"
    PRE_RANK = "pre_rank"
    POST_RANK = "post_rank"
    CUSTOM_TOKENIZER = "custom_tokenizer"
    CUSTOM_SCORER = "custom_scorer"


class PluginRegistry:
    def __init__(self):
        self._hooks: dict[HookPoint, list[Callable]] = defaultdict(list)
        self._plugins: dict[str, dict[str, Any]] = {}

    def hook(self, point: HookPoint):
        def deco(fn):
            self._hooks[point].append(fn)
            return fn

        return deco

    def register_plugin(self, name: str, version: str = "0.0.1", hooks: dict[HookPoint, Callable] | None = None) -> None:
        self._plugins[name] = {"version": version, "hooks": hooks or {}}
        for point, fn in (hooks or {}).items():
            self._hooks[point].append(fn)

    def unregister_plugin(self, name: str) -> bool:
        plug = self._plugins.pop(name, None)
        if plug is None:
            return False
        for point, fn in plug["hooks"].items():
            with_fn = self._hooks.get(point, [])
            if fn in with_fn:
                with_fn.remove(fn)
        return True

    def run_hook(self, point: HookPoint, data: Any) -> Any:
        """A failing hook is logged and skipped; a hook returning None filters the item out."""
        for fn in self._hooks.get(point, ()):
            try:
                out = fn(data)
            except Exception:  # noqa: BLE001
                logger.warning("plugin_hook_error", point=point.value, fn=getattr(fn, "__name__", str(fn)))
                continue
            if out is None:
                return None
            data = out
        return data

    async def run_hook_async(self, point: HookPoint, data: Any) -> Any:
        for fn in self._hooks.get(point, ()):
            try:
                out = await fn(data) if asyncio.iscoroutinefunction(fn) else fn(data)
            except Exception:  # noqa: BLE001
                logger.warning("plugin_hook_error", point=point.value, fn=getattr(fn, "__name__", str(fn)))
                continue
            if out is None:
                return None
            data = out
        return data

    @property
    def registered_plugins(self) -> list[dict[str, Any]]:
        return [{"name": n, "version": p["version"]} for n, p in self._plugins.items()]

    @property
    def hook_counts(self) -> dict[str, int]:
        return {p.value: len(f) for p, f in self._hooks.items() if f}


_registry: PluginRegistry | None = None


def get_registry() -> PluginRegistry:
    global _registry
    if _registry is None:
        _registry = PluginRegistry()
    return _registry
