"""Pipeline extension points: callables bound to one of ten stages rewrite -- or, by returning ``None``, drop -- the item
passing through that stage.

Contract (SURVEY §2.1 "plugins"; reference infomesh/plugins.py): stages ``pre/post_{crawl,index,search,rank}`` plus
``custom_tokenizer`` / ``custom_scorer``; handlers run in registration order; a handler that raises is logged and skipped;
a ``None`` result ends the chain with ``None``; a named plugin bundles handlers and can be removed as a unit; the async
runner awaits coroutine handlers and calls plain ones.

Implementation: one ordered list of bindings ``(stage, handler, owner)`` instead of a per-stage table, and one driver
(a generator that yields each handler and receives its outcome) shared by the sync and async runners, so the chain
semantics exist once."""
from __future__ import annotations

import inspect
from collections import Counter
from dataclasses import dataclass
from enum import StrEnum
from typing import Any, Callable, Generator

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


@dataclass(frozen=True)
class _Binding:
    stage: HookPoint
    handler: Callable[[Any], Any]
    owner: str | None          # plugin name, or None for a bare @hook registration


class _Failed:
    """Marker sent into the chain driver when a handler raised."""


class PluginRegistry:
    def __init__(self):
        self._bindings: list[_Binding] = []
        self._versions: dict[str, str] = {}       # plugin name -> version, registration order

    # ---- registration
    def hook(self, point: HookPoint):
        """Decorator form: ``@registry.hook(HookPoint.PRE_INDEX)``."""
        def bind(fn):
            self._bindings.append(_Binding(point, fn, None))
            return fn

        return bind

    def register_plugin(self, name: str, version: str = "0.0.1", hooks: dict[HookPoint, Callable] | None = None) -> None:
        self._versions[name] = version
        self._bindings.extend(_Binding(stage, fn, name) for stage, fn in (hooks or {}).items())

    def unregister_plugin(self, name: str) -> bool:
        if self._versions.pop(name, None) is None:
            return False
        self._bindings = [b for b in self._bindings if b.owner != name]
        return True

    # ---- execution
    def _chain(self, point: HookPoint, item: Any) -> Generator[tuple[Callable, Any], Any, Any]:
        """Drive ``item`` through the handlers of ``point``: yields ``(handler, current item)``, is sent the handler's
        result (or :class:`_Failed`), returns the final item."""
        for binding in [b for b in self._bindings if b.stage == point]:
            outcome = yield binding.handler, item
            if outcome is _Failed:
                logger.warning("plugin_hook_error", point=point.value, fn=getattr(binding.handler, "__name__", repr(binding.handler)))
                continue
            if outcome is None:
                return None
            item = outcome
        return item

    def run_hook(self, point: HookPoint, data: Any) -> Any:
        chain = self._chain(point, data)
        try:
            handler, item = next(chain)
            while True:
                try:
                    result = handler(item)
                except Exception:  # noqa: BLE001 -- a broken plugin must not break the pipeline
                    result = _Failed
                handler, item = chain.send(result)
        except StopIteration as done:
            return done.value

    async def run_hook_async(self, point: HookPoint, data: Any) -> Any:
        chain = self._chain(point, data)
        try:
            handler, item = next(chain)
            while True:
                try:
                    result = handler(item)
                    if inspect.isawaitable(result):
                        result = await result
                except Exception:  # noqa: BLE001
                    result = _Failed
                handler, item = chain.send(result)
        except StopIteration as done:
            return done.value

    # ---- introspection
    @property
    def registered_plugins(self) -> list[dict[str, Any]]:
        return [{"name": name, "version": version} for name, version in self._versions.items()]

    @property
    def hook_counts(self) -> dict[str, int]:
        return dict(Counter(b.stage.value for b in self._bindings))


_shared: list[PluginRegistry] = []


def get_registry() -> PluginRegistry:
    """Process-wide registry (created on first use)."""
    if not _shared:
        _shared.append(PluginRegistry())
    return _shared[0]
