"""Orderly exit on SIGTERM / SIGINT.

Contract (SURVEY §2.1 "shutdown"; reference infomesh/shutdown.py): the first signal -- and only the first -- runs the
registered callbacks in order and then closes the application context (``close_async`` when it has one and a loop is
running, else ``close``); a callback that raises is logged and does not stop the rest; inside an event loop the work runs
as a task, outside one it runs synchronously (coroutine callbacks cannot, and are skipped) and ends the process with exit
status 0; installing handlers from a non-main thread is tolerated.

Implementation: a one-shot latch guards the sequence; the signal wiring picks the loop or ``signal.signal`` per signal; the
synchronous and asynchronous paths share one ordered list of steps and differ only in how a step is invoked."""
from __future__ import annotations

import asyncio
import inspect
import signal
import threading
from typing import Any, Callable

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_SIGNALS = (signal.SIGTERM, signal.SIGINT)


def _running_loop() -> asyncio.AbstractEventLoop | None:
    try:
        return asyncio.get_running_loop()
    except RuntimeError:
        return None


class GracefulShutdown:
    def __init__(self):
        self._latch = threading.Lock()
        self._started = False
        self._context: Any | None = None
        self._callbacks: list[Callable[[], Any]] = []
        self._task: asyncio.Future | None = None

    # ---- wiring
    def register(self, context: Any) -> None:
        """Remember the context to close and hook both termination signals."""
        self._context = context
        loop = _running_loop()
        for sig in _SIGNALS:
            try:
                if loop is None:
                    signal.signal(sig, self._on_signal_sync)
                else:
                    loop.add_signal_handler(sig, self._on_signal_in_loop)
            except (ValueError, NotImplementedError, RuntimeError):    # wrong thread, or a platform without the API
                logger.debug("signal_handler_not_installed", signal=int(sig))

    def add_callback(self, callback: Callable[[], Any]) -> None:
        self._callbacks.append(callback)

    # ---- one-shot latch
    def _try_set_shutting_down(self) -> bool:
        with self._latch:
            first, self._started = not self._started, True
        return first

    @property
    def is_shutting_down(self) -> bool:
        return self._started

    # ---- the sequence
    def _closer(self) -> Callable[[], Any] | None:
        ctx = self._context
        if ctx is None:
            return None
        return getattr(ctx, "close_async", None) or ctx.close

    async def cleanup(self) -> None:
        steps = [(cb, "shutdown_callback_failed") for cb in self._callbacks if callable(cb)]
        closer = self._closer()
        if closer is not None:
            steps.append((closer, "shutdown_context_close_failed"))
        for step, event in steps:
            try:
                outcome = step()
                if inspect.isawaitable(outcome):
                    await outcome
            except Exception:  # noqa: BLE001
                logger.exception(event)

    def _on_signal_in_loop(self) -> None:
        if self._try_set_shutting_down():
            self._task = asyncio.ensure_future(self.cleanup())

    def _on_signal_sync(self, signum: int, frame: Any) -> None:
        if not self._try_set_shutting_down():
            return
        plain = [cb for cb in self._callbacks if callable(cb) and not inspect.iscoroutinefunction(cb)]
        if self._context is not None:
            plain.append(self._context.close)
        for step in plain:
            try:
                step()
            except Exception:  # noqa: BLE001 -- keep going: the process is on its way out
                pass
        raise SystemExit(0)
