"""Graceful shutdown: SIGTERM / SIGINT run the registered callbacks once, then close the AppContext (async close when
a loop is running) (reference infomesh/shutdown.py:17-110)."""
from __future__ import annotations

import asyncio
import contextlib
import signal
import threading
from typing import Any

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class GracefulShutdown:
    def __init__(self):
        self._flag = threading.Event()
        self._context: Any | None = None
        self._callbacks: list[Any] = []
        self._task: asyncio.Future | None = None

    def register(self, context: Any) -> None:
        self._context = context
        loop = None
        with contextlib.suppress(RuntimeError):
            loop = asyncio.get_running_loop()
        for sig in (signal.SIGTERM, signal.SIGINT):
            try:
                if loop is not None:
                    loop.add_signal_handler(sig, self._handle_signal)
                else:
                    signal.signal(sig, self._sync_handler)
            except (ValueError, NotImplementedError, RuntimeError):    # not the main thread / unsupported platform
                logger.debug("signal_handler_not_installed", signal=int(sig))

    def add_callback(self, callback: Any) -> None:
        self._callbacks.append(callback)

    def _try_set_shutting_down(self) -> bool:
        if self._flag.is_set():
            return False
        self._flag.set()
        return True

    def _handle_signal(self) -> None:
        if self._try_set_shutting_down():
            self._task = asyncio.ensure_future(self.cleanup())

    def _sync_handler(self, signum: int, frame: Any) -> None:
        if not self._try_set_shutting_down():
            return
        for cb in self._callbacks:
            if callable(cb) and not asyncio.iscoroutinefunction(cb):
                with contextlib.suppress(Exception):
                    cb()
        if self._context is not None:
            with contextlib.suppress(Exception):
                self._context.close()
        raise SystemExit(0)

    async def cleanup(self) -> None:
        for cb in self._callbacks:
            try:
                if asyncio.iscoroutinefunction(cb):
                    await cb()
                elif callable(cb):
                    cb()
            except Exception:  # noqa: BLE001
                logger.exception("shutdown_callback_failed")
        ctx = self._context
        if ctx is not None:
            try:
                if hasattr(ctx, "close_async"):
                    await ctx.close_async()
                else:
                    ctx.close()
            except Exception:  # noqa: BLE001
                logger.exception("shutdown_context_close_failed")

    @property
    def is_shutting_down(self) -> bool:
        return self._flag.is_set()
