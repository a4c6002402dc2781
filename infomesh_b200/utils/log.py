"""structlog-style logger shim on top of stdlib logging.

The reference logs through structlog with an event name + keyword arguments
(reference infomesh/cli/__init__.py:22-31).  structlog is not installed here, so ``get_logger()`` returns an
object with the same call style (``log.info("event_name", key=value)``) that renders ``event key=value ...``.
If structlog is importable it is used directly.
"""
from __future__ import annotations

import logging
import sys
from typing import Any

try:  # pragma: no cover - optional dependency
    import structlog as _structlog
except Exception:  # noqa: BLE001
    _structlog = None


class _KVLogger:
    def __init__(self, name: str, bound: dict[str, Any] | None = None):
        self._log = logging.getLogger(name)
        self._bound = bound or {}

    def bind(self, **kw: Any) -> "_KVLogger":
        return _KVLogger(self._log.name, {**self._bound, **kw})

    def _emit(self, level: int, event: str, kw: dict[str, Any]) -> None:
        if not self._log.isEnabledFor(level):
            return
        exc_info = kw.pop("exc_info", None)
        fields = {**self._bound, **kw}
        tail = " ".join(f"{k}={v!r}" if isinstance(v, str) and " " in v else f"{k}={v}" for k, v in fields.items())
        self._log.log(level, f"{event} {tail}".rstrip(), exc_info=exc_info)

    def debug(self, event: str, /, **kw: Any) -> None:
        self._emit(logging.DEBUG, event, kw)

    def info(self, event: str, /, **kw: Any) -> None:
        self._emit(logging.INFO, event, kw)

    def warning(self, event: str, /, **kw: Any) -> None:
        self._emit(logging.WARNING, event, kw)

    warn = warning

    def error(self, event: str, /, **kw: Any) -> None:
        self._emit(logging.ERROR, event, kw)

    def critical(self, event: str, /, **kw: Any) -> None:
        self._emit(logging.CRITICAL, event, kw)

    def exception(self, event: str, /, **kw: Any) -> None:
        kw.setdefault("exc_info", True)
        self._emit(logging.ERROR, event, kw)


def get_logger(name: str = "infomesh") -> Any:
    if _structlog is not None:
        return _structlog.get_logger(name)
    return _KVLogger(name)


def configure(level: str = "info", stream=None, logfile: str | None = None, max_bytes: int = 10 * 1024 * 1024,
              backups: int = 5) -> None:
    """Console (or rotating-file) logging like the reference's CLI / ``_serve`` worker set-up."""
    root = logging.getLogger()
    root.setLevel(getattr(logging, level.upper(), logging.INFO))
    for h in list(root.handlers):
        root.removeHandler(h)
    if logfile:
        from logging.handlers import RotatingFileHandler

        handler: logging.Handler = RotatingFileHandler(logfile, maxBytes=max_bytes, backupCount=backups)
    else:
        handler = logging.StreamHandler(stream or sys.stderr)
    handler.setFormatter(logging.Formatter("%(asctime)s [%(levelname)-7s] %(message)s", "%Y-%m-%dT%H:%M:%S"))
    root.addHandler(handler)
