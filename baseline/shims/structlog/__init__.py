"""Stand-in for the `structlog` wheel (absent from this image) so the UNMODIFIED reference can be imported.

Only what the reference touches: ``get_logger()`` returning an object with the usual level methods that accept an
event name plus keyword fields, ``configure()`` and the processor / factory names its CLI passes to it.  Events go to
the stdlib ``logging`` tree under the "infomesh" logger (WARNING and above by default), which is what a quiet
structlog configuration would do.  Nothing here is benchmark code."""
from __future__ import annotations

import logging as _logging
import types as _types

_root = _logging.getLogger("infomesh")


class BoundLogger:
    __slots__ = ("_fields",)

    def __init__(self, **fields):
        self._fields = fields

    def bind(self, **kw):
        return BoundLogger(**{**self._fields, **kw})

    new = bind

    def unbind(self, *keys):
        return BoundLogger(**{k: v for k, v in self._fields.items() if k not in keys})

    def _emit(self, level, event, kw):
        if _root.isEnabledFor(level):
            kw.pop("exc_info", None)
            fields = {**self._fields, **kw}
            _root.log(level, "%s %s", event, " ".join(f"{k}={v!r}" for k, v in fields.items()))

    def debug(self, event=None, *a, **kw):
        self._emit(_logging.DEBUG, event, kw)

    def info(self, event=None, *a, **kw):
        self._emit(_logging.INFO, event, kw)

    def warning(self, event=None, *a, **kw):
        self._emit(_logging.WARNING, event, kw)

    warn = warning

    def error(self, event=None, *a, **kw):
        self._emit(_logging.ERROR, event, kw)

    def critical(self, event=None, *a, **kw):
        self._emit(_logging.CRITICAL, event, kw)

    def exception(self, event=None, *a, **kw):
        self._emit(_logging.ERROR, event, kw)


def get_logger(*_a, **kw):
    return BoundLogger(**kw)


getLogger = get_logger


def configure(*_a, **_k):
    return None


def is_configured():
    return True


def _ns(name, **members):
    m = _types.ModuleType(f"structlog.{name}")
    for k, v in members.items():
        setattr(m, k, v)
    return m


def _proc(*_a, **_k):
    return lambda _logger, _name, event_dict: event_dict


stdlib = _ns("stdlib", BoundLogger=BoundLogger, LoggerFactory=lambda *a, **k: None, add_log_level=_proc(),
             add_logger_name=_proc(), filter_by_level=_proc(), ProcessorFormatter=_proc)
processors = _ns("processors", TimeStamper=_proc, JSONRenderer=_proc, StackInfoRenderer=_proc, format_exc_info=_proc(),
                 UnicodeDecoder=_proc, add_log_level=_proc())
dev = _ns("dev", ConsoleRenderer=_proc)
contextvars = _ns("contextvars", bind_contextvars=lambda **k: None, clear_contextvars=lambda: None,
                  merge_contextvars=_proc())


def PrintLoggerFactory(*_a, **_k):     # noqa: N802 -- structlog API name; the CLI entry point configures it
    return None


def make_filtering_bound_logger(*_a, **_k):
    return BoundLogger
