"""Cross-process coordination behind ``infomesh start`` / ``stop`` / ``status``.

Contract (SURVEY §2.1 "runtime"; reference infomesh/runtime.py): three files in the data directory --

* ``infomesh.pid``: the running node's pid.  A pid counts as live only if the process exists AND (where its command line
  can be read) looks like infomesh, so a recycled pid is not mistaken for the node; unreadable or stale files are removed
  on sight; a process only clears its own pid file.
* ``infomesh.start.lock``: an exclusive advisory lock held while a node starts (waits up to 5 s, then fails).
* ``runtime_status.json``: a heartbeat snapshot (role, uptime, governor readings, optional GPU block); older than 30 s it
  reads as ``{"status": "stopped", "stale": true}``; corrupt files are discarded.

Writes are atomic (temp file + rename).  Implementation: one ``_StateFile`` value object does all file handling (atomic
write, tolerant read, discard); the public functions are thin, named operations on the three files; the status snapshot is
assembled from a field table rather than a hand-written dict."""
from __future__ import annotations

import contextlib
import json
import os
import signal
import sys
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable

try:
    import fcntl
except ImportError:  # pragma: no cover -- non-Unix
    fcntl = None  # type: ignore[assignment]

PID_FILE_NAME = "infomesh.pid"
STARTUP_LOCK_FILE_NAME = "infomesh.start.lock"
RUNTIME_STATUS_FILE_NAME = "runtime_status.json"
RUNTIME_STATUS_MAX_AGE_SECONDS = 30.0


# ----------------------------------------------------------------------------- file plumbing
@dataclass(frozen=True)
class _StateFile:
    path: Path

    def put(self, text: str) -> None:
        self.path.parent.mkdir(parents=True, exist_ok=True)
        scratch = self.path.with_name(f"{self.path.name}.tmp")
        scratch.write_text(text, encoding="utf-8")
        os.replace(scratch, self.path)

    def take(self, parse: Callable[[str], Any]) -> Any:
        """Parsed content; ``None`` if the file is absent; unreadable / unparsable files are deleted and read as ``None``."""
        try:
            raw = self.path.read_text(encoding="utf-8")
        except FileNotFoundError:
            return None
        except OSError:
            self.discard()
            return None
        try:
            return parse(raw)
        except ValueError:
            self.discard()
            return None

    def discard(self) -> None:
        self.path.unlink(missing_ok=True)


def pid_path(data_dir: Path) -> Path:
    return Path(data_dir, PID_FILE_NAME)


def startup_lock_path(data_dir: Path) -> Path:
    return Path(data_dir, STARTUP_LOCK_FILE_NAME)


def runtime_status_path(data_dir: Path) -> Path:
    return Path(data_dir, RUNTIME_STATUS_FILE_NAME)


def _pid_file(data_dir: Path) -> _StateFile:
    return _StateFile(pid_path(data_dir))


def _status_file(data_dir: Path) -> _StateFile:
    return _StateFile(runtime_status_path(data_dir))


# ----------------------------------------------------------------------------- processes
def is_process_running(pid: int) -> bool:
    if pid <= 0:
        return False
    try:
        os.kill(pid, 0)                      # signal 0: existence / permission probe only
    except ProcessLookupError:
        return False
    except PermissionError:                  # exists, owned by someone else
        pass
    return True


def _process_cmdline(pid: int) -> str | None:
    if pid == os.getpid():
        return " ".join((sys.executable, *sys.argv))
    try:
        raw = Path(f"/proc/{pid}/cmdline").read_bytes()
    except OSError:
        return None
    return raw.replace(b"\x00", b" ").decode("utf-8", "replace").strip()


def is_infomesh_process(pid: int) -> bool:
    if not is_process_running(pid):
        return False
    if pid == os.getpid():
        return True
    line = _process_cmdline(pid)
    return True if line is None else "infomesh" in line     # no /proc: existence is all we can check


def wait_for_process_exit(pid: int, *, timeout_seconds: float = 5.0, poll_interval_seconds: float = 0.05) -> bool:
    give_up = time.monotonic() + timeout_seconds
    while is_process_running(pid):
        if time.monotonic() >= give_up:
            return False
        time.sleep(poll_interval_seconds)
    return True


def request_graceful_stop(pid: int, *, timeout_seconds: float = 5.0) -> bool:
    os.kill(pid, signal.SIGTERM)
    return wait_for_process_exit(pid, timeout_seconds=timeout_seconds)


# ----------------------------------------------------------------------------- pid file
def _recorded_pid(data_dir: Path) -> int | None:
    return _pid_file(data_dir).take(lambda raw: int(raw.strip()))


def write_pid_file(data_dir: Path, pid: int) -> None:
    _pid_file(data_dir).put(str(pid))


def read_live_pid(data_dir: Path) -> int | None:
    pid = _recorded_pid(data_dir)
    if pid is None:
        return None
    if is_infomesh_process(pid):
        return pid
    _pid_file(data_dir).discard()
    return None


def clear_pid_file(data_dir: Path, pid: int) -> None:
    if _recorded_pid(data_dir) == pid:       # (a corrupt file was already discarded by the read)
        _pid_file(data_dir).discard()


# ----------------------------------------------------------------------------- startup lock
class StartupLock:
    """``with StartupLock(data_dir): ...`` -- at most one node start-up per data directory at a time."""

    def __init__(self, data_dir: Path, *, timeout_seconds: float = 5.0, poll_interval_seconds: float = 0.05):
        self._file = startup_lock_path(Path(data_dir))
        self._patience, self._tick = timeout_seconds, poll_interval_seconds
        self._handle: Any | None = None
        self.acquired = False

    def _try_lock(self) -> bool:
        if fcntl is None:
            return True
        try:
            fcntl.flock(self._handle.fileno(), fcntl.LOCK_EX | fcntl.LOCK_NB)
        except BlockingIOError:
            return False
        return True

    def acquire(self) -> bool:
        self._file.parent.mkdir(parents=True, exist_ok=True)
        self._handle = self._file.open("a+", encoding="utf-8")
        give_up = time.monotonic() + self._patience
        while not self._try_lock():
            if time.monotonic() >= give_up:
                self.release()
                return False
            time.sleep(self._tick)
        self.acquired = True
        if fcntl is not None:                # leave the holder's pid in the file for humans
            self._handle.seek(0)
            self._handle.truncate()
            self._handle.write(str(os.getpid()))
            self._handle.flush()
        return True

    def release(self) -> None:
        handle, self._handle = self._handle, None
        if handle is None:
            return
        if self.acquired and fcntl is not None:
            try:
                fcntl.flock(handle.fileno(), fcntl.LOCK_UN)
            except OSError:
                pass
        handle.close()
        self.acquired = False

    def __enter__(self) -> "StartupLock":
        if self.acquire():
            return self
        raise RuntimeError("another InfoMesh startup is already in progress")

    def __exit__(self, *exc: object) -> None:
        self.release()


# ----------------------------------------------------------------------------- status heartbeat
# (key, governor attribute, conversion); attributes missing on older governor objects fall back to the default
_GOVERNOR_FIELDS: tuple[tuple[str, str, Callable[[Any], Any], Any], ...] = (
    ("cpu_percent", "cpu_percent", lambda v: round(float(v), 1), 0.0),
    ("memory_percent", "memory_percent", lambda v: round(float(v), 1), 0.0),
    ("process_memory_mb", "process_memory_mb", lambda v: round(float(v), 1), 0.0),
    ("process_memory_limit_mb", "process_memory_limit_mb", int, 0),
    ("process_memory_ratio", "process_memory_ratio", lambda v: round(float(v), 3), 0.0),
    ("throttle_factor", "throttle_factor", lambda v: round(float(v), 3), 1.0),
    ("checks_performed", "checks_performed", int, 0),
)


def build_runtime_status(*, pid: int, role: str, started_at: float, no_crawl: bool, governor_state: Any,
                         gpu: dict[str, Any] | None = None) -> dict[str, Any]:
    stamp = time.time()
    snapshot: dict[str, Any] = {"status": "running", "pid": pid, "role": role, "no_crawl": no_crawl,
                                "started_at": round(started_at, 3), "updated_at": round(stamp, 3),
                                "uptime_seconds": round(stamp - started_at, 1),
                                "degrade_level": getattr(governor_state.degrade_level, "name", "UNKNOWN")}
    for key, attr, convert, default in _GOVERNOR_FIELDS:
        snapshot[key] = convert(getattr(governor_state, attr, default))
    if gpu:
        snapshot["gpu"] = gpu
    return snapshot


def write_runtime_status(data_dir: Path, status: dict[str, Any]) -> None:
    _status_file(data_dir).put(json.dumps(status, sort_keys=True))


def _parse_status(raw: str) -> dict[str, Any]:
    doc = json.loads(raw)                    # JSONDecodeError is a ValueError
    if not isinstance(doc, dict):
        raise ValueError("status snapshot must be an object")
    return doc


def read_runtime_status(data_dir: Path, *, max_age_seconds: float | None = RUNTIME_STATUS_MAX_AGE_SECONDS) -> dict[str, Any]:
    doc = _status_file(data_dir).take(_parse_status)
    if doc is None:
        return {}
    if max_age_seconds is None:
        return doc
    try:
        age = time.time() - float(doc.get("updated_at", 0.0))
    except (TypeError, ValueError):
        age = float("inf")
    if age <= max_age_seconds:
        return doc
    return {"status": "stopped", "pid": doc.get("pid"), "stale": True, "age_seconds": round(min(age, 1e12), 1)}


def mark_runtime_stopped(data_dir: Path, pid: int) -> None:
    """Overwrite the heartbeat with a 'stopped' marker -- unless another process has taken the file over."""
    owner = read_runtime_status(data_dir, max_age_seconds=None).get("pid")
    if owner in (None, pid):
        write_runtime_status(data_dir, {"status": "stopped", "pid": pid, "updated_at": round(time.time(), 3)})


class contextlib_suppress_os_error(contextlib.suppress):
    """``with contextlib_suppress_os_error(): ...`` swallows ``OSError`` (name kept for API parity with the reference)."""

    def __init__(self) -> None:
        super().__init__(OSError)
