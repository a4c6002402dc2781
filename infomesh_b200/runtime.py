"""Process coordination for ``infomesh start/stop/status``: PID file with liveness + command-line check, an flock'ed
startup lock, and the JSON runtime-status snapshot (stale after 30 s) (reference infomesh/runtime.py:26-313)."""
from __future__ import annotations

import contextlib
import json
import os
import signal
import sys
import time
from pathlib import Path
from typing import Any

try:
    import fcntl
except ImportError:  # pragma: no cover — non-Unix
    fcntl = None  # type: ignore[assignment]

PID_FILE_NAME = "infomesh.pid"
STARTUP_LOCK_FILE_NAME = "infomesh.start.lock"
RUNTIME_STATUS_FILE_NAME = "runtime_status.json"
RUNTIME_STATUS_MAX_AGE_SECONDS = 30.0


def pid_path(data_dir: Path) -> Path:
    return Path(data_dir) / PID_FILE_NAME


def startup_lock_path(data_dir: Path) -> Path:
    return Path(data_dir) / STARTUP_LOCK_FILE_NAME


def runtime_status_path(data_dir: Path) -> Path:
    return Path(data_dir) / RUNTIME_STATUS_FILE_NAME


def _atomic_write(path: Path, text: str) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_name(path.name + ".tmp")
    tmp.write_text(text, encoding="utf-8")
    tmp.replace(path)


def is_process_running(pid: int) -> bool:
    if pid <= 0:
        return False
    try:
        os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    return True


def _process_cmdline(pid: int) -> str | None:
    if pid == os.getpid():
        return " ".join([sys.executable, *sys.argv])
    try:
        return (Path("/proc") / str(pid) / "cmdline").read_bytes().replace(b"\0", b" ").decode("utf-8", "replace").strip()
    except OSError:
        return None


def is_infomesh_process(pid: int) -> bool:
    """Alive and (when the command line is readable) an infomesh process — guards against PID reuse."""
    if not is_process_running(pid):
        return False
    if pid == os.getpid():
        return True
    cmd = _process_cmdline(pid)
    return cmd is None or "infomesh" in cmd


def read_live_pid(data_dir: Path) -> int | None:
    path = pid_path(data_dir)
    try:
        pid = int(path.read_text(encoding="utf-8").strip())
    except FileNotFoundError:
        return None
    except (OSError, ValueError):
        path.unlink(missing_ok=True)
        return None
    if is_infomesh_process(pid):
        return pid
    path.unlink(missing_ok=True)
    return None


def write_pid_file(data_dir: Path, pid: int) -> None:
    _atomic_write(pid_path(data_dir), str(pid))


def clear_pid_file(data_dir: Path, pid: int) -> None:
    path = pid_path(data_dir)
    try:
        owner = int(path.read_text(encoding="utf-8").strip())
    except (OSError, ValueError):
        path.unlink(missing_ok=True)
        return
    if owner == pid:
        path.unlink(missing_ok=True)


def wait_for_process_exit(pid: int, *, timeout_seconds: float = 5.0, poll_interval_seconds: float = 0.05) -> bool:
    deadline = time.monotonic() + timeout_seconds
    while time.monotonic() < deadline:
        if not is_process_running(pid):
            return True
        time.sleep(poll_interval_seconds)
    return not is_process_running(pid)


class StartupLock:
    """``with StartupLock(data_dir):`` serialises node start-up across processes (exclusive flock, 5 s wait)."""

    def __init__(self, data_dir: Path, *, timeout_seconds: float = 5.0, poll_interval_seconds: float = 0.05):
        self._dir, self._timeout, self._poll = Path(data_dir), timeout_seconds, poll_interval_seconds
        self._fh: Any | None = None
        self.acquired = False

    def acquire(self) -> bool:
        self._dir.mkdir(parents=True, exist_ok=True)
        self._fh = startup_lock_path(self._dir).open("a+", encoding="utf-8")
        if fcntl is None:
            self.acquired = True
            return True
        deadline = time.monotonic() + self._timeout
        while True:
            try:
                fcntl.flock(self._fh.fileno(), fcntl.LOCK_EX | fcntl.LOCK_NB)
            except BlockingIOError:
                if time.monotonic() >= deadline:
                    self.release()
                    return False
                time.sleep(self._poll)
                continue
            self._fh.seek(0)
            self._fh.truncate()
            self._fh.write(str(os.getpid()))
            self._fh.flush()
            self.acquired = True
            return True

    def release(self) -> None:
        if self._fh is None:
            return
        if fcntl is not None and self.acquired:
            with contextlib.suppress(OSError):
                fcntl.flock(self._fh.fileno(), fcntl.LOCK_UN)
        self._fh.close()
        self._fh, self.acquired = None, False

    def __enter__(self) -> "StartupLock":
        if not self.acquire():
            raise RuntimeError("another InfoMesh startup is already in progress")
        return self

    def __exit__(self, *exc: object) -> None:
        self.release()


def build_runtime_status(*, pid: int, role: str, started_at: float, no_crawl: bool, governor_state: Any,
                         gpu: dict[str, Any] | None = None) -> dict[str, Any]:
    now, g = time.time(), governor_state
    out = {
        "status": "running", "pid": pid, "role": role, "no_crawl": no_crawl, "started_at": round(started_at, 3),
        "updated_at": round(now, 3), "uptime_seconds": round(now - started_at, 1),
        "degrade_level": getattr(g.degrade_level, "name", "UNKNOWN"), "cpu_percent": round(float(g.cpu_percent), 1),
        "memory_percent": round(float(g.memory_percent), 1), "process_memory_mb": round(float(g.process_memory_mb), 1),
        "process_memory_limit_mb": int(getattr(g, "process_memory_limit_mb", 0)),
        "process_memory_ratio": round(float(getattr(g, "process_memory_ratio", 0.0)), 3),
        "throttle_factor": round(float(g.throttle_factor), 3), "checks_performed": int(g.checks_performed),
    }
    if gpu:
        out["gpu"] = gpu
    return out


def write_runtime_status(data_dir: Path, status: dict[str, Any]) -> None:
    _atomic_write(runtime_status_path(data_dir), json.dumps(status, sort_keys=True))


def read_runtime_status(data_dir: Path, *, max_age_seconds: float | None = RUNTIME_STATUS_MAX_AGE_SECONDS) -> dict[str, Any]:
    path = runtime_status_path(data_dir)
    try:
        data = json.loads(path.read_text(encoding="utf-8"))
    except FileNotFoundError:
        return {}
    except (OSError, json.JSONDecodeError):
        path.unlink(missing_ok=True)
        return {}
    if not isinstance(data, dict):
        path.unlink(missing_ok=True)
        return {}
    if max_age_seconds is None:
        return data
    try:
        age = time.time() - float(data.get("updated_at", 0.0))
    except (TypeError, ValueError):
        age = max_age_seconds + 1.0
    if age > max_age_seconds:
        return {"status": "stopped", "pid": data.get("pid"), "stale": True, "age_seconds": round(age, 1)}
    return data


def mark_runtime_stopped(data_dir: Path, pid: int) -> None:
    cur = read_runtime_status(data_dir, max_age_seconds=None)
    if cur and cur.get("pid") not in (pid, None):
        return
    write_runtime_status(data_dir, {"status": "stopped", "pid": pid, "updated_at": round(time.time(), 3)})


def request_graceful_stop(pid: int, *, timeout_seconds: float = 5.0) -> bool:
    os.kill(pid, signal.SIGTERM)
    return wait_for_process_exit(pid, timeout_seconds=timeout_seconds)


class contextlib_suppress_os_error(contextlib.suppress):
    """``with contextlib_suppress_os_error(): ...`` — kept for callers written against reference runtime.py:202."""

    def __init__(self) -> None:
        super().__init__(OSError)
