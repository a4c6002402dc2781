"""Optional dashboard background music / sound effects through an external player (mpv, ffplay, afplay, paplay):
detection, looped playback in its own process group, volume mapping, health check + restart, orphan cleanup by the
PIDs recorded in ``~/.infomesh/bgm/players.pid`` — never by process-name pattern
(reference infomesh/dashboard/bgm.py:49-537; no auto-installer here)."""
from __future__ import annotations

import contextlib
import os
import shutil
import signal
import subprocess
from pathlib import Path
from typing import Any

_BGM_CACHE_DIR = Path.home() / ".infomesh" / "bgm"
_PID_FILE = _BGM_CACHE_DIR / "players.pid"
_PLAYERS: tuple[tuple[str, list[str]], ...] = (
    ("mpv", ["--no-video", "--really-quiet", "--loop=inf"]), ("ffplay", ["-nodisp", "-autoexit", "-loglevel", "quiet", "-loop", "0"]),
    ("afplay", []), ("paplay", []))


def _find_player() -> tuple[str, list[str]] | None:
    for name, args in _PLAYERS:
        path = shutil.which(name)
        if path:
            return path, list(args)
    return None


def _build_volume_args(player_cmd: str, volume: int) -> list[str]:
    v = max(0, min(int(volume), 100))
    base = os.path.basename(player_cmd)
    if base == "mpv":
        return [f"--volume={v}"]
    if base == "ffplay":
        return ["-volume", str(v)]
    if base == "afplay":
        return ["-v", f"{v / 100:.2f}"]
    if base == "paplay":
        return [f"--volume={int(v / 100 * 65536)}"]
    return []


def ensure_bgm_assets() -> Path:
    _BGM_CACHE_DIR.mkdir(parents=True, exist_ok=True)
    return _BGM_CACHE_DIR


def _record_pid(pid: int) -> None:
    with contextlib.suppress(OSError):
        ensure_bgm_assets()
        with open(_PID_FILE, "a", encoding="utf-8") as f:
            f.write(f"{pid}\n")


def kill_orphaned_bgm() -> int:
    """Terminate players this installation started earlier (exact PIDs from the pid file, verified by /proc cmdline)."""
    try:
        pids = [int(x) for x in _PID_FILE.read_text().split()]
    except (OSError, ValueError):
        return 0
    killed = 0
    names = tuple(n for n, _ in _PLAYERS)
    for pid in pids:
        try:
            cmd = Path(f"/proc/{pid}/cmdline").read_bytes().split(b"\0")[0].decode("utf-8", "replace")
        except OSError:
            continue
        if os.path.basename(cmd) in names:
            with contextlib.suppress(ProcessLookupError, PermissionError):
                os.kill(pid, signal.SIGTERM)
                killed += 1
    with contextlib.suppress(OSError):
        _PID_FILE.unlink()
    return killed


def _popen_kwargs() -> dict[str, Any]:
    return {"stdin": subprocess.DEVNULL, "stdout": subprocess.DEVNULL, "stderr": subprocess.DEVNULL, "start_new_session": True}


class BGMPlayer:
    def __init__(self, *, auto_install_mpv: bool = False):
        self._player = _find_player()
        self._proc: subprocess.Popen | None = None
        self._sfx: list[subprocess.Popen] = []
        self._track: tuple[str, int] | None = None

    @property
    def available(self) -> bool:
        return self._player is not None

    @property
    def is_playing(self) -> bool:
        return self._proc is not None and self._proc.poll() is None

    def _spawn(self, path: str | Path, volume: int, loop: bool) -> subprocess.Popen | None:
        if self._player is None or not Path(path).exists():
            return None
        cmd, args = self._player
        if not loop:
            args = [a for a in args if "loop" not in a and a != "0"]
        try:
            proc = subprocess.Popen([cmd, *args, *_build_volume_args(cmd, volume), str(path)], **_popen_kwargs())
        except OSError:
            return None
        _record_pid(proc.pid)
        return proc

    def play(self, path: str | Path, *, loop: bool = True, volume: int = 50) -> bool:
        self.stop()
        self._proc = self._spawn(path, volume, loop=loop)
        self._track = (str(path), volume) if self._proc else None
        return self._proc is not None

    def play_sfx(self, path: str | Path, *, volume: int = 100) -> bool:
        self.reap_sfx()
        p = self._spawn(path, volume, loop=False)
        if p is not None:
            self._sfx.append(p)
        return p is not None

    def reap_sfx(self) -> None:
        self._sfx = [p for p in self._sfx if p.poll() is None]

    def check_and_restart(self) -> bool:
        """True when a player that should be running had died and was restarted."""
        if self._track is None or self.is_playing:
            return False
        self._proc = self._spawn(self._track[0], self._track[1], loop=True)
        return self._proc is not None

    def stop(self) -> None:
        for p in [self._proc, *self._sfx]:
            if p is not None and p.poll() is None:
                with contextlib.suppress(ProcessLookupError, PermissionError):
                    os.killpg(p.pid, signal.SIGTERM)
        self._proc, self._sfx, self._track = None, [], None

    def toggle(self, path: str | Path, *, volume: int | None = None) -> bool:
        if self.is_playing:
            self.stop()
            return False
        return self.play(path, volume=volume if volume is not None else 50)
