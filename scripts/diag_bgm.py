#!/usr/bin/env python3
"""Why did the dashboard music stop?  Drive ``dashboard.bgm.BGMPlayer`` outside the TUI and log what its player process does.

    python scripts/diag_bgm.py                       # play a generated tone for 5 minutes, report every state change
    python scripts/diag_bgm.py --duration 1800       # ... for half an hour
    python scripts/diag_bgm.py --kill-after 20       # kill the player after 20 s: does check_and_restart() bring it back?
    python scripts/diag_bgm.py --check-orphans       # only look for players an earlier run left behind (exact PIDs)
    python scripts/diag_bgm.py --track ~/music.mp3 --volume 30

(The reference ships a script of the same name for its ffplay / mpv wrapper: scripts/diag_bgm.py; this one exercises this
package's player, which records the PIDs it starts and never matches processes by name.)"""
from __future__ import annotations

import argparse
import math
import os
import signal
import struct
import sys
import time
import wave
from datetime import datetime
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from infomesh_b200.dashboard import bgm  # noqa: E402


def say(msg: str) -> None:
    print(f"[{datetime.now().strftime('%H:%M:%S')}] {msg}", flush=True)


def tone(path: Path, seconds: float = 2.0, hz: float = 440.0, rate: int = 22050) -> Path:
    """A short sine wave, so the script needs no audio asset."""
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(b"".join(struct.pack("<h", int(8000 * math.sin(2 * math.pi * hz * i / rate))) for i in range(int(seconds * rate))))
    return path


def proc_state(pid: int) -> str:
    try:
        fields = Path(f"/proc/{pid}/stat").read_text().rsplit(")", 1)[1].split()
        return {"R": "running", "S": "sleeping", "D": "disk wait", "Z": "zombie", "T": "stopped"}.get(fields[0], fields[0])
    except OSError:
        return "gone"


def orphans() -> list[tuple[int, str]]:
    try:
        pids = [int(x) for x in bgm._PID_FILE.read_text().split()]
    except (OSError, ValueError):
        return []
    return [(pid, proc_state(pid)) for pid in pids if proc_state(pid) != "gone"]


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--duration", type=float, default=300.0, help="seconds to monitor")
    ap.add_argument("--volume", type=int, default=50)
    ap.add_argument("--track", type=Path, default=None, help="audio file (default: a generated tone)")
    ap.add_argument("--kill-after", type=float, default=0.0, help="SIGKILL the player after this many seconds to test the restart path")
    ap.add_argument("--check-orphans", action="store_true")
    a = ap.parse_args()

    found = bgm._find_player()
    say(f"player: {found[0] if found else 'none of ' + ', '.join(n for n, _ in bgm._PLAYERS) + ' on PATH'}")
    left = orphans()
    say(f"recorded players still alive: {left or 'none'}")
    if a.check_orphans:
        if left:
            say(f"terminated {bgm.kill_orphaned_bgm()} of them")
        return 0
    if not found:
        say("nothing to play with; install mpv or ffmpeg (ffplay)")
        return 1
    track = a.track.expanduser() if a.track else tone(bgm.ensure_bgm_assets() / "diag_tone.wav")
    player = bgm.BGMPlayer()
    if not player.play(track, volume=a.volume):
        say(f"could not start the player on {track}")
        return 1
    pid = player._proc.pid
    say(f"playing {track} at volume {a.volume} (pid {pid}, argv volume flags {bgm._build_volume_args(found[0], a.volume)})")
    t0, last, restarts, killed = time.monotonic(), None, 0, False
    try:
        while time.monotonic() - t0 < a.duration:
            state = proc_state(player._proc.pid) if player._proc else "gone"
            if state != last:
                say(f"pid {player._proc.pid if player._proc else pid}: {last or 'start'} -> {state}  (+{time.monotonic() - t0:.1f}s)")
                last = state
            if a.kill_after and not killed and time.monotonic() - t0 >= a.kill_after:
                say(f"killing pid {player._proc.pid} to test the restart path")
                os.kill(player._proc.pid, signal.SIGKILL)
                killed = True
            if player.check_and_restart():
                restarts += 1
                say(f"player had exited (return code seen by poll) -> restarted as pid {player._proc.pid}")
                last = None
            time.sleep(0.5)
    except KeyboardInterrupt:
        say("interrupted")
    finally:
        player.stop()
        time.sleep(0.3)
        say(f"stopped; restarts: {restarts}; still alive afterwards: {orphans() or 'none'}")
        bgm.kill_orphaned_bgm()
    return 0


if __name__ == "__main__":
    sys.exit(main())
