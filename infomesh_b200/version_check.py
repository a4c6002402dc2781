"""Update detection: the package index (24 h on-disk cache, 5 s timeout) and version gossip from PING/PONG exchanges
(reference infomesh/version_check.py:27-280).  Comparison keeps only the numeric prefix of each dotted segment."""
from __future__ import annotations

import contextlib
import json
import re
import time
from dataclasses import dataclass
from pathlib import Path

from infomesh_b200 import __version__

_PYPI_URL = "https://pypi.org/pypi/infomesh/json"
_CACHE_TTL_SECONDS = 86400
_CACHE_FILE_NAME = "version_cache.json"
_REQUEST_TIMEOUT = 5.0
_LEADING_INT = re.compile(r"\d+")


@dataclass(frozen=True)
class UpdateInfo:
    current: str
    latest: str
    source: str      # "pypi" | "peer"


def _parse_version(v: str) -> tuple[int, ...]:
    nums = []
    for seg in v.split("."):
        m = _LEADING_INT.match(seg)
        if m:
            nums.append(int(m.group()))
    return tuple(nums) or (0,)


def is_newer(candidate: str, current: str | None = None) -> bool:
    return _parse_version(candidate) > _parse_version(current or __version__)


def _read_cache(data_dir: Path) -> str | None:
    try:
        raw = json.loads((Path(data_dir) / _CACHE_FILE_NAME).read_text(encoding="utf-8"))
        if time.time() - float(raw.get("ts", 0)) > _CACHE_TTL_SECONDS:
            return None
        ver = raw.get("version", "")
        return ver if isinstance(ver, str) and ver else None
    except Exception:  # noqa: BLE001
        return None


def _write_cache(data_dir: Path, version: str) -> None:
    with contextlib.suppress(OSError):
        (Path(data_dir) / _CACHE_FILE_NAME).write_text(json.dumps({"version": version, "ts": time.time()}), encoding="utf-8")


def _fetch_latest_from_pypi() -> str | None:
    try:
        import urllib.request

        with urllib.request.urlopen(_PYPI_URL, timeout=_REQUEST_TIMEOUT) as resp:  # noqa: S310 — constant https URL
            if resp.status != 200:
                return None
            ver = json.loads(resp.read(1 << 22)).get("info", {}).get("version", "")
        return ver if isinstance(ver, str) and ver else None
    except Exception:  # noqa: BLE001
        return None


def check_pypi_update(data_dir: Path) -> UpdateInfo | None:
    latest = _read_cache(data_dir)
    if latest is None:
        latest = _fetch_latest_from_pypi()
        if latest is None:
            return None
        _write_cache(data_dir, latest)
    return UpdateInfo(__version__, latest, "pypi") if is_newer(latest) else None


class PeerVersionTracker:
    def __init__(self):
        self._peer_versions: dict[str, str] = {}

    def record(self, peer_id: str, version: str) -> None:
        if version and isinstance(version, str) and len(version) <= 32:
            self._peer_versions[peer_id] = version

    def get_newest_peer_version(self) -> str | None:
        return max(self._peer_versions.values(), key=_parse_version) if self._peer_versions else None

    def check_peer_update(self) -> UpdateInfo | None:
        newest = self.get_newest_peer_version()
        return UpdateInfo(__version__, newest, "peer") if newest and is_newer(newest) else None

    @property
    def peer_versions(self) -> dict[str, str]:
        return dict(self._peer_versions)


def check_for_update(data_dir: Path | None = None, peer_tracker: PeerVersionTracker | None = None) -> UpdateInfo | None:
    best: UpdateInfo | None = None
    if data_dir is not None:
        with contextlib.suppress(Exception):
            best = check_pypi_update(data_dir)
    if peer_tracker is not None:
        peer = peer_tracker.check_peer_update()
        if peer is not None and (best is None or is_newer(peer.latest, best.latest)):
            best = peer
    return best


def format_update_banner(info: UpdateInfo) -> str:
    src = "PyPI" if info.source == "pypi" else "P2P peer"
    return f"\n  ⬆ Update available ({src}): v{info.current} → v{info.latest}\n    Run: infomesh update\n"
