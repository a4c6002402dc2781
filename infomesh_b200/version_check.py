"""Is a newer release out?  Two sources: the package index (asked at most once a day, answer memoised on disk, 5 s
network budget) and the versions peers announce in PING/PONG.

Contract (SURVEY §2.1 "version check"; reference infomesh/version_check.py): versions compare by the leading integer of each
dotted segment (``1.2.3rc1`` -> ``(1, 2, 3)``, nothing numeric -> ``(0,)``); every failure (offline, bad JSON, unwritable
cache) degrades to "no update known", never to an exception; peer-reported versions longer than 32 characters are ignored;
when both sources know an update the higher version wins.

Implementation: versions are turned into sort keys by one function used everywhere; the on-disk memo is a tiny record
class with ``load`` / ``store``; ``check_for_update`` gathers candidate announcements and takes the maximum."""
from __future__ import annotations

import json
import time
from dataclasses import dataclass
from itertools import takewhile
from pathlib import Path

from infomesh_b200 import DISTRIBUTION, __version__

_PYPI_URL = f"https://pypi.org/pypi/{DISTRIBUTION}/json"      # releases of THIS distribution, never the reference's
_CACHE_TTL_SECONDS = 86400
_CACHE_FILE_NAME = "version_cache.json"
_REQUEST_TIMEOUT = 5.0
_MAX_PEER_VERSION_LEN = 32
_SOURCE_LABEL = {"pypi": "PyPI", "peer": "P2P peer"}


@dataclass(frozen=True)
class UpdateInfo:
    current: str
    latest: str
    source: str      # "pypi" | "peer"


# ----------------------------------------------------------------------------- ordering
def _parse_version(text: str) -> tuple[int, ...]:
    """Sort key: leading digits of every dotted segment; segments without any are skipped."""
    key = tuple(int(digits) for part in text.split(".") if (digits := "".join(takewhile(str.isdigit, part))))
    return key or (0,)


def is_newer(candidate: str, current: str | None = None) -> bool:
    return _parse_version(candidate) > _parse_version(__version__ if not current else current)


def _clean(value: object) -> str | None:
    return value if isinstance(value, str) and value else None


# ----------------------------------------------------------------------------- package index, memoised on disk
class _Memo:
    """``{"version": ..., "ts": ...}`` next to the node's other state files."""

    def __init__(self, data_dir: Path):
        self.file = Path(data_dir) / _CACHE_FILE_NAME

    def load(self) -> str | None:
        try:
            rec = json.loads(self.file.read_text(encoding="utf-8"))
            fresh = time.time() - float(rec.get("ts", 0)) <= _CACHE_TTL_SECONDS
            return _clean(rec.get("version", "")) if fresh else None
        except Exception:  # noqa: BLE001 -- missing / corrupt memo == no memo
            return None

    def store(self, version: str) -> None:
        try:
            self.file.write_text(json.dumps({"version": version, "ts": time.time()}), encoding="utf-8")
        except OSError:
            pass


def _read_cache(data_dir: Path) -> str | None:
    return _Memo(data_dir).load()


def _write_cache(data_dir: Path, version: str) -> None:
    _Memo(data_dir).store(version)


def _fetch_latest_from_pypi() -> str | None:
    try:
        from urllib.request import urlopen

        with urlopen(_PYPI_URL, timeout=_REQUEST_TIMEOUT) as reply:  # noqa: S310 -- constant https URL
            if reply.status != 200:
                return None
            body = json.loads(reply.read(4 << 20))
        return _clean(body.get("info", {}).get("version", ""))
    except Exception:  # noqa: BLE001
        return None


def check_pypi_update(data_dir: Path) -> UpdateInfo | None:
    published = _read_cache(data_dir)
    if published is None:
        published = _fetch_latest_from_pypi()
        if published is None:
            return None
        _write_cache(data_dir, published)
    return UpdateInfo(__version__, published, "pypi") if is_newer(published) else None


# ----------------------------------------------------------------------------- peer gossip
class PeerVersionTracker:
    def __init__(self):
        self._seen: dict[str, str] = {}

    def record(self, peer_id: str, version: str) -> None:
        if _clean(version) and len(version) <= _MAX_PEER_VERSION_LEN:
            self._seen[peer_id] = version

    @property
    def peer_versions(self) -> dict[str, str]:
        return dict(self._seen)

    def get_newest_peer_version(self) -> str | None:
        return max(self._seen.values(), key=_parse_version, default=None)

    def check_peer_update(self) -> UpdateInfo | None:
        top = self.get_newest_peer_version()
        return UpdateInfo(__version__, top, "peer") if top is not None and is_newer(top) else None


# ----------------------------------------------------------------------------- combined
def check_for_update(data_dir: Path | None = None, peer_tracker: PeerVersionTracker | None = None) -> UpdateInfo | None:
    found: list[UpdateInfo] = []
    if data_dir is not None:
        try:
            found.append(check_pypi_update(data_dir))
        except Exception:  # noqa: BLE001
            pass
    if peer_tracker is not None:
        found.append(peer_tracker.check_peer_update())
    found = [f for f in found if f is not None]
    # max() keeps the first of equal keys: the package index wins a tie, as it is the authoritative source
    return max(found, key=lambda f: _parse_version(f.latest), default=None)


def format_update_banner(info: UpdateInfo) -> str:
    where = _SOURCE_LABEL.get(info.source, _SOURCE_LABEL["peer"])
    return f"\n  ⬆ Update available ({where}): v{info.current} → v{info.latest}\n    Run: infomesh update\n"
