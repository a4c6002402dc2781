"""Small helpers shared by the dashboard views (reference infomesh/dashboard/utils.py:11-201)."""
from __future__ import annotations

import json
import time
from typing import Any

from infomesh_b200.config import Config
from infomesh_b200.runtime import read_live_pid, read_runtime_status


_UPTIME_UNITS = (("d", 86400), ("h", 3600))
_BYTE_UNITS = ("B", "KB", "MB", "GB", "TB")


def format_uptime(seconds: float) -> str:
    """``1d 2h 5m`` -- zero days / hours are left out, minutes always shown, an em dash for "not running" (<= 0)."""
    if seconds <= 0:
        return "—"
    left, shown = int(seconds), []
    for suffix, span in _UPTIME_UNITS:
        count, left = divmod(left, span)
        if count:
            shown.append(f"{count}{suffix}")
    shown.append(f"{left // 60}m")
    return " ".join(shown)


def format_bytes(n: int | float) -> str:
    """One decimal and a 1024-based unit (``0 B`` for zero, PB past the terabytes)."""
    if n == 0:
        return "0 B"
    value = float(n)
    for unit in _BYTE_UNITS:
        if abs(value) < 1024:
            return f"{value:.1f} {unit}"
        value /= 1024
    return f"{value:.1f} PB"


def get_peer_id(config: Config) -> str:
    try:
        from infomesh_b200.p2p.keys import KeyPair

        return KeyPair.load(config.node.data_dir / "keys").peer_id
    except Exception:  # noqa: BLE001
        return "(not generated)"


def is_node_running(config: Config) -> bool:
    return read_live_pid(config.node.data_dir) is not None


def is_node_running_with_uptime(config: Config) -> tuple[bool, float]:
    if not is_node_running(config):
        return False, 0.0
    return True, float(read_runtime_status(config.node.data_dir).get("uptime_seconds", 0.0) or 0.0)


def read_p2p_status(config: Config, *, max_age: float = 30.0) -> dict[str, object]:
    try:
        data = json.loads((config.node.data_dir / "p2p_status.json").read_text())
    except (OSError, ValueError):
        return {}
    if not isinstance(data, dict):
        return {}
    if time.time() - float(data.get("timestamp", 0) or 0) > max_age:
        data = {**data, "state": "stopped", "peers": 0, "peer_ids": []}
    return data


_TIER_STARS = {"TIER_1": 1, "TIER_2": 2, "TIER_3": 3}


def tier_label(tier: Any) -> str:
    """``⭐⭐ Tier 2`` for anything whose ``.name`` is a contribution tier, ``Unknown`` otherwise (no enum import needed)."""
    stars = _TIER_STARS.get(getattr(tier, "name", ""))
    return f"{'⭐' * stars} Tier {stars}" if stars else "Unknown"


def format_doc_line(url: str, title: str) -> str:
    """``url  (title)`` for the live log; titles are cut to 40 characters + ellipsis."""
    if not title:
        return url
    label = title if len(title) <= 40 else title[:40] + "…"
    return f"{url}  ({label})"


def push_new_docs_to_log(log: Any, docs: list[Any], seen: set[int]) -> int:
    """Write not-yet-seen recent documents (oldest first) to a log widget exposing ``write_line``."""
    fresh = [d for d in reversed(docs) if d.doc_id not in seen]
    for d in fresh:
        seen.add(d.doc_id)
        log.write_line(f"{time.strftime('%H:%M:%S', time.localtime(d.crawled_at))}  ✔ {format_doc_line(d.url, d.title)}")
    return len(fresh)
