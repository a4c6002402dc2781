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


def push_new_docs_to_log(recent_docs: "list[Any]", doc_count: int, seen_ids: set[int], last_count: int, log_widget: Any) -> tuple[set[int], int]:
    """Feed documents that were not shown yet (oldest first) to a live log and return the updated ``(seen_ids, last_count)``.

    Nothing new and no growth -> unchanged; the seen set is mutated in place and trimmed to its newest 300 ids once it passes
    500, so a pane that runs for weeks does not accumulate every document id it ever displayed."""
    if doc_count <= last_count and not recent_docs:
        return seen_ids, last_count
    fresh = sorted((d for d in recent_docs if d.doc_id not in seen_ids), key=lambda d: d.crawled_at)
    if not fresh:
        return seen_ids, doc_count
    for doc in fresh:
        try:
            log_widget.log_crawl(format_doc_line(doc.url, doc.title), success=True)
        except Exception:  # noqa: BLE001 -- a log widget that is being torn down must not stop the pane's refresh
            break
        seen_ids.add(doc.doc_id)
    if len(seen_ids) > 500:
        seen_ids = set(sorted(seen_ids)[-300:])
    return seen_ids, doc_count
