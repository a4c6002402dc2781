"""Small helpers shared by the dashboard views (reference infomesh/dashboard/utils.py:11-201)."""
from __future__ import annotations

import json
import time
from typing import Any

from infomesh_b200.config import Config
from infomesh_b200.runtime import read_live_pid, read_runtime_status


def format_uptime(seconds: float) -> str:
    s = int(max(seconds, 0))
    d, s = divmod(s, 86400)
    h, s = divmod(s, 3600)
    m, s = divmod(s, 60)
    return f"{d}d {h}h {m}m" if d else f"{h}h {m}m {s}s" if h else f"{m}m {s}s" if m else f"{s}s"


def format_bytes(n: int | float) -> str:
    v = float(n)
    for unit in ("B", "KB", "MB", "GB", "TB"):
        if abs(v) < 1024 or unit == "TB":
            return f"{v:.0f} {unit}" if unit == "B" else f"{v:.1f} {unit}"
        v /= 1024
    return f"{v:.1f} TB"


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


def tier_label(tier: Any) -> str:
    name = getattr(tier, "name", str(tier))
    return {"TIER_1": "Tier 1 (×1.0 search cost)", "TIER_2": "⭐ Tier 2 (×0.67)", "TIER_3": "⭐⭐ Tier 3 (×0.33)"}.get(name, name)


def format_doc_line(url: str, title: str, width: int = 90) -> str:
    label = (title or url).strip().replace("\n", " ")
    line = f"{label}  ·  {url}" if title else url
    return line if len(line) <= width else line[:width - 1] + "…"


def push_new_docs_to_log(log: Any, docs: list[Any], seen: set[int]) -> int:
    """Write not-yet-seen recent documents (oldest first) to a log widget exposing ``write_line``."""
    fresh = [d for d in reversed(docs) if d.doc_id not in seen]
    for d in fresh:
        seen.add(d.doc_id)
        log.write_line(f"{time.strftime('%H:%M:%S', time.localtime(d.crawled_at))}  ✔ {format_doc_line(d.url, d.title)}")
    return len(fresh)
