"""Diagnostics: peer-count partition alerts, DHT latency benchmarking and the ``infomesh doctor`` report
(reference infomesh/diagnostics.py:22-373).  The doctor additionally checks the GPU plane."""
from __future__ import annotations

import shutil
import socket
import sys
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Awaitable, Callable

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass
class PartitionAlert:
    timestamp: float
    previous_peers: int
    current_peers: int
    drop_pct: float
    severity: str           # warning | critical


class PartitionDetector:
    """Compares the current peer count with the mean of the previous (up to nine) samples."""

    def __init__(self, *, warning_threshold: float = 0.5, critical_threshold: float = 0.8, min_peers_for_alert: int = 3):
        self._warn, self._crit, self._min = warning_threshold, critical_threshold, min_peers_for_alert
        self._history: list[tuple[float, int]] = []
        self._alerts: list[PartitionAlert] = []

    def record(self, peer_count: int) -> PartitionAlert | None:
        now = time.time()
        self._history = (self._history + [(now, peer_count)])[-60:]
        if len(self._history) < 3:
            return None
        prev = [c for _, c in self._history[-10:-1]]
        avg = sum(prev) / len(prev)
        if avg < self._min:
            return None
        drop = 1.0 - peer_count / avg
        if drop < self._warn:
            return None
        alert = PartitionAlert(now, int(avg), peer_count, round(drop * 100, 1), "critical" if drop >= self._crit else "warning")
        self._alerts.append(alert)
        logger.warning("network_partition_detected", severity=alert.severity, peer_drop_pct=alert.drop_pct)
        return alert

    @property
    def alerts(self) -> list[PartitionAlert]:
        return list(self._alerts)


@dataclass
class DHTBenchResult:
    operation: str
    samples: int
    avg_ms: float
    p50_ms: float
    p95_ms: float
    p99_ms: float
    errors: int


def _percentile(values: list[float], pct: float) -> float:
    if not values:
        return 0.0
    s = sorted(values)
    return s[min(int(len(s) * pct / 100), len(s) - 1)]


async def benchmark_dht(op: Callable[[int], Awaitable[Any]], *, operation: str = "get", samples: int = 20) -> DHTBenchResult:
    """Time ``await op(i)`` ``samples`` times (e.g. ``lambda i: dht.get(f"k{i}")``)."""
    times, errors = [], 0
    for i in range(samples):
        t0 = time.monotonic()
        try:
            await op(i)
            times.append((time.monotonic() - t0) * 1000)
        except Exception:  # noqa: BLE001
            errors += 1
    avg = sum(times) / len(times) if times else 0.0
    return DHTBenchResult(operation, len(times), round(avg, 2), round(_percentile(times, 50), 2), round(_percentile(times, 95), 2),
                          round(_percentile(times, 99), 2), errors)


@dataclass
class DiagnosticCheck:
    name: str
    status: str             # ok | warning | error
    message: str
    details: str = ""


@dataclass
class DiagnosticReport:
    checks: list[DiagnosticCheck] = field(default_factory=list)
    timestamp: float = field(default_factory=time.time)

    @property
    def ok(self) -> bool:
        return all(c.status == "ok" for c in self.checks)

    @property
    def summary(self) -> str:
        n = {s: sum(c.status == s for c in self.checks) for s in ("ok", "warning", "error")}
        return f"{n['ok']} ok, {n['warning']} warnings, {n['error']} errors"


def _port_in_use(port: int) -> bool | None:
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
            s.settimeout(1)
            return s.connect_ex(("127.0.0.1", port)) == 0
    except OSError:
        return None


def run_diagnostics(data_dir: Path | None = None, *, p2p_port: int = 4001, admin_port: int = 8080) -> DiagnosticReport:
    from infomesh_b200.config import DEFAULT_DATA_DIR

    d = Path(data_dir) if data_dir else DEFAULT_DATA_DIR
    rep = DiagnosticReport()
    add = lambda name, status, msg, details="": rep.checks.append(DiagnosticCheck(name, status, msg, details))  # noqa: E731
    add("data_dir", "ok" if d.exists() else "error", f"Data directory {'exists' if d.exists() else 'missing'}: {d}")
    have_key = (d / "keys" / "private.pem").exists() or (d / "keys" / "private.key").exists()
    add("key_pair", "ok" if have_key else "warning", "Ed25519 key pair present" if have_key else "No key pair found (will be generated on start)")
    db = d / "index.db"
    if db.exists():
        add("index_db", "ok", f"Index DB: {db.stat().st_size / 2 ** 20:.1f} MB")
    else:
        add("index_db", "warning", "No index database (empty node)")
    add("config", "ok", "Config file present" if (d / "config.toml").exists() else "Using default config (no config.toml)")
    used = _port_in_use(p2p_port)
    add("p2p_port", "warning" if used is None else "ok",
        f"Cannot check port {p2p_port}" if used is None else
        f"Port {p2p_port} is in use (node may be running)" if used else f"Port {p2p_port} available")
    up = _port_in_use(admin_port)
    add("admin_api", "ok" if up else "warning", f"Admin API responding on port {admin_port}" if up else "Admin API not responding")
    free_gb = shutil.disk_usage(d if d.exists() else "/").free / 2 ** 30
    add("disk_space", "error" if free_gb < 1 else "warning" if free_gb < 5 else "ok",
        (f"Low disk space: {free_gb:.1f} GB free" if free_gb < 1 else f"Disk space low: {free_gb:.1f} GB free" if free_gb < 5
         else f"Disk space: {free_gb:.1f} GB free"))
    add("python", "ok", f"Python {sys.version_info.major}.{sys.version_info.minor}")
    add("credits", "ok", "Credit ledger present" if (d / "credits.db").exists() else "No credit ledger (will be created)")
    from infomesh_b200.p2p.bootstrap import bundled_nodes

    nodes = bundled_nodes()
    add("bootstrap", "ok" if nodes else "warning", f"{len(nodes)} bundled bootstrap node(s) (check with `infomesh status`)")
    # --- GPU plane
    try:
        from infomesh_b200.resources.preflight import IssueSeverity, check_gpu

        issues = check_gpu()
        if not issues:
            import torch

            add("gpu", "ok", f"{torch.cuda.get_device_name(0)} x{torch.cuda.device_count()}, native kernels loadable")
        for it in issues:
            add("gpu", "error" if it.severity == IssueSeverity.ERROR else "warning", it.message)
    except Exception as exc:  # noqa: BLE001
        add("gpu", "warning", f"GPU check failed: {exc}")
    # --- GPU health (ECC / Xid / throttle) through NVML
    try:
        from infomesh_b200.resources.gpu_health import GpuHealthMonitor

        mon = GpuHealthMonitor()
        hs = [h for h in mon.poll() if h.available]
        mon.close()
        bad = [h for h in hs if not h.healthy]
        if bad:
            add("gpu_health", "error", "; ".join(f"GPU {h.index}: uncorrected ECC {h.ecc_uncorrected}, Xid {h.last_xid} x{h.xid_events}" for h in bad))
        elif hs:
            hot = sorted({r for h in hs for r in h.throttle_reasons if r != "sw_power_cap"})
            add("gpu_health", "warning" if hot else "ok", f"{len(hs)} GPU(s) healthy (no uncorrected ECC errors, no critical Xid)"
                + (f"; throttling: {', '.join(hot)}" if hot else ""))
    except Exception as exc:  # noqa: BLE001
        add("gpu_health", "warning", f"GPU health check failed: {exc}")
    return rep
