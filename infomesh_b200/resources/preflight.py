"""Preflight checks before a node starts: free disk space at the data dir (error < 500 MB, warning < 1 GB, crawl
pause < 200 MB), outbound HTTPS reachability, and — new in this build — GPU readiness (device visible, native
kernels loadable, HBM head-room) (reference infomesh/resources/preflight.py:27-193)."""
from __future__ import annotations

import shutil
import socket
from dataclasses import dataclass
from enum import StrEnum
from pathlib import Path

MIN_DISK_SPACE_MB = 500
LOW_DISK_SPACE_MB = 200
CONNECTIVITY_TARGETS: list[tuple[str, int]] = [("docs.python.org", 443), ("developer.mozilla.org", 443), ("1.1.1.1", 443)]
CONNECT_TIMEOUT = 3.0


class IssueSeverity(StrEnum):
    ERROR = "error"
    WARNING = "warning"


@dataclass(frozen=True)
class PreflightIssue:
    severity: IssueSeverity
    check: str
    message: str


def get_disk_free_mb(path: Path) -> float:
    p = Path(path)
    while not p.exists() and p != p.parent:
        p = p.parent
    return shutil.disk_usage(p).free / 2 ** 20


def check_disk_space(data_dir: Path) -> list[PreflightIssue]:
    try:
        free = get_disk_free_mb(data_dir)
    except OSError as exc:
        return [PreflightIssue(IssueSeverity.ERROR, "disk_space", f"Cannot check disk space: {exc}")]
    if free < MIN_DISK_SPACE_MB:
        return [PreflightIssue(IssueSeverity.ERROR, "disk_space",
                               f"Insufficient disk space: {free:.0f} MB free (minimum {MIN_DISK_SPACE_MB} MB required). "
                               f"Free up space at {data_dir} before starting.")]
    if free < 2 * MIN_DISK_SPACE_MB:
        return [PreflightIssue(IssueSeverity.WARNING, "disk_space",
                               f"Low disk space: {free:.0f} MB free. Consider freeing space at {data_dir}.")]
    return []


def is_disk_critically_low(data_dir: Path) -> bool:
    try:
        return get_disk_free_mb(data_dir) < LOW_DISK_SPACE_MB
    except OSError:
        return False


def check_outbound_connectivity(targets: list[tuple[str, int]] | None = None) -> list[PreflightIssue]:
    dead = []
    for host, port in (targets if targets is not None else CONNECTIVITY_TARGETS):
        try:
            with socket.create_connection((host, port), timeout=CONNECT_TIMEOUT):
                pass
        except (OSError, TimeoutError):
            dead.append(f"{host}:{port}")
    n = len(targets if targets is not None else CONNECTIVITY_TARGETS)
    if dead and len(dead) == n:
        return [PreflightIssue(IssueSeverity.ERROR, "network",
                               "No outbound connectivity — cannot reach any external host. "
                               f"Tried: {', '.join(dead)}. Check your firewall or proxy settings.")]
    if dead:
        return [PreflightIssue(IssueSeverity.WARNING, "network",
                               f"Some hosts unreachable: {', '.join(dead)}. Crawling may be limited.")]
    return []


def check_gpu(min_free_gb: float = 2.0) -> list[PreflightIssue]:
    """GPU plane readiness.  A missing GPU is only a warning: every component has a CPU path."""
    try:
        import torch
    except ImportError:
        return [PreflightIssue(IssueSeverity.WARNING, "gpu", "PyTorch not importable: GPU search/encode disabled.")]
    if not torch.cuda.is_available():
        return [PreflightIssue(IssueSeverity.WARNING, "gpu", "No CUDA device visible: running the CPU search plane only.")]
    out = []
    major, minor = torch.cuda.get_device_capability(0)
    if (major, minor) != (10, 0):
        out.append(PreflightIssue(IssueSeverity.ERROR, "gpu",
                                  f"Device is sm_{major}{minor}; the native kernels are built for sm_100a (B200) only."))
    from infomesh_b200 import _native

    if not _native.available():
        out.append(PreflightIssue(IssueSeverity.ERROR, "gpu", "Native kernel library is not built: run `python -m infomesh_b200.build`."))
    free, _ = torch.cuda.mem_get_info()
    if free / 2 ** 30 < min_free_gb:
        out.append(PreflightIssue(IssueSeverity.WARNING, "gpu", f"Only {free / 2 ** 30:.1f} GB of HBM free."))
    return out


def run_preflight_checks(data_dir: Path, *, network: bool = True, gpu: bool = False) -> list[PreflightIssue]:
    issues = check_disk_space(data_dir)
    if network:
        issues += check_outbound_connectivity()
    if gpu:
        issues += check_gpu()
    return issues
