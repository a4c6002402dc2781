"""Resource presets (minimal / balanced / contributor / dedicated / custom) for CPU, memory, network and crawl
concurrency (reference infomesh/resources/profiles.py:17-165).  Each profile additionally carries the GPU-side
budget this build adds: how much HBM the resident index may take and the query micro-batch."""
from __future__ import annotations

from dataclasses import dataclass, fields, replace
from enum import StrEnum

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class ProfileName(StrEnum):
    MINIMAL = "minimal"
    BALANCED = "balanced"
    CONTRIBUTOR = "contributor"
    DEDICATED = "dedicated"
    CUSTOM = "custom"


@dataclass(frozen=True)
class ResourceProfile:
    name: ProfileName
    cpu_cores_limit: int            # 0 = unlimited
    cpu_nice: int
    memory_limit_mb: int            # 0 = unlimited
    disk_io_priority: str           # low | normal | high
    upload_limit_mbps: float
    download_limit_mbps: float
    max_concurrent_crawl: int
    llm_enabled: bool
    llm_off_peak_only: bool
    gpu_memory_fraction: float = 0.5    # share of HBM the resident index + model weights may occupy
    gpu_query_batch: int = 64


_P = ProfileName
PROFILES: dict[ProfileName, ResourceProfile] = {
    _P.MINIMAL: ResourceProfile(_P.MINIMAL, 1, 19, 512, "low", 0.5, 1.0, 1, False, True, 0.15, 8),
    _P.BALANCED: ResourceProfile(_P.BALANCED, 2, 10, 2048, "low", 2.0, 5.0, 3, True, True, 0.4, 32),
    _P.CONTRIBUTOR: ResourceProfile(_P.CONTRIBUTOR, 4, 5, 4096, "normal", 5.0, 10.0, 5, True, False, 0.6, 64),
    _P.DEDICATED: ResourceProfile(_P.DEDICATED, 0, 0, 0, "high", 25.0, 50.0, 10, True, False, 0.9, 128),
}


def get_profile(name: str | ProfileName) -> ResourceProfile:
    try:
        key = ProfileName(name)
    except ValueError:
        raise ValueError(f"Unknown profile '{name}'. Valid profiles: {', '.join(p.value for p in ProfileName)}") from None
    if key == ProfileName.CUSTOM:
        return replace(PROFILES[ProfileName.BALANCED], name=ProfileName.CUSTOM)
    return PROFILES[key]


def build_custom_profile(**overrides: object) -> ResourceProfile:
    """Balanced defaults with the given fields replaced; unknown names are reported and ignored."""
    known = {f.name for f in fields(ResourceProfile)} - {"name"}
    good = {}
    for k, v in overrides.items():
        if k in known:
            good[k] = v
        else:
            logger.warning("profile_unknown_field", field=k)
    return replace(get_profile(ProfileName.CUSTOM), **good)
