"""Configuration: frozen dataclasses per section, TOML file + ``INFOMESH_{SECTION}_{KEY}`` env overrides,
range clamping and enum whitelists, save-only-non-defaults.

Behavioural parity with reference infomesh/config.py:45-525 (same sections, keys, defaults, constraints and
precedence env > file > default) plus a ``[gpu]`` section for the B200 engine.
"""
from __future__ import annotations

import dataclasses
import json
import os
import sysconfig
import tomllib
from dataclasses import MISSING, dataclass, field, fields
from dataclasses import replace as dc_replace
from pathlib import Path
from typing import Any

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_DATA_DIR = Path(os.environ.get("INFOMESH_HOME", str(Path.home() / ".infomesh")))
DEFAULT_CONFIG_PATH = DEFAULT_DATA_DIR / "config.toml"


class NodeRole:
    """Which components a node runs (reference config.py:27-42)."""

    FULL = "full"
    CRAWLER = "crawler"
    SEARCH = "search"
    ALL = frozenset({FULL, CRAWLER, SEARCH})


@dataclass(frozen=True)
class NodeConfig:
    data_dir: Path = DEFAULT_DATA_DIR
    listen_port: int = 4001
    listen_address: str = "0.0.0.0"
    role: str = NodeRole.FULL
    log_level: str = "info"
    github_email: str = ""


@dataclass(frozen=True)
class CrawlConfig:
    max_concurrent: int = 5
    politeness_delay: float = 1.0
    max_depth: int = 0
    urls_per_hour: int = 60
    pending_per_domain: int = 10
    user_agent: str = "InfoMesh/0.1 (+https://github.com/dotnetpower/infomesh)"
    respect_robots: bool = True
    js_rendering: bool = False
    js_max_tabs: int = 3
    js_timeout_ms: int = 30_000
    js_max_memory_mb: int = 512
    rss_enabled: bool = False
    rss_default_interval: int = 900
    rss_max_feeds: int = 100
    rss_discovery: bool = True


@dataclass(frozen=True)
class NetworkConfig:
    upload_limit_mbps: float = 5.0
    download_limit_mbps: float = 10.0
    replication_factor: int = 3
    bootstrap_nodes: list[str] = field(default_factory=list)
    index_submit_peers: list[str] = field(default_factory=list)
    peer_acl: list[str] = field(default_factory=list)
    bootstrap_dns: bool = True
    bootstrap_github: bool = True
    bootstrap_dns_domain: str = "infomesh.io"
    encrypt: bool = True              # offer / accept the IMN1 encrypted, mutually authenticated channel (p2p/secure_channel.py)
    require_encrypted: bool = False   # refuse peers that only speak plaintext frames


@dataclass(frozen=True)
class IndexConfig:
    db_path: Path = field(default_factory=lambda: DEFAULT_DATA_DIR / "index.db")
    fts_tokenizer: str = "unicode61"
    max_doc_size_kb: int = 100
    vector_search: bool = False
    embedding_model: str = "all-MiniLM-L6-v2"


@dataclass(frozen=True)
class LLMConfig:
    enabled: bool = False
    runtime: str = "ollama"
    model: str = "qwen2.5:3b"
    off_peak_start: str = "23:00"
    off_peak_end: str = "07:00"
    timezone: str = "auto"


@dataclass(frozen=True)
class StorageConfig:
    compression_enabled: bool = True
    compression_level: int = 3
    max_cache_size_mb: int = 500
    max_index_size_gb: int = 50
    cache_ttl_days: int = 7


@dataclass(frozen=True)
class ResourceConfig:
    profile: str = "balanced"
    cpu_cores_limit: int = 2
    cpu_nice: int = 10
    memory_limit_mb: int = 2048
    disk_io_priority: str = "low"


@dataclass(frozen=True)
class DashboardConfig:
    bgm_auto_start: bool = False
    bgm_auto_install_mpv: bool = True
    bgm_volume: int = 50
    bgm_idle_stop: bool = False
    refresh_interval: float = 0.5
    theme: str = "catppuccin-mocha"


@dataclass(frozen=True)
class McpConfig:
    default_format: str = "text"
    max_response_chars: int = 0
    show_attribution: bool = True
    show_copyright: bool = True
    debug: bool = False


@dataclass(frozen=True)
class SearchConfig:
    feedback_tracking: bool = True
    cjk_auto_detect: bool = True


@dataclass(frozen=True)
class GpuConfig:
    """B200 engine settings (new in this rebuild)."""

    enabled: bool = False            # build the GPU engine inside AppContext
    devices: int = 0                 # 0 = all visible
    tp: int = 1                      # tensor-parallel degree of encoder / reranker / summariser
    shard_dtype: str = "bf16"        # dense shard storage: bf16 | fp8
    use_nvls: bool = True            # NVLink-SHARP multicast for one-to-all
    backend: str = "fused"           # fused | nccl  (A/B switch)
    encoder_model: str = "bge-small-en"
    reranker_model: str = "bge-reranker-base"
    summarizer_model: str = "t5-small"
    autotune_cache: str = ""
    device: int = 0                  # CUDA ordinal used by the single-process serving path
    query_batch: int = 64            # queries packed into one device pass by the store-backed GPU index
    rerank: bool = True              # cross-encoder second stage on the device
    encoder_path: str = ""           # HF-style checkpoint dir (config.json + model.safetensors + vocab.txt) for the encoder
    reranker_path: str = ""          # same for the cross-encoder; without checkpoints the device path serves BM25 order only
    allow_untrained_models: bool = False   # let random-init models rank (benchmarks / demos only)
    segments_dir: str = ""           # device segment files (one sub-directory per GPU when sharded): cold starts load these
                                     # instead of re-encoding the corpus, and every full build refreshes them


@dataclass(frozen=True)
class Config:
    node: NodeConfig = field(default_factory=NodeConfig)
    crawl: CrawlConfig = field(default_factory=CrawlConfig)
    network: NetworkConfig = field(default_factory=NetworkConfig)
    index: IndexConfig = field(default_factory=IndexConfig)
    llm: LLMConfig = field(default_factory=LLMConfig)
    storage: StorageConfig = field(default_factory=StorageConfig)
    resources: ResourceConfig = field(default_factory=ResourceConfig)
    dashboard: DashboardConfig = field(default_factory=DashboardConfig)
    mcp: McpConfig = field(default_factory=McpConfig)
    search: SearchConfig = field(default_factory=SearchConfig)
    gpu: GpuConfig = field(default_factory=GpuConfig)


_SECTIONS: dict[str, type] = {
    "node": NodeConfig, "crawl": CrawlConfig, "network": NetworkConfig, "index": IndexConfig, "llm": LLMConfig,
    "storage": StorageConfig, "resources": ResourceConfig, "dashboard": DashboardConfig, "mcp": McpConfig,
    "search": SearchConfig, "gpu": GpuConfig,
}

# numeric ranges (out-of-range values are clamped to the nearest bound, with a warning)
_VALUE_CONSTRAINTS: dict[str, tuple[float, float]] = {
    "listen_port": (1, 65535), "max_concurrent": (1, 100), "politeness_delay": (0.1, 60.0),
    "urls_per_hour": (1, 10000), "pending_per_domain": (1, 1000), "upload_limit_mbps": (0.1, 1000.0),
    "download_limit_mbps": (0.1, 1000.0), "replication_factor": (1, 10), "max_doc_size_kb": (1, 10240),
    "compression_level": (1, 22), "max_cache_size_mb": (10, 100000), "max_index_size_gb": (1, 10000),
    "cache_ttl_days": (1, 365), "cpu_cores_limit": (1, 256), "cpu_nice": (0, 19),
    "memory_limit_mb": (64, 1048576), "bgm_volume": (0, 100), "refresh_interval": (0.2, 5.0),
    "max_response_chars": (0, 10000000), "tp": (1, 8), "devices": (0, 8),
}

_ALLOWED_VALUES: dict[str, frozenset[str]] = {
    "role": frozenset(NodeRole.ALL),
    "log_level": frozenset({"debug", "info", "warning", "error", "critical"}),
    "runtime": frozenset({"ollama", "llama_cpp", "vllm", "b200"}),
    "profile": frozenset({"minimal", "balanced", "contributor", "dedicated"}),
    "disk_io_priority": frozenset({"low", "normal", "high"}),
    "fts_tokenizer": frozenset({"unicode61", "ascii", "porter", "trigram"}),
    "default_format": frozenset({"text", "json"}),
    "theme": frozenset({"catppuccin-mocha", "textual-dark", "textual-light", "dracula", "tokyo-night", "monokai",
                        "nord", "gruvbox", "textual-ansi", "solarized-light"}),
    "shard_dtype": frozenset({"bf16", "fp8"}),
    "backend": frozenset({"fused", "nccl"}),
}


def _env_override(section: str, key: str) -> str | None:
    return os.environ.get(f"INFOMESH_{section.upper()}_{key.upper()}")


def _field_type(f: dataclasses.Field) -> type:
    names = {"int": int, "float": float, "bool": bool, "str": str, "Path": Path, "list[str]": list}
    if isinstance(f.type, str) and f.type in names:
        return names[f.type]
    if f.default is not MISSING:
        return type(f.default)
    if f.default_factory is not MISSING:  # type: ignore[misc]
        return type(f.default_factory())  # type: ignore[misc]
    return str


def _coerce(value: Any, target: type) -> Any:
    """Coerce TOML / env values; env strings use 'true/1/yes' booleans and comma-separated lists."""
    if isinstance(value, str):
        if target is bool:
            return value.strip().lower() in ("true", "1", "yes")
        if target is int:
            return int(value)
        if target is float:
            return float(value)
        if target is Path:
            return Path(value).expanduser()
        if target is list:
            return [x.strip() for x in value.split(",") if x.strip()]
        return value
    if target is Path and value is not None:
        return Path(str(value)).expanduser()
    if target is float and isinstance(value, int) and not isinstance(value, bool):
        return float(value)
    if target is list and isinstance(value, (list, tuple)):
        return [str(x) for x in value]
    return value


def _default_of(f: dataclasses.Field) -> Any:
    if f.default is not MISSING:
        return f.default
    return f.default_factory()  # type: ignore[misc]


def _validate(key: str, value: Any, default: Any) -> Any:
    """Numbers outside their range are pulled to the nearest bound (an operator who asks for nice 25 gets 19, not the default);
    strings outside their vocabulary -- compared case-insensitively, kept as written -- fall back to the default."""
    if key in _VALUE_CONSTRAINTS and isinstance(value, (int, float)) and not isinstance(value, bool):
        lo, hi = _VALUE_CONSTRAINTS[key]
        if not lo <= value <= hi:
            clamped = type(value)(min(max(value, lo), hi))
            logger.warning("config_value_out_of_range", key=key, value=value, min=lo, max=hi, using=clamped)
            return clamped
    if key in _ALLOWED_VALUES and isinstance(value, str) and value.lower() not in _ALLOWED_VALUES[key]:
        logger.warning("config_invalid_value", key=key, value=value, allowed=sorted(_ALLOWED_VALUES[key]), using=default)
        return default
    return value


def _build_section(cls: type, raw: dict[str, Any], section: str):
    kwargs: dict[str, Any] = {}
    for f in fields(cls):
        default = _default_of(f)
        target = _field_type(f)
        value: Any = default
        if isinstance(raw, dict) and f.name in raw:
            value = raw[f.name]
        env = _env_override(section, f.name)
        if env is not None:
            value = env
        try:
            value = _coerce(value, target)
            if target in (int, float) and isinstance(value, bool):
                raise ValueError("bool is not numeric")
            if target is int and isinstance(value, float):
                value = int(value)
        except (ValueError, TypeError):
            logger.warning("config_value_invalid", section=section, key=f.name, value=str(value))
            value = default
        kwargs[f.name] = _validate(f.name, value, default)
    return cls(**kwargs)


def _load_default_bootstrap_nodes() -> list[str]:
    """Bundled bootstrap peers: ``bootstrap/nodes.json`` next to the package (or the installed share dir)."""
    here = Path(__file__).parent
    candidates = [here / "bootstrap" / "nodes.json", here.parent / "bootstrap" / "nodes.json",
                  Path(sysconfig.get_path("data") or "") / "share" / "infomesh" / "bootstrap" / "nodes.json"]
    for cand in candidates:
        try:
            if cand.exists():
                entries = json.loads(cand.read_text("utf-8"))
                addrs = [e["addr"] for e in entries if isinstance(e, dict) and "addr" in e]
                if addrs:
                    return addrs
        except (OSError, ValueError, KeyError, TypeError):
            continue
    return []


def load_config(config_path: Path | None = None) -> Config:
    """env > config.toml > defaults; creates the data directory."""
    path = Path(config_path) if config_path else DEFAULT_CONFIG_PATH
    raw: dict[str, Any] = {}
    if path.exists():
        try:
            with open(path, "rb") as fh:
                raw = tomllib.load(fh)
            logger.info("config_loaded", path=str(path))
        except (OSError, tomllib.TOMLDecodeError) as exc:
            logger.warning("config_unreadable", path=str(path), error=str(exc))
            raw = {}
    sections = {name: _build_section(cls, raw.get(name, {}), name) for name, cls in _SECTIONS.items()}
    cfg = Config(**sections)
    if not cfg.network.bootstrap_nodes:
        nodes = _load_default_bootstrap_nodes()
        if nodes:
            cfg = dc_replace(cfg, network=dc_replace(cfg.network, bootstrap_nodes=nodes))
    # a custom data_dir moves the default index path with it
    if cfg.node.data_dir != DEFAULT_DATA_DIR and cfg.index.db_path == DEFAULT_DATA_DIR / "index.db":
        cfg = dc_replace(cfg, index=dc_replace(cfg.index, db_path=cfg.node.data_dir / "index.db"))
    try:
        cfg.node.data_dir.expanduser().resolve().mkdir(parents=True, exist_ok=True)
    except OSError as exc:
        logger.warning("data_dir_unwritable", path=str(cfg.node.data_dir), error=str(exc))
    return cfg


def _toml_value(v: Any) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return repr(v)
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(_toml_value(x) for x in v) + "]"
    s = str(v).replace("\\", "\\\\").replace('"', '\\"')
    return f'"{s}"'


def save_config(config: Config, config_path: Path | None = None) -> None:
    """Write only the keys that differ from the defaults (reference config.py:460-525)."""
    path = Path(config_path) if config_path else DEFAULT_CONFIG_PATH
    path.parent.mkdir(parents=True, exist_ok=True)
    lines: list[str] = []
    for name, cls in _SECTIONS.items():
        cur = getattr(config, name)
        diff = []
        for f in fields(cls):
            val, dflt = getattr(cur, f.name), _default_of(f)
            if f.name == "db_path" and val == config.node.data_dir / "index.db":
                continue
            if val != dflt:
                diff.append(f"{f.name} = {_toml_value(val)}")
        if diff:
            lines.append(f"[{name}]")
            lines.extend(diff)
            lines.append("")
    tmp = path.with_suffix(".tmp")
    tmp.write_text("\n".join(lines), "utf-8")
    tmp.replace(path)


def config_to_dict(config: Config, redact: bool = True) -> dict[str, dict[str, Any]]:
    out: dict[str, dict[str, Any]] = {}
    for name in _SECTIONS:
        sec = {}
        for f in fields(getattr(config, name)):
            v = getattr(getattr(config, name), f.name)
            if isinstance(v, Path):
                v = str(v)
            if redact and f.name in ("github_email",) and v:
                v = "***"
            sec[f.name] = v
        out[name] = sec
    return out


def set_config_value(config: Config, dotted_key: str, value: str) -> Config:
    """``section.key = value`` with the same coercion / validation as loading (CLI ``config set``)."""
    section, _, key = dotted_key.partition(".")
    if section not in _SECTIONS or not key:
        raise KeyError(f"unknown config key {dotted_key!r}")
    cls = _SECTIONS[section]
    fmap = {f.name: f for f in fields(cls)}
    if key not in fmap:
        raise KeyError(f"unknown config key {dotted_key!r}")
    f = fmap[key]
    coerced = _validate(key, _coerce(value, _field_type(f)), _default_of(f))
    return dc_replace(config, **{section: dc_replace(getattr(config, section), **{key: coerced})})
