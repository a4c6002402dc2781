"""Bootstrap-node discovery: bundled static list, on-disk cache (1 h TTL), DNS SRV ``_infomesh._tcp.<domain>``,
DNS TXT ``_infomesh-bootstrap.<domain>``, a hosted nodes.json; results merged + de-duplicated, TCP health probes,
seed-peer selection (0.6 recency + 0.4 uptime) and a 10 req/min limiter for nodes acting as bootstrap servers
(reference infomesh/p2p/bootstrap.py:30-685).  The network sources degrade to empty lists without connectivity."""
from __future__ import annotations

import asyncio
import contextlib
import json
import struct
import time
from dataclasses import asdict, dataclass, field
from pathlib import Path

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_DNS_DOMAIN = "infomesh.io"
SRV_SERVICE = "_infomesh._tcp"
TXT_PREFIX = "_infomesh-bootstrap"
GITHUB_NODES_URL = "https://raw.githubusercontent.com/dotnetpower/infomesh/main/bootstrap/nodes.json"
BOOTSTRAP_CACHE_FILE = "bootstrap_cache.json"
BOOTSTRAP_CACHE_TTL = 3600
BOOTSTRAP_RATE_LIMIT_PER_MIN = 10
BOOTSTRAP_MAX_PEERS_SEED = 50
HEALTH_CHECK_TIMEOUT = 5.0
HEALTH_CHECK_INTERVAL = 60.0


@dataclass
class BootstrapNode:
    addr: str
    source: str                 # static | dns_srv | dns_txt | github | cache
    region: str = ""
    last_seen: float = 0.0
    healthy: bool = True
    latency_ms: float = 0.0

    @property
    def host_port(self) -> tuple[str, int]:
        parts = self.addr.split("/")
        host, port = "", 0
        for tag, val in zip(parts, parts[1:]):
            if tag in ("ip4", "ip6", "dns4", "dns6", "dns"):
                host = val
            elif tag == "tcp":
                with contextlib.suppress(ValueError):
                    port = int(val)
        return host, port


@dataclass
class BootstrapResult:
    nodes: list[BootstrapNode] = field(default_factory=list)
    sources_tried: list[str] = field(default_factory=list)
    sources_succeeded: list[str] = field(default_factory=list)
    discovery_ms: float = 0.0

    @property
    def addrs(self) -> list[str]:
        return list(dict.fromkeys(n.addr for n in self.nodes))


@dataclass
class BootstrapHealth:
    addr: str
    reachable: bool
    latency_ms: float = 0.0
    peer_count: int = 0
    uptime_seconds: float = 0.0
    last_check: float = 0.0


def bundled_nodes() -> list[dict[str, str]]:
    path = Path(__file__).resolve().parent.parent / "bootstrap" / "nodes.json"
    try:
        data = json.loads(path.read_text())
        return data if isinstance(data, list) else []
    except (OSError, ValueError):
        return []


def discover_from_static(nodes_json: list[dict[str, str]]) -> list[BootstrapNode]:
    now = time.time()
    return [BootstrapNode(e["addr"], "static", e.get("region", ""), now) for e in nodes_json
            if isinstance(e, dict) and "addr" in e]


# ------------------------------------------------------------------ DNS (stdlib-only resolver over UDP)
def _dns_query(name: str, qtype: int, *, server: str | None = None, timeout: float = 3.0) -> list[bytes]:
    """Minimal RFC 1035 client: returns the RDATA blobs of answers of ``qtype``."""
    import random
    import socket

    if server is None:
        server = "127.0.0.53"
        with contextlib.suppress(OSError):
            for line in Path("/etc/resolv.conf").read_text().splitlines():
                if line.startswith("nameserver"):
                    server = line.split()[1]
                    break
    qid = random.randrange(65536)
    q = struct.pack("!HHHHHH", qid, 0x0100, 1, 0, 0, 0)
    q += b"".join(bytes([len(p)]) + p.encode() for p in name.strip(".").split(".")) + b"\0" + struct.pack("!HH", qtype, 1)
    with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:
        s.settimeout(timeout)
        s.sendto(q, (server, 53))
        data, _ = s.recvfrom(4096)
    if len(data) < 12 or struct.unpack("!H", data[:2])[0] != qid:
        return []
    _, _, qd, an, _, _ = struct.unpack("!HHHHHH", data[:12])

    def skip_name(off: int) -> int:
        while off < len(data):
            n = data[off]
            if n == 0:
                return off + 1
            if n & 0xC0 == 0xC0:
                return off + 2
            off += 1 + n
        return off

    off = 12
    for _ in range(qd):
        off = skip_name(off) + 4
    out: list[bytes] = []
    for _ in range(an):
        off = skip_name(off)
        if off + 10 > len(data):
            break
        rtype, _, _, rdlen = struct.unpack("!HHIH", data[off:off + 10])
        off += 10
        if rtype == qtype:
            out.append(_expand(data, off, rdlen, qtype))
        off += rdlen
    return out


def _read_name(data: bytes, off: int, depth: int = 0) -> str:
    labels = []
    while off < len(data) and depth < 8:
        n = data[off]
        if n == 0:
            break
        if n & 0xC0 == 0xC0:
            ptr = ((n & 0x3F) << 8) | data[off + 1]
            labels.append(_read_name(data, ptr, depth + 1))
            break
        labels.append(data[off + 1:off + 1 + n].decode("ascii", "replace"))
        off += 1 + n
    return ".".join(x for x in labels if x)


def _expand(data: bytes, off: int, rdlen: int, qtype: int) -> bytes:
    if qtype == 33 and rdlen >= 7:      # SRV: prio, weight, port, target (possibly compressed)
        prio, weight, port = struct.unpack("!HHH", data[off:off + 6])
        return struct.pack("!HHH", prio, weight, port) + _read_name(data, off + 6).encode()
    return data[off:off + rdlen]


def _resolve_srv(name: str) -> list[tuple[str, int]]:
    out = []
    for r in _dns_query(name, 33):
        _, _, port = struct.unpack("!HHH", r[:6])
        host = r[6:].decode()
        if host and port:
            out.append((host, port))
    return out


def _resolve_txt(name: str) -> list[str]:
    out = []
    for r in _dns_query(name, 16):
        parts, off = [], 0
        while off < len(r):
            n = r[off]
            parts.append(r[off + 1:off + 1 + n].decode("utf-8", "replace"))
            off += 1 + n
        out.append("".join(parts))
    return out


async def discover_from_dns_srv(domain: str = DEFAULT_DNS_DOMAIN) -> list[BootstrapNode]:
    try:
        answers = await asyncio.get_running_loop().run_in_executor(None, _resolve_srv, f"{SRV_SERVICE}.{domain}")
    except Exception as exc:  # noqa: BLE001
        logger.debug("bootstrap_dns_srv_failed", domain=domain, error=str(exc))
        return []
    now = time.time()
    return [BootstrapNode(f"/dns4/{h}/tcp/{p}", "dns_srv", last_seen=now) for h, p in answers]


async def discover_from_dns_txt(domain: str = DEFAULT_DNS_DOMAIN) -> list[BootstrapNode]:
    try:
        records = await asyncio.get_running_loop().run_in_executor(None, _resolve_txt, f"{TXT_PREFIX}.{domain}")
    except Exception as exc:  # noqa: BLE001
        logger.debug("bootstrap_dns_txt_failed", domain=domain, error=str(exc))
        return []
    now = time.time()
    return [BootstrapNode(r.strip().strip('"'), "dns_txt", last_seen=now) for r in records
            if r.strip().strip('"').startswith("/")]


async def discover_from_github(url: str = GITHUB_NODES_URL, timeout: float = 10.0) -> list[BootstrapNode]:
    def _fetch():
        import urllib.request

        with urllib.request.urlopen(url, timeout=timeout) as resp:  # noqa: S310 — fixed https URL from config
            return json.loads(resp.read(1 << 20))

    try:
        entries = await asyncio.get_running_loop().run_in_executor(None, _fetch)
    except Exception as exc:  # noqa: BLE001
        logger.debug("bootstrap_github_failed", url=url, error=str(exc))
        return []
    now = time.time()
    return [BootstrapNode(e["addr"], "github", e.get("region", ""), now) for e in (entries if isinstance(entries, list) else [])
            if isinstance(e, dict) and "addr" in e]


# ------------------------------------------------------------------ cache
def _load_cache(cache_dir: Path) -> list[BootstrapNode]:
    path = Path(cache_dir) / BOOTSTRAP_CACHE_FILE
    try:
        blob = json.loads(path.read_text())
    except (OSError, ValueError):
        return []
    if not isinstance(blob, dict) or time.time() - float(blob.get("saved_at", 0)) > BOOTSTRAP_CACHE_TTL:
        return []
    out = []
    for e in blob.get("nodes", []):
        if isinstance(e, dict) and isinstance(e.get("addr"), str):
            out.append(BootstrapNode(e["addr"], "cache", str(e.get("region", "")), float(e.get("last_seen", 0.0))))
    return out


def _save_cache(cache_dir: Path, nodes: list[BootstrapNode]) -> None:
    path = Path(cache_dir) / BOOTSTRAP_CACHE_FILE
    try:
        path.parent.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(".tmp")
        tmp.write_text(json.dumps({"saved_at": time.time(), "nodes": [asdict(n) for n in nodes]}))
        tmp.replace(path)
    except OSError as exc:
        logger.debug("bootstrap_cache_save_failed", error=str(exc))


async def discover_bootstrap_nodes(static_nodes: list[dict[str, str]] | None = None, dns_domain: str = DEFAULT_DNS_DOMAIN,
                                   github_url: str = GITHUB_NODES_URL, cache_dir: Path | None = None,
                                   use_dns: bool = True, use_github: bool = True) -> BootstrapResult:
    t0 = time.time()
    res = BootstrapResult()
    found: list[BootstrapNode] = []

    def take(name: str, nodes: list[BootstrapNode]) -> None:
        res.sources_tried.append(name)
        if nodes:
            found.extend(nodes)
            res.sources_succeeded.append(name)

    if static_nodes:
        take("static", discover_from_static(static_nodes))
    if cache_dir:
        take("cache", _load_cache(cache_dir))
    jobs: list[tuple[str, asyncio.Future]] = []
    if use_dns:
        jobs += [("dns_srv", discover_from_dns_srv(dns_domain)), ("dns_txt", discover_from_dns_txt(dns_domain))]
    if use_github:
        jobs.append(("github", discover_from_github(github_url)))
    if jobs:
        done = await asyncio.gather(*(j for _, j in jobs), return_exceptions=True)
        for (name, _), r in zip(jobs, done):
            take(name, r if isinstance(r, list) else [])
    seen: set[str] = set()
    for n in found:
        if n.addr not in seen:
            seen.add(n.addr)
            res.nodes.append(n)
    res.discovery_ms = (time.time() - t0) * 1000
    if cache_dir and res.nodes:
        _save_cache(cache_dir, res.nodes)
    return res


async def check_bootstrap_health(node: BootstrapNode, timeout: float = HEALTH_CHECK_TIMEOUT) -> BootstrapHealth:
    host, port = node.host_port
    now = time.time()
    if not host or not port:
        return BootstrapHealth(node.addr, False, last_check=now)
    t0 = time.monotonic()
    try:
        _, w = await asyncio.wait_for(asyncio.open_connection(host, port), timeout=timeout)
        w.close()
        with contextlib.suppress(Exception):
            await w.wait_closed()
    except (OSError, asyncio.TimeoutError):
        node.healthy = False
        return BootstrapHealth(node.addr, False, last_check=now)
    node.healthy, node.latency_ms = True, (time.monotonic() - t0) * 1000
    return BootstrapHealth(node.addr, True, node.latency_ms, last_check=now)


async def check_all_bootstrap_health(nodes: list[BootstrapNode], timeout: float = HEALTH_CHECK_TIMEOUT) -> list[BootstrapHealth]:
    return list(await asyncio.gather(*(check_bootstrap_health(n, timeout) for n in nodes)))


def select_seed_peers(known_peers: list[dict[str, object]], max_peers: int = BOOTSTRAP_MAX_PEERS_SEED) -> list[dict[str, object]]:
    now = time.time()

    def num(v: object) -> float:
        return float(v) if isinstance(v, (int, float)) else 0.0

    def score(p: dict[str, object]) -> float:
        recency = max(0.0, 1.0 - (now - num(p.get("last_seen"))) / 86400)
        return 0.6 * recency + 0.4 * min(num(p.get("uptime")) / 86400, 1.0)

    return sorted(known_peers, key=score, reverse=True)[:max_peers]


class BootstrapRateLimiter:
    def __init__(self, max_per_minute: int = BOOTSTRAP_RATE_LIMIT_PER_MIN, window_seconds: float = 60.0):
        self._max, self._window = max_per_minute, window_seconds
        self._hits: dict[str, list[float]] = {}

    def allow(self, client_id: str) -> bool:
        now = time.time()
        live = [t for t in self._hits.get(client_id, []) if t > now - self._window]
        if len(live) >= self._max:
            self._hits[client_id] = live
            return False
        live.append(now)
        self._hits[client_id] = live
        return True

    def reset(self, client_id: str) -> None:
        self._hits.pop(client_id, None)

    @property
    def tracked_clients(self) -> int:
        return len(self._hits)

    def cleanup(self) -> int:
        cutoff = time.time() - self._window
        dead = [k for k, v in self._hits.items() if all(t <= cutoff for t in v)]
        for k in dead:
            del self._hits[k]
        return len(dead)
