"""Local admin HTTP API (FastAPI): health / readiness / status / config / index / credits / peers / analytics /
Prometheus metrics / OpenAPI / HTML dashboard; loopback clients only, optional ``x-api-key``, 10 requests per second,
security headers (reference infomesh/api/local_api.py:50-791).

Additions of this build: ``/search`` is served by the attached :class:`ToolRuntime` (so it uses the GPU hybrid pipeline
when present) instead of re-opening the SQLite store per request; ``/gpu/stats``; ``POST /index/submit`` — the HTTP
bridge the crawler-role ``IndexSubmitSender`` posts msgpack frames to; and :func:`serve_admin_api`, which actually
runs the app (in the reference the app factory exists but nothing serves it)."""
from __future__ import annotations

import hmac
import json
import os
import shutil
import time
from collections import deque
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Any

from infomesh_b200 import __version__
from infomesh_b200.config import Config, load_config
from infomesh_b200.runtime import read_runtime_status
from infomesh_b200.utils.log import get_logger

try:  # module-level so FastAPI can resolve the postponed ``Request`` annotations of the route functions
    from fastapi import FastAPI, Request
    from fastapi.responses import HTMLResponse, JSONResponse, PlainTextResponse
except ImportError:  # pragma: no cover — create_admin_app() raises a clear error instead
    FastAPI = Request = HTMLResponse = JSONResponse = PlainTextResponse = None  # type: ignore[assignment,misc]

logger = get_logger(__name__)

_LOCAL_ADMIN_HOSTS = frozenset({"127.0.0.1", "::1", "localhost", "testclient"})
# never leave the box (filesystem layout, peer addresses) — same four keys as reference api/local_api.py:39-46 ...
_ALWAYS_REDACTED = frozenset({"data_dir", "db_path", "bootstrap_nodes", "persist_dir"})
# ... plus credentials and identity, masked when set
_SENSITIVE_KEYS = frozenset({"api_key", "secret", "password", "token", "private_key", "github_email"})
_RATE_LIMIT_PER_SECOND = 10
_RATE_MAX_CLIENTS = 10_000


@dataclass
class AdminState:
    config: Config
    config_path: Path | None = None
    start_time: float = field(default_factory=time.time)
    total_searches: int = 0
    total_crawls: int = 0
    total_fetches: int = 0
    avg_latency_ms: float = 0.0
    _latency_sum: float = 0.0
    runtime: Any | None = None          # mcp.handlers.ToolRuntime (optional)
    index_submit_receiver: Any | None = None

    def record_search(self, latency_ms: float) -> None:
        self.total_searches += 1
        self._latency_sum += latency_ms
        self.avg_latency_ms = self._latency_sum / self.total_searches

    def record_crawl(self) -> None:
        self.total_crawls += 1

    def record_fetch(self) -> None:
        self.total_fetches += 1


def _format_duration(seconds: float) -> str:
    if seconds < 60:
        return f"{seconds:.0f}s"
    if seconds < 3600:
        return f"{seconds / 60:.0f}m {seconds % 60:.0f}s"
    return f"{seconds / 3600:.0f}h {(seconds % 3600) / 60:.0f}m"


def _redact_paths(cfg: dict[str, Any]) -> None:
    for k, v in list(cfg.items()):
        if k in _ALWAYS_REDACTED:
            cfg[k] = "***REDACTED***"
        elif k in _SENSITIVE_KEYS:
            cfg[k] = "***REDACTED***" if v else v
        elif isinstance(v, Path):
            cfg[k] = str(v)
        elif isinstance(v, dict):
            _redact_paths(v)


def _get_index_stats(config: Config) -> dict[str, Any]:
    db = Path(config.index.db_path)
    if not db.exists():
        return {"document_count": 0, "db_size_mb": 0.0}
    try:
        from infomesh_b200.index.local_store import LocalStore

        with LocalStore(db, compression_enabled=config.storage.compression_enabled, compression_level=config.storage.compression_level) as st:
            n = st.get_stats().get("document_count", 0)
        return {"document_count": n, "db_size_mb": round(db.stat().st_size / 2 ** 20, 2)}
    except Exception:  # noqa: BLE001
        return {"error": "unable to read index stats"}


def _get_credit_stats(config: Config) -> dict[str, Any]:
    path = Path(config.node.data_dir) / "credits.db"
    empty = {"balance": 0.0, "total_earned": 0.0, "total_spent": 0.0}
    if not path.exists():
        return empty
    try:
        from infomesh_b200.credits.ledger import CreditLedger

        led = CreditLedger(path)
        try:
            s = led.stats()
        finally:
            led.close()
        return {"total_earned": s.total_earned, "total_spent": s.total_spent, "balance": s.balance}
    except Exception:  # noqa: BLE001
        return empty


def create_admin_app(config: Config | None = None, config_path: Path | None = None, *, runtime: Any | None = None,
                     index_submit_receiver: Any | None = None):
    if FastAPI is None:
        raise RuntimeError("the admin API needs the 'fastapi' package")
    cfg = config or load_config(config_path)
    state = AdminState(cfg, config_path, runtime=runtime, index_submit_receiver=index_submit_receiver)
    app = FastAPI(title="InfoMesh Admin API", description="Local node administration and monitoring.", version=__version__,
                  docs_url="/docs" if cfg.node.log_level.lower() == "debug" else None, redoc_url=None)
    app.state.admin = state
    buckets: dict[str, deque[float]] = {}

    @app.middleware("http")
    async def guard(request: Request, call_next):
        client = request.client.host if request.client else None
        if client not in _LOCAL_ADMIN_HOSTS:
            return JSONResponse(status_code=403, content={"detail": "Admin API is only accessible from localhost"})
        key = os.environ.get("INFOMESH_API_KEY")
        if key is not None and not hmac.compare_digest(request.headers.get("x-api-key", "").encode(), key.encode()):
            return JSONResponse(status_code=401, content={"detail": "Invalid API key"})
        now = time.time()
        if client not in buckets and len(buckets) >= _RATE_MAX_CLIENTS:
            for k in [k for k, v in buckets.items() if not v or v[-1] < now - 10.0]:
                del buckets[k]
        q = buckets.setdefault(client, deque(maxlen=_RATE_LIMIT_PER_SECOND))
        while q and now - q[0] >= 1.0:
            q.popleft()
        if len(q) >= _RATE_LIMIT_PER_SECOND:
            return JSONResponse(status_code=429, content={"detail": "Rate limit exceeded"})
        q.append(now)
        resp = await call_next(request)
        resp.headers["X-Content-Type-Options"] = "nosniff"
        resp.headers["X-Frame-Options"] = "DENY"
        resp.headers["Content-Security-Policy"] = "default-src 'self'; script-src 'self'; style-src 'unsafe-inline'; object-src 'none'; base-uri 'none'"
        return resp

    def st(request: Request) -> AdminState:
        return request.app.state.admin

    @app.get("/health")
    async def health(request: Request) -> dict[str, Any]:
        s = st(request)
        if not request.query_params.get("detail", ""):
            return {"status": "ok"}
        out: dict[str, str] = {"status": "ok", "db": "ok" if Path(s.config.index.db_path).exists() else "missing"}
        try:
            d = Path(s.config.node.data_dir)
            free = shutil.disk_usage(d if d.exists() else "/").free
            out["disk_free_gb"] = str(round(free / 2 ** 30, 1))
            out["disk"] = "ok" if free >= 2 ** 30 else "low"
            if free < 2 ** 30:
                out["status"] = "degraded"
        except OSError:
            out["disk"] = "unknown"
        try:
            import psutil

            pct = psutil.virtual_memory().percent
            out["memory_pct"], out["memory"] = str(round(pct, 1)), "critical" if pct > 95 else "ok"
            if pct > 95:
                out["status"] = "degraded"
        except ImportError:
            out["memory"] = "unknown"
        out["uptime_s"] = str(round(time.time() - s.start_time))
        rt = read_runtime_status(s.config.node.data_dir)
        if rt:
            out["runtime"] = str(rt.get("status", "unknown"))
            out["runtime_degrade_level"] = str(rt.get("degrade_level", "unknown"))
            out["runtime_process_memory_mb"] = str(rt.get("process_memory_mb", "unknown"))
        return out

    @app.get("/search")
    async def search_api(request: Request, q: str = "", limit: int = 5) -> dict[str, Any]:
        s = st(request)
        if not q:
            return {"results": [], "error": "query required"}
        t0 = time.monotonic()
        try:
            lim = max(1, min(limit, 20))
            if s.runtime is not None:
                data = json.loads(await s.runtime.call("search_local", {"query": q, "limit": lim, "format": "json"}))
                rows = [{"url": r.get("url", ""), "title": r.get("title", ""), "snippet": str(r.get("snippet", ""))[:300],
                         "score": round(float(r.get("score", r.get("combined_score", 0.0)) or 0.0), 4)} for r in data.get("results", [])]
                out = {"query": q, "total": data.get("total", len(rows)), "elapsed_ms": data.get("elapsed_ms", 0.0), "results": rows}
            else:
                from infomesh_b200.index.local_store import LocalStore
                from infomesh_b200.search.query import search_local

                with LocalStore(s.config.index.db_path, compression_enabled=s.config.storage.compression_enabled,
                                compression_level=s.config.storage.compression_level) as store:
                    res = search_local(store, q, limit=lim)
                out = {"query": q, "total": res.total, "elapsed_ms": round(res.elapsed_ms, 1),
                       "results": [{"url": r.url, "title": r.title, "snippet": r.snippet[:300], "score": round(r.combined_score, 4)} for r in res.results]}
            s.record_search((time.monotonic() - t0) * 1000)
            return out
        except Exception as exc:  # noqa: BLE001
            return {"results": [], "error": str(exc)[:200]}

    @app.get("/readiness")
    async def readiness(request: Request):
        ok = Path(st(request).config.index.db_path).exists()
        return JSONResponse(status_code=200 if ok else 503, content={"status": "ready" if ok else "not_ready", "db": "accessible" if ok else "missing"})

    @app.get("/status")
    async def status(request: Request) -> dict[str, Any]:
        s = st(request)
        up = time.time() - s.start_time
        return {"status": "running", "uptime_seconds": round(up, 1), "uptime_human": _format_duration(up), "index": _get_index_stats(s.config),
                "runtime": read_runtime_status(s.config.node.data_dir), "version": __version__}

    @app.get("/config")
    async def get_config(request: Request) -> dict[str, Any]:
        cfg_d = asdict(st(request).config)
        _redact_paths(cfg_d)
        return cfg_d

    @app.post("/config/reload", response_model=None)
    async def reload_config(request: Request):
        s = st(request)
        try:
            s.config = load_config(s.config_path)
            return {"status": "reloaded"}
        except Exception:  # noqa: BLE001
            logger.exception("config_reload_failed")
            return JSONResponse(status_code=500, content={"status": "error", "detail": "Failed to reload configuration"})

    @app.get("/index/stats")
    async def index_stats(request: Request) -> dict[str, Any]:
        return _get_index_stats(st(request).config)

    @app.get("/index/compression")
    async def index_compression(request: Request) -> dict[str, Any]:
        s = st(request)
        stats = _get_index_stats(s.config)
        n, mb = stats.get("document_count", 0), stats.get("db_size_mb", 0.0)
        return {"documents": n, "db_size_mb": mb, "avg_doc_kb": round(float(mb) * 1024 / n, 2) if isinstance(n, int) and n > 0 else 0.0,
                "compression_enabled": s.config.storage.compression_enabled, "compression_level": s.config.storage.compression_level}

    @app.get("/credits/balance")
    async def credits_balance(request: Request) -> dict[str, Any]:
        return _get_credit_stats(st(request).config)

    @app.get("/network/peers")
    async def network_peers(request: Request) -> dict[str, Any]:
        path = Path(st(request).config.node.data_dir) / "p2p_status.json"
        try:
            data = json.loads(path.read_text())
            if time.time() - float(data.get("timestamp", 0)) < 30:
                n = data.get("peers", data.get("connected_peers", 0))
                return {"total_peers": n, "connected": n, "peer_id": data.get("peer_id", ""), "state": data.get("state", ""),
                        "listen_addrs": data.get("listen_addrs", []), "dht": data.get("dht", {}), "bandwidth": data.get("bandwidth", {})}
        except (OSError, ValueError):
            pass
        return {"total_peers": 0, "connected": 0, "note": "P2P metrics available when node is networked"}

    @app.get("/analytics")
    async def analytics(request: Request) -> dict[str, Any]:
        s = st(request)
        return {"total_searches": s.total_searches, "total_crawls": s.total_crawls, "total_fetches": s.total_fetches,
                "avg_latency_ms": round(s.avg_latency_ms, 1), "uptime_seconds": round(time.time() - s.start_time, 1)}

    @app.get("/analytics/tools")
    async def tool_stats(request: Request) -> dict[str, Any]:
        s = st(request)
        usage = {"web_search": s.total_searches, "crawl_url": s.total_crawls, "fetch_page": s.total_fetches,
                 "total": s.total_searches + s.total_crawls + s.total_fetches}
        if s.runtime is not None:
            usage["mcp"] = dict(s.runtime.analytics.tool_calls)
        return {"tool_usage": usage, "search_fetch_rate": round(s.total_fetches / s.total_searches * 100, 1) if s.total_searches else 0.0}

    @app.get("/metrics")
    async def metrics(request: Request):
        from infomesh_b200.observability.metrics import MetricsCollector, get_collector

        s = st(request)
        mc = MetricsCollector()
        mc.inc("infomesh_search_total", float(s.total_searches))
        mc.inc("infomesh_crawl_total", float(s.total_crawls))
        mc.inc("infomesh_fetch_total", float(s.total_fetches))
        mc.set_gauge("infomesh_search_latency_ms_avg", s.avg_latency_ms)
        mc.set_gauge("infomesh_documents_indexed", float(_get_index_stats(s.config).get("document_count", 0) or 0))
        rt = read_runtime_status(s.config.node.data_dir)
        if isinstance(rt.get("process_memory_mb"), (int, float)):
            mc.set_gauge("infomesh_process_memory_mb", float(rt["process_memory_mb"]))
        text = mc.format_prometheus() + get_collector().format_prometheus()
        if "json" in request.headers.get("accept", ""):
            return JSONResponse(content={"metrics": text})
        return PlainTextResponse(text, media_type="text/plain; version=0.0.4")

    @app.get("/gpu/stats")
    async def gpu_stats(request: Request) -> dict[str, Any]:
        rt = st(request).runtime
        gi = getattr(rt, "gpu_index", None) if rt is not None else None
        return gi.stats() if gi is not None else {"enabled": False}

    @app.post("/index/submit", response_model=None)
    async def index_submit(request: Request):
        from infomesh_b200.p2p.protocol import MessageType, decode_message

        recv = st(request).index_submit_receiver
        if recv is None:
            return JSONResponse(status_code=404, content={"detail": "this node does not accept index submissions"})
        body = await request.body()
        try:
            kind, payload = decode_message(body)
        except ValueError as exc:
            return JSONResponse(status_code=400, content={"detail": str(exc)})
        if kind != MessageType.INDEX_SUBMIT:
            return JSONResponse(status_code=400, content={"detail": "expected INDEX_SUBMIT"})
        ack = recv.handle_submit(payload)
        return JSONResponse(status_code=200 if ack.success else 422, content={"url": ack.url, "doc_id": ack.doc_id, "success": ack.success, "error": ack.error})

    @app.get("/openapi-spec")
    async def openapi_spec() -> dict[str, Any]:
        from infomesh_b200.api.extensions import generate_openapi_spec

        return generate_openapi_spec()

    @app.get("/dashboard", response_class=HTMLResponse)
    async def dashboard_page():
        return HTMLResponse(content=_DASHBOARD_HTML)

    @app.get("/dashboard.js")
    async def dashboard_js():
        from starlette.responses import Response

        return Response(content=_DASHBOARD_JS, media_type="application/javascript")

    return app


async def serve_admin_api(config: Config, *, host: str = "127.0.0.1", port: int = 8080, runtime: Any | None = None,
                          index_submit_receiver: Any | None = None) -> None:
    import uvicorn

    app = create_admin_app(config, runtime=runtime, index_submit_receiver=index_submit_receiver)
    await uvicorn.Server(uvicorn.Config(app, host=host, port=port, log_level="warning", access_log=False)).serve()


_DASHBOARD_HTML = """<!doctype html><html><head><meta charset="utf-8"><title>InfoMesh node</title>
<style>body{font:14px system-ui;margin:2rem;background:#111;color:#ddd}h1{font-size:1.3rem}.c{display:inline-block;min-width:11rem;
margin:.4rem;padding:.8rem 1rem;background:#1c1c1c;border-radius:.5rem}.c b{display:block;font-size:1.4rem;color:#7fd}input{padding:.4rem;width:22rem}
li{margin:.5rem 0}a{color:#8bf}</style></head><body><h1>InfoMesh node</h1><div id="cards"></div>
<p><input id="q" placeholder="search the local index"> <button id="go">Search</button></p><ol id="res"></ol>
<script src="/dashboard.js"></script></body></html>"""

# Served as a same-origin file so the CSP can be `script-src 'self'` (no inline script, no inline event handlers).
# Everything that originates from crawled pages (title, snippet, url) is inserted with textContent / createElement only:
# the snippet's <b>…</b> highlight marks are re-created as elements, any other markup stays inert text, and links are
# kept only for http(s) URLs.
_DASHBOARD_JS = """'use strict';
async function j(u){return (await fetch(u)).json()}
function el(tag,text){const e=document.createElement(tag);if(text!==undefined)e.textContent=String(text);return e}
function safeUrl(u){try{const p=new URL(u);return (p.protocol==='http:'||p.protocol==='https:')?p.href:null}catch(e){return null}}
function snippetNodes(parent,s){const parts=String(s||'').split(/(<b>|<\\/b>)/);let bold=false;
for(const p of parts){if(p==='<b>'){bold=true;continue}if(p==='</b>'){bold=false;continue}if(!p)continue;
parent.appendChild(bold?el('b',p):document.createTextNode(p))}}
async function refresh(){const s=await j('/status'),a=await j('/analytics'),p=await j('/network/peers'),g=await j('/gpu/stats');
const cards=[['Documents',s.index.document_count],['DB size (MB)',s.index.db_size_mb],['Uptime',s.uptime_human],['Searches',a.total_searches],
['Avg latency (ms)',a.avg_latency_ms],['Peers',p.connected],['GPU docs',g.documents??'off']];
const box=document.getElementById('cards');box.replaceChildren();
for(const c of cards){const d=el('div');d.className='c';d.appendChild(document.createTextNode(c[0]));d.appendChild(el('b',c[1]));box.appendChild(d)}}
async function go(){const q=document.getElementById('q').value;const r=await j('/search?q='+encodeURIComponent(q)+'&limit=10');
const ol=document.getElementById('res');ol.replaceChildren();
for(const x of (r.results||[])){const li=el('li');const href=safeUrl(x.url);const head=href?el('a',x.title||x.url):el('span',x.title||x.url);
if(href){head.href=href;head.rel='noopener noreferrer'}li.appendChild(head);li.appendChild(document.createTextNode(' '));
li.appendChild(el('small',x.score));li.appendChild(el('br'));snippetNodes(li,x.snippet);ol.appendChild(li)}}
document.getElementById('go').addEventListener('click',go);
document.getElementById('q').addEventListener('keydown',e=>{if(e.key==='Enter')go()});
refresh();setInterval(refresh,5000);
"""
