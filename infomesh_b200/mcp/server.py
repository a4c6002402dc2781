"""MCP server wiring: stdio transport (``infomesh mcp``) and streamable HTTP (``infomesh mcp --http``) with /mcp,
/health and CORS; every call is API-key checked when ``INFOMESH_API_KEY`` is set and dispatched through
:class:`ToolRuntime` (reference infomesh/mcp/server.py:87-672)."""
from __future__ import annotations

import asyncio
import contextlib
import json
import os
from typing import Any

from infomesh_b200.config import Config, load_config
from infomesh_b200.mcp.handlers import ToolRuntime
from infomesh_b200.mcp.tools import check_api_key, get_all_tools
from infomesh_b200.services import AppContext, republish_local_index
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_CORS_HEADERS = [[b"access-control-allow-origin", b"*"], [b"access-control-allow-methods", b"GET, POST, DELETE, OPTIONS"],
                 [b"access-control-allow-headers", b"content-type, mcp-session-id, x-api-key, authorization"],
                 [b"access-control-expose-headers", b"mcp-session-id"]]


def _env_api_key() -> str | None:
    return os.environ.get("INFOMESH_API_KEY") or None


def create_runtime(config: Config, *, distributed_index: Any | None = None, p2p_node: Any | None = None,
                   ctx: AppContext | None = None) -> tuple[ToolRuntime, AppContext, Any]:
    """-> (runtime, AppContext, PersistentStore).  Also attaches the GPU index when ``[gpu] enabled`` and a device
    is visible."""
    from infomesh_b200.persistence.store import PersistentStore

    ctx = ctx or AppContext(config, apply_os_priority=True)
    pstore = PersistentStore(str(config.node.data_dir / "persistent.db"))
    attach_gpu_index(ctx)
    rt = ToolRuntime(ctx, distributed_index=distributed_index, p2p_node=p2p_node, pstore=pstore)
    for url in pstore.get_webhooks():
        rt.webhooks.register(url)
    return rt, ctx, pstore


def attach_gpu_index(ctx: AppContext) -> Any | None:
    """Build the HBM mirror of the local store when the GPU plane is enabled and usable; never fatal."""
    if getattr(ctx, "gpu_index", None) is not None:
        return ctx.gpu_index
    ctx.gpu_index = None
    gcfg = getattr(ctx.config, "gpu", None)
    if gcfg is None or not getattr(gcfg, "enabled", False):
        return None
    try:
        import torch

        from infomesh_b200 import _native

        if not torch.cuda.is_available() or not _native.available():
            return None
        from infomesh_b200.engine.multigpu import make_index

        # [gpu] devices > 1 (or 0 = all visible): one worker process per GPU behind the same interface
        gi = make_index(ctx.store, gcfg, rerank=getattr(gcfg, "rerank", True), query_batch=getattr(gcfg, "query_batch", 64))
        from infomesh_b200.engine.multigpu import warm_start

        warm_start(gi, ctx.store, getattr(gcfg, "segments_dir", ""))
        ctx.gpu_index = gi
    except Exception as exc:  # noqa: BLE001
        logger.warning("gpu_index_unavailable", error=str(exc))
    return ctx.gpu_index


def _create_app(config: Config, distributed_index: Any | None = None, p2p_node: Any | None = None, *, api_key: str | None = None):
    from mcp.server import Server
    from mcp.types import TextContent, Tool

    app = Server("infomesh")
    rt, ctx, pstore = create_runtime(config, distributed_index=distributed_index, p2p_node=p2p_node)

    @app.list_tools()
    async def list_tools() -> list[Tool]:
        return get_all_tools(api_key_required=api_key is not None)

    @app.call_tool()
    async def call_tool(name: str, arguments: dict[str, Any]) -> list[TextContent]:
        err = check_api_key(arguments or {}, api_key)
        if err is not None:
            return [TextContent(type="text", text=err)]
        return [TextContent(type="text", text=await rt.call(name, arguments))]

    return app, ctx, pstore


async def _serve(app: Any, ctx: AppContext, pstore: Any, run, *, distributed_index, p2p_node) -> None:
    republish = asyncio.create_task(republish_local_index(ctx.store, p2p_node=p2p_node, distributed_index=distributed_index))
    try:
        await run()
    finally:
        republish.cancel()
        with contextlib.suppress(asyncio.CancelledError, Exception):
            await republish
        pstore.close()


async def run_mcp_server(config: Config | None = None, *, distributed_index: Any | None = None, p2p_node: Any | None = None) -> None:
    from mcp.server.stdio import stdio_server

    config = config or load_config()
    app, ctx, pstore = _create_app(config, distributed_index, p2p_node, api_key=_env_api_key())
    async with ctx, stdio_server() as (rs, ws):
        await _serve(app, ctx, pstore, lambda: app.run(rs, ws, app.create_initialization_options()),
                     distributed_index=distributed_index, p2p_node=p2p_node)


def make_asgi_app(transport: Any):
    """CORS pre-flight, ``/health`` and ``/mcp`` routing around a streamable-HTTP transport."""
    async def respond(send, status: int, body: bytes = b"", ctype: bytes | None = None) -> None:
        headers = list(_CORS_HEADERS) + ([[b"content-type", ctype]] if ctype else [])
        await send({"type": "http.response.start", "status": status, "headers": headers})
        await send({"type": "http.response.body", "body": body})

    async def asgi(scope, receive, send) -> None:
        if scope.get("type") != "http":
            return
        method, path = scope.get("method", ""), scope.get("path", "")
        if method == "OPTIONS":
            await respond(send, 204)
        elif path == "/health":
            await respond(send, 200, json.dumps({"status": "ok"}).encode(), b"application/json")
        elif path == "/mcp":
            await transport.handle_request(scope, receive, send)
        else:
            await respond(send, 404, b"Not Found")

    return asgi


async def run_mcp_http_server(config: Config | None = None, *, host: str = "127.0.0.1", port: int = 8081,
                              distributed_index: Any | None = None, p2p_node: Any | None = None) -> None:
    import uvicorn
    from mcp.server.streamable_http import StreamableHTTPServerTransport

    config = config or load_config()
    app, ctx, pstore = _create_app(config, distributed_index, p2p_node, api_key=_env_api_key())
    transport = StreamableHTTPServerTransport(mcp_session_id=None)
    server = uvicorn.Server(uvicorn.Config(make_asgi_app(transport), host=host, port=port, log_level="info"))
    logger.info("mcp_server_starting", transport="http", host=host, port=port)
    async with ctx, transport.connect() as (rs, ws):
        mcp_task = asyncio.create_task(app.run(rs, ws, app.create_initialization_options()))

        async def run():
            try:
                await server.serve()
            finally:
                mcp_task.cancel()
                with contextlib.suppress(asyncio.CancelledError):
                    await mcp_task

        await _serve(app, ctx, pstore, run, distributed_index=distributed_index, p2p_node=p2p_node)
