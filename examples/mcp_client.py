"""Call the MCP tools.  Default: in-process through ToolRuntime (what the server itself dispatches to).
With --stdio: spawn `python -m infomesh_b200 mcp` and speak MCP over its stdin/stdout."""
import asyncio
import json
import sys
import tempfile
from dataclasses import replace
from pathlib import Path

from infomesh_b200.config import Config


async def in_process() -> None:
    from infomesh_b200.mcp.handlers import ToolRuntime
    from infomesh_b200.sdk import InfoMeshClient
    from infomesh_b200.services import AppContext

    with tempfile.TemporaryDirectory() as d:
        with InfoMeshClient(d) as c:
            c.add_document("https://example.org/python", "History of Python", "Python was created by Guido van Rossum and first released in 1991. " * 4)
        base = Config()
        cfg = replace(base, node=replace(base.node, data_dir=Path(d)), index=replace(base.index, db_path=Path(d) / "index.db", vector_search=False))
        async with AppContext(cfg) as ctx:
            rt = ToolRuntime(ctx)
            print(json.loads(await rt.call("web_search", {"query": "python created", "top_k": 3}))["results"][0]["url"])
            print(json.loads(await rt.call("fact_check", {"claim": "Python was first released in 1991"}))["verdict"])
            print(json.loads(await rt.call("status", {}))["documents_indexed"], "documents")


async def over_stdio() -> None:
    from mcp import ClientSession, StdioServerParameters
    from mcp.client.stdio import stdio_client

    params = StdioServerParameters(command=sys.executable, args=["-m", "infomesh_b200", "mcp"])
    async with stdio_client(params) as (r, w), ClientSession(r, w) as session:
        await session.initialize()
        print([t.name for t in (await session.list_tools()).tools])
        print((await session.call_tool("status", {})).content[0].text[:300])


asyncio.run(over_stdio() if "--stdio" in sys.argv else in_process())
