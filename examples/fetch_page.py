"""fetch_page through the service layer: served from the local cache when fresh, crawled live otherwise."""
import asyncio
import sys
import tempfile
from dataclasses import replace
from pathlib import Path

from infomesh_b200.config import Config
from infomesh_b200.services import AppContext, fetch_page_async

url = sys.argv[1] if len(sys.argv) > 1 else "https://example.com/"


async def main() -> None:
    with tempfile.TemporaryDirectory() as d:
        base = Config()
        cfg = replace(base, node=replace(base.node, data_dir=Path(d)), index=replace(base.index, db_path=Path(d) / "index.db", vector_search=False))
        async with AppContext(cfg) as ctx:
            for attempt in ("live", "cached"):
                fp = await fetch_page_async(url, store=ctx.store, worker=ctx.worker)
                print(attempt, "->", "ok" if fp.success else fp.error, "| cached:", fp.is_cached, "|", fp.title, "|", len(fp.text), "chars")


asyncio.run(main())
