#!/usr/bin/env python3
"""Crawl the curated seed lists into a throw-away node and export the result as ``starter.infomesh-snapshot`` — the
release asset that ``infomesh index import --starter`` downloads (counterpart of reference scripts/build_starter.py).

    python scripts/build_starter.py                                  # quickstart seeds, 500 pages
    python scripts/build_starter.py --category tech-docs --max-pages 2000 --depth 1
    python scripts/build_starter.py --seeds my_urls.txt --output out.infomesh-snapshot

The crawl runs through the same AppContext + crawl loop as a real node (robots, politeness, dedup, SimHash), so the
snapshot contains exactly what a node would have indexed itself."""
from __future__ import annotations

import argparse
import asyncio
import sys
import tempfile
import time
from dataclasses import replace
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


async def build(category: str, seed_file: Path | None, output: Path, max_pages: int, depth: int, max_minutes: float) -> int:
    from infomesh_b200.config import Config, NodeRole
    from infomesh_b200.crawler import crawl_loop
    from infomesh_b200.index.snapshot import export_snapshot
    from infomesh_b200.services import AppContext

    with tempfile.TemporaryDirectory(prefix="infomesh-starter-") as tmp:
        data = Path(tmp)
        base = Config()
        cfg = replace(base, node=replace(base.node, data_dir=data, role=NodeRole.FULL),
                      index=replace(base.index, db_path=data / "index.db", vector_search=False),
                      crawl=replace(base.crawl, max_depth=depth, urls_per_hour=0, politeness_delay=0.5, rss_enabled=False),
                      llm=replace(base.llm, enabled=False), gpu=replace(base.gpu, enabled=False))
        if seed_file is not None:                      # a private list replaces the packaged category files
            from infomesh_b200.crawler import seeds as S

            urls = S._parse_seed_file(seed_file)
            crawl_loop.load_seeds = lambda category=None: urls if category == "custom" else []   # type: ignore[assignment]
            category = "custom"
        t0 = time.monotonic()
        async with AppContext(cfg) as ctx:
            ctx.ledger = None                          # no credit accounting for a build job
            try:                                       # the loop idles (re-seeding) once the frontier is exhausted
                await asyncio.wait_for(crawl_loop.seed_and_crawl_loop(ctx, category, max_pages=max_pages), max_minutes * 60)
            except (TimeoutError, asyncio.TimeoutError):
                print(f"time budget of {max_minutes:g} min reached; exporting what has been crawled")
            stats = export_snapshot(ctx.store, output)
        print(f"crawled {stats.total_documents} pages in {time.monotonic() - t0:.0f} s; snapshot: {stats.total_documents} documents, "
              f"{output.stat().st_size / 2 ** 20:.1f} MB -> {output}")
        return stats.total_documents


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--category", default="quickstart", help="packaged seed category")
    ap.add_argument("--seeds", type=Path, default=None, help="text file with one URL per line (overrides --category)")
    ap.add_argument("--output", type=Path, default=Path("starter.infomesh-snapshot"))
    ap.add_argument("--max-pages", type=int, default=500)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--max-minutes", type=float, default=120.0, help="stop and export after this long even if --max-pages was not reached")
    a = ap.parse_args()
    docs = asyncio.run(build(a.category, a.seeds, a.output, a.max_pages, a.depth, a.max_minutes))
    return 0 if docs > 0 else 1


if __name__ == "__main__":
    raise SystemExit(main())
