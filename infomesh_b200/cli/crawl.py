"""``infomesh crawl URL``, ``infomesh mcp [--http]``, ``infomesh dashboard [--tab] [--text]``,
``infomesh feeds import OPML | list`` (reference infomesh/cli/crawl.py:17-355)."""
from __future__ import annotations

import asyncio
import json

import click

from infomesh_b200.config import load_config


@click.command()
@click.argument("url")
@click.option("--depth", "-d", default=None, type=int, help="Follow links this many levels deep (same site)")
@click.option("--force", "-f", is_flag=True, help="Re-crawl even if the URL was crawled before")
def crawl(url: str, depth: int | None, force: bool) -> None:
    """Crawl one URL (and optionally its links), index it and print the outcome."""
    from infomesh_b200.security import SSRFError, validate_url
    from infomesh_b200.services import AppContext, crawl_and_index

    try:
        validate_url(url)
    except SSRFError as exc:
        raise click.ClickException(f"URL blocked: {exc}") from None
    cfg = load_config()
    want = depth if depth is not None else 0

    async def run() -> None:
        async with AppContext(cfg) as ctx:
            if ctx.worker is None:
                raise click.ClickException("this node role has no crawler (set node.role to full or crawler)")
            if want > 0:
                scoped = ctx.worker.set_scope(url)
                if asyncio.iscoroutine(scoped):
                    await scoped
            todo, seen, done = [(url, 0)], {url}, 0
            while todo:
                u, d = todo.pop(0)
                res = await ctx.worker.crawl_url(u, depth=d, force=force and d == 0)
                if res.success and res.page:
                    from infomesh_b200.services import index_document

                    if ctx.link_graph is not None and res.discovered_links:
                        ctx.link_graph.add_links(u, res.discovered_links)
                    index_document(res.page, ctx.store, ctx.vector_store, js_required=res.js_required)
                    done += 1
                    click.secho(f"  ✔ [{d}] {res.page.title[:60] or u}  ({len(res.page.text)} chars, {len(res.discovered_links)} links, {res.elapsed_ms:.0f} ms)", fg="green")
                    if d < want:
                        for link in res.discovered_links:
                            if link not in seen and len(seen) < 500:
                                seen.add(link)
                                todo.append((link, d + 1))
                else:
                    click.secho(f"  ✖ [{d}] {u}: {res.error}", fg="yellow")
            click.echo(f"Crawled {done} page(s).")

    del crawl_and_index
    asyncio.run(run())


@click.command("mcp")
@click.option("--http", is_flag=True, help="Serve streamable HTTP instead of stdio")
@click.option("--host", default="127.0.0.1", help="HTTP bind address")
@click.option("--port", default=8081, type=int, help="HTTP port")
def mcp_cmd(http: bool, host: str, port: int) -> None:
    """Run the MCP server (stdio by default) for an LLM client."""
    from infomesh_b200.mcp.server import run_mcp_http_server, run_mcp_server

    cfg = load_config()
    if http:
        click.echo(f"MCP server on http://{host}:{port}/mcp", err=True)
        asyncio.run(run_mcp_http_server(cfg, host=host, port=port))
    else:
        asyncio.run(run_mcp_server(cfg))


@click.command()
@click.option("--tab", "-t", default="overview", type=click.Choice(["overview", "crawl", "search", "network", "credits", "settings"]))
@click.option("--text", is_flag=True, help="Print a one-shot text report instead of the interactive TUI")
def dashboard(tab: str, text: bool) -> None:
    """Open the terminal dashboard."""
    cfg = load_config()
    if text:
        from infomesh_b200.dashboard.text_report import render_text_report

        click.echo(render_text_report(cfg))
        return
    from infomesh_b200.dashboard.app import run_dashboard

    run_dashboard(config=cfg, initial_tab=tab)


@click.group("feeds")
def feeds_group() -> None:
    """Manage RSS/Atom feed monitoring."""


def _feeds_file():
    return load_config().node.data_dir / "feeds.json"


@feeds_group.command("import")
@click.argument("opml_file", type=click.Path(exists=True))
def feeds_import(opml_file: str) -> None:
    """Import RSS/Atom feeds from an OPML file (the reference only lists them: cli/crawl.py:318-337; here they are also
    remembered in ``<data_dir>/feeds.json`` so that the feed monitor polls them)."""
    from pathlib import Path

    from infomesh_b200.crawler.feed_monitor import parse_opml

    feeds = parse_opml(Path(opml_file).read_text(encoding="utf-8"))
    if not feeds:
        click.echo("No feeds found in OPML file.")
        return
    click.echo(f"Found {len(feeds)} feeds:")
    for feed in feeds:
        click.echo(f"  {feed.url}{f' ({feed.label})' if feed.label else ''}")
    path = _feeds_file()
    known = set(json.loads(path.read_text())) if path.exists() else set()
    new = [f.url for f in feeds if f.url not in known]
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(sorted(known | set(new)), indent=2))
    click.secho(f"✔ {len(new)} new feeds ({len(known) + len(new)} total). Enable polling with: infomesh config set crawl.rss_enabled true", fg="green")


@feeds_group.command("list")
def feeds_list() -> None:
    """List currently configured RSS/Atom feeds."""
    crawl = load_config().crawl
    if crawl.rss_enabled:
        click.echo(f"RSS monitoring: enabled (interval={crawl.rss_default_interval}s, max={crawl.rss_max_feeds}, "
                   f"discovery={'on' if crawl.rss_discovery else 'off'})")
    else:
        click.echo("RSS feed monitoring is disabled. Enable with: infomesh config set crawl.rss_enabled true")
    path = _feeds_file()
    urls = json.loads(path.read_text()) if path.exists() else []
    if urls:
        click.echo(f"Subscribed feeds ({len(urls)}):")
        for u in urls:
            click.echo(f"  {u}")
