"""``infomesh search QUERY [-n] [--local] [--vector] [--gpu]`` and ``infomesh feedback stats | top-urls``
(reference infomesh/cli/search.py:16-277)."""
from __future__ import annotations

import asyncio
import time

import click

from infomesh_b200.config import load_config


def _print_ranked(results, elapsed_ms: float, source: str) -> None:
    if not results:
        click.echo("No results found.")
        return
    click.secho(f"Found {len(results)} results ({elapsed_ms:.0f} ms, {source}):\n", bold=True)
    for i, r in enumerate(results, 1):
        get = (lambda k, d="": r.get(k, d)) if isinstance(r, dict) else (lambda k, d="": getattr(r, k, d))
        score = get("combined_score", None)
        score = get("score", 0.0) if score in (None, "") else score
        peer = get("peer_id", "")
        click.secho(f"{i}. {get('title') or get('url')}", fg="cyan")
        click.echo(f"   {get('url')}")
        click.echo(f"   score {float(score):.4f}" + (f"  · peer {str(peer)[:12]}" if peer else ""))
        snip = str(get("snippet")).replace("<b>", "").replace("</b>", "").replace("\n", " ")
        click.echo(f"   {snip[:220]}\n")


async def _wait_for_peer(node, timeout: float = 8.0) -> bool:
    deadline = time.monotonic() + timeout
    while time.monotonic() < deadline:
        if node.get_connected_peers():
            return True
        await asyncio.sleep(0.25)
    return False


@click.command()
@click.argument("query")
@click.option("--limit", "-n", default=10, type=click.IntRange(1, 100), help="Number of results")
@click.option("--local", "--local-only", "local_only", is_flag=True, help="Search the local index only")
@click.option("--vector", is_flag=True, help="Hybrid keyword + vector search")
@click.option("--gpu", is_flag=True, help="Use the fused GPU pipeline (builds the HBM index first)")
def search(query: str, limit: int, local_only: bool, vector: bool, gpu: bool) -> None:
    """Search the index (network search first, local fallback)."""
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.search import formatter as F      # the same renderers the MCP tools use (reference cli/search.py:33-136)
    from infomesh_b200.search import query as Q

    cfg = load_config()
    store = LocalStore(db_path=cfg.index.db_path, tokenizer=cfg.index.fts_tokenizer, compression_enabled=cfg.storage.compression_enabled,
                       compression_level=cfg.storage.compression_level)
    try:
        if gpu:
            from infomesh_b200.engine.multigpu import make_index

            gi = make_index(store, getattr(cfg, "gpu", None), query_batch=8)
            try:
                from infomesh_b200.engine.multigpu import warm_start

                warm_start(gi, store, getattr(getattr(cfg, "gpu", None), "segments_dir", ""))   # `infomesh search --gpu` cold-starts from files
                t0 = time.monotonic()
                hits = gi.search(query, limit)
                label = f"gpu hybrid x{gi.world}" if hasattr(gi, "world") else "gpu hybrid"
                _print_ranked(hits, (time.monotonic() - t0) * 1000, label)
            finally:
                gi.close()
            return
        if vector:
            from infomesh_b200.index.vector_store import VectorStore

            vs = VectorStore(persist_dir=cfg.node.data_dir / "vectors", model_name=cfg.index.embedding_model)
            try:
                res = Q.search_hybrid(store, vs, query, limit=limit)
            finally:
                vs.close()
            click.echo(F.format_hybrid_results(res))
            return
        if not local_only:
            from infomesh_b200.services import bootstrap_p2p, create_local_search_fn

            node, dist_index = bootstrap_p2p(cfg, local_search_fn=create_local_search_fn(cfg, store), enable_mdns=False)
            if node is not None:
                try:
                    async def go():
                        await _wait_for_peer(node)
                        return await Q.search_distributed(store, dist_index, query, limit=limit, network_search_fn=node.search_network)

                    res = asyncio.run(go())
                    click.echo(F.format_distributed_results(res))
                    return
                finally:
                    node.stop()
        click.echo(F.format_fts_results(Q.search_local(store, query, limit=limit)))
    finally:
        store.close()


@click.group("feedback")
def feedback_group() -> None:
    """Inspect implicit search quality signals."""


def _feedback_store():
    """The feedback database, or None while no search has recorded a signal yet (the file does not exist)."""
    from infomesh_b200.search.feedback import FeedbackStore

    path = load_config().node.data_dir / "feedback.db"
    return FeedbackStore(str(path)) if path.exists() else None


@feedback_group.command("stats")
def feedback_stats() -> None:
    """Show feedback signal statistics."""
    fb = _feedback_store()
    if fb is None:
        click.echo("No feedback data yet. Search more to collect signals.")
        return
    try:
        count, top = fb.signal_count(), fb.top_boosted_urls(5)
        click.echo(f"Total signals: {count}")
        click.echo(f"Boosted URLs:  {len(top)}")
        if top:
            click.echo(f"\n{'URL':<60} {'Boost':>8} {'Fetch':>6} {'Cite':>6}")
            click.echo("─" * 82)
            for u in top:
                shown = u.url if len(u.url) <= 60 else u.url[:58] + ".."
                click.echo(f"{shown:<60} {u.boost_score:>8.2f} {u.fetch_count:>6} {u.cite_count:>6}")
    finally:
        fb.close()


@feedback_group.command("top-urls")
@click.option("--limit", "-n", default=20, help="Number of URLs to show")
def feedback_top_urls(limit: int) -> None:
    """Show URLs with highest quality signals."""
    fb = _feedback_store()
    if fb is None:
        click.echo("No feedback data yet.")
        return
    try:
        top = fb.top_boosted_urls(limit)
        if not top:
            click.echo("No boosted URLs yet.")
        for u in top:
            click.echo(f"  {u.boost_score:+.2f}  {u.url}  (fetch={u.fetch_count} cite={u.cite_count})")
    finally:
        fb.close()
