"""``infomesh index stats | export | import [--starter] | import-wet | import-urls | gpu-build``
(reference infomesh/cli/index.py:13-282; ``gpu-build`` is new: it builds the HBM-resident index once and reports its
footprint and build time)."""
from __future__ import annotations

import asyncio

import click

from infomesh_b200.config import load_config


def _store(cfg):
    from infomesh_b200.index.local_store import LocalStore

    return LocalStore(db_path=cfg.index.db_path, tokenizer=cfg.index.fts_tokenizer, compression_enabled=cfg.storage.compression_enabled,
                      compression_level=cfg.storage.compression_level)


@click.group("index")
def index_group() -> None:
    """Inspect, export and import the local index."""


@index_group.command("stats")
def index_stats() -> None:
    """Document count, database size and top domains."""
    cfg = load_config()
    with _store(cfg) as st:
        n = st.get_stats().get("document_count", 0)
        click.echo(f"Index Statistics\n{'=' * 30}")                       # field labels as in reference cli/index.py:32-44
        click.echo(f"Database:        {cfg.index.db_path}")
        click.echo(f"Documents:       {n}")
        click.echo(f"Tokenizer:       {cfg.index.fts_tokenizer}")
        click.echo(f"Compression:     {'on' if cfg.storage.compression_enabled else 'off'}")
        if cfg.index.db_path.exists():
            click.echo(f"DB size:         {cfg.index.db_path.stat().st_size / 2 ** 20:.2f} MB")
        top = st.get_top_domains(limit=7)
        if top:
            click.echo("Top domains:")
            for dom, cnt in top:
                click.echo(f"  {cnt:6d}  {dom}")


def _report(headline: str, details: list[tuple[str, object]], elapsed_ms: float | None = None) -> None:
    """A headline followed by indented ``label value`` lines and the elapsed time: the layout every import / export prints."""
    click.echo(headline)
    for label, value in details:
        click.echo(f"  {label} {value}")
    if elapsed_ms is not None:
        click.echo(f"  Time: {elapsed_ms:.0f}ms")


@index_group.command("export")
@click.argument("output", default="infomesh-index.infomesh-snapshot")
def index_export(output: str) -> None:
    """Export the local index to a snapshot file."""
    from infomesh_b200.index.snapshot import export_snapshot

    with _store(load_config()) as st:
        stats = export_snapshot(st, output)
    click.echo(f"Exported {stats.total_documents} documents to {output}")
    click.echo(f"  File size: {stats.file_size_bytes / 2 ** 20:.2f} MB ({stats.elapsed_ms:.0f}ms)")


def _import_starter(cfg, info_only: bool) -> None:
    from infomesh_b200.index import starter as S
    from infomesh_b200.index.snapshot import import_snapshot

    click.echo("  ⏳ Checking for starter index on GitHub Releases...")
    asset = asyncio.run(S.find_starter_asset(cache_dir=cfg.node.data_dir))
    if asset is None:
        click.secho("  No starter snapshot found in GitHub Releases.", fg="yellow")
        click.echo("  The project maintainer has not published a starter index yet.")
        raise SystemExit(0)
    if info_only:
        _report("Starter Index (remote)", [("Release: ", asset.release_tag), ("Size:    ", f"{asset.size_mb:.1f} MB"), ("Created: ", asset.created_at),
                                           ("URL:     ", asset.download_url)])
        return
    click.echo(f"  Found starter index: {asset.size_mb:.1f} MB (release {asset.release_tag})")
    import sys

    if sys.stdin.isatty() and not click.confirm("  Download and import it?", default=True):
        return
    click.echo(f"  ⏳ Downloading {asset.size_mb:.1f} MB...")
    shown = {"pct": -1}

    def progress(done: int, total: int) -> None:
        pct = int(done * 100 / max(total, 1)) // 10 * 10
        if pct != shown["pct"]:
            shown["pct"] = pct
            click.echo(f"  ... {pct}%")

    path = S.download_starter_sync(cfg.node.data_dir, progress_callback=progress)
    if path is None:
        click.secho("  Download failed.", fg="red")
        raise SystemExit(1)
    click.echo(f"  ✔ Downloaded to {path}")
    click.echo("  ⏳ Importing into local index...")
    with _store(cfg) as st:
        stats = import_snapshot(st, path)
    _report(f"  ✔ Imported {stats.exported} documents", [("Skipped (duplicate):", stats.skipped), ("Total in snapshot:  ", stats.total_documents)], stats.elapsed_ms)
    click.secho("\n  Starter index loaded! Run 'infomesh search <query>' to try it out.", fg="green", bold=True)


@index_group.command("import")
@click.argument("input_path", required=False, default=None)
@click.option("--info", "info_only", is_flag=True, default=False, help="Show snapshot metadata only")
@click.option("--starter", is_flag=True, default=False, help="Download and import the community starter index from GitHub Releases")
def index_import(input_path: str | None, info_only: bool, starter: bool) -> None:
    """Import a snapshot file into the local index.

    \b
    With --starter, downloads the community starter index from GitHub
    Releases automatically (no INPUT_PATH needed). Use --starter --info
    to check what's available without downloading."""
    from infomesh_b200.index.snapshot import import_snapshot, read_snapshot_metadata

    cfg = load_config()
    if starter:
        _import_starter(cfg, info_only)
        return
    if input_path is None:
        click.echo("Error: Missing argument 'INPUT_PATH'.")
        click.echo("  Use --starter to download the community index.")
        raise SystemExit(1)
    if info_only:
        import datetime

        meta = read_snapshot_metadata(input_path)
        created = datetime.datetime.fromtimestamp(meta.get("created_at", 0), tz=datetime.UTC).isoformat()
        _report(f"Snapshot: {input_path}", [("Format version:", meta.get("format_version")), ("Documents:     ", meta.get("document_count")), ("Created:       ", created)])
        return
    with _store(cfg) as st:
        stats = import_snapshot(st, input_path)
    _report(f"Imported {stats.exported} documents from {input_path}", [("Skipped (duplicate):", stats.skipped), ("Total in snapshot:  ", stats.total_documents)],
            stats.elapsed_ms)


def _with_importer(run):
    """Open the store + dedup database, hand a ``CommonCrawlImporter`` to ``run`` (a coroutine function), close both."""
    from infomesh_b200.crawler.dedup import DeduplicatorDB
    from infomesh_b200.index.commoncrawl import CommonCrawlImporter

    cfg = load_config()
    with _store(cfg) as st:
        dedup = DeduplicatorDB(str(cfg.node.data_dir / "dedup.db"))
        try:
            return asyncio.run(run(CommonCrawlImporter(st, dedup)))
        finally:
            dedup.close()


@index_group.command("import-wet")
@click.argument("path_or_url")
def index_import_wet(path_or_url: str) -> None:
    """Import a Common Crawl WET file (local path or URL, .gz supported)."""
    s = _with_importer(lambda imp: imp.import_wet_file(path_or_url))
    _report(f"Imported {s.imported} documents from WET file", [("Total records:   ", s.total_records), ("Skipped (dup):   ", s.skipped_duplicate),
                                                                ("Skipped (short): ", s.skipped_too_short), ("Skipped (error): ", s.skipped_error)], s.elapsed_ms)


@index_group.command("import-urls")
@click.argument("url_file")
@click.option("--max", "-m", "max_urls", default=10000, help="Maximum URLs to import")
def index_import_urls(url_file: str, max_urls: int) -> None:
    """Register the URLs of a text file for crawling."""
    s = _with_importer(lambda imp: imp.import_url_list(url_file, max_urls=max_urls))
    _report(f"Registered {s.imported} URLs from {url_file}", [("Skipped (already seen):", s.skipped_duplicate)])


@index_group.command("gpu-build")
@click.option("--no-rerank", is_flag=True, help="Skip loading the cross-encoder")
@click.option("--save", "save_dir", default=None, help="Also write the device segments + manifest to this directory")
def index_gpu_build(no_rerank: bool, save_dir: str | None) -> None:
    """Build the HBM-resident mirror of the index (one GPU, or ``[gpu] devices`` of them) and report its footprint."""
    from infomesh_b200.engine.multigpu import make_index

    cfg = load_config()
    gcfg = getattr(cfg, "gpu", None)
    save_dir = save_dir or getattr(gcfg, "segments_dir", "") or None
    with _store(cfg) as st:
        gi = make_index(st, gcfg, rerank=not no_rerank)
        try:
            n = gi.rebuild()
            info = gi.stats()
            if save_dir and n:
                man = gi.save(save_dir)
                parts = man["shards"] if "shards" in man else [{"bytes": sum(f["bytes"] for f in man["files"].values())}]
                click.echo(f"  segments written to {save_dir} ({sum(p_['bytes'] for p_ in parts) / 2 ** 20:.1f} MB, {len(parts)} shard(s))")
        finally:
            gi.close()
    where = f"{info['gpus']} GPUs" if "gpus" in info else "1 GPU"
    click.secho(f"✔ {n} documents resident on {where}: {info['hbm_bytes'] / 2 ** 20:.1f} MB HBM, built in {info['build_seconds']} s", fg="green")
