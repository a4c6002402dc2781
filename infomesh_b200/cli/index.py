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


@index_group.command("export")
@click.argument("output", default="infomesh-index.infomesh-snapshot")
def index_export(output: str) -> None:
    """Write a portable compressed snapshot of the index."""
    from infomesh_b200.index.snapshot import export_snapshot

    with _store(load_config()) as st:
        stats = export_snapshot(st, output)
    click.secho(f"✔ Exported {stats.total_documents} documents to {output} ({stats.file_size_bytes / 2 ** 20:.1f} MB, {stats.elapsed_ms:.0f} ms)", fg="green")


@index_group.command("import")
@click.argument("input_path", required=False, default=None)
@click.option("--starter", is_flag=True, help="Download and import the community starter snapshot")
@click.option("--info", "info_only", is_flag=True, help="With --starter: only show the remote snapshot's metadata")
def index_import(input_path: str | None, starter: bool, info_only: bool) -> None:
    """Import a snapshot file (or the starter snapshot)."""
    from infomesh_b200.index.snapshot import import_snapshot

    cfg = load_config()
    if starter:
        from infomesh_b200.index import starter as S

        info = asyncio.run(S.find_starter_asset(cache_dir=cfg.node.data_dir))
        if info is None:
            raise click.ClickException("no starter snapshot found (offline, or none has been published)")
        click.echo(f"Starter snapshot {info.release_tag}: {info.size_mb:.1f} MB, created {info.created_at}")
        if info_only:
            return
        bar = {"last": -1}

        def progress(done: int, total: int) -> None:
            pct = int(done * 100 / max(total, 1))
            if pct // 10 != bar["last"]:
                bar["last"] = pct // 10
                click.echo(f"  {pct}%")

        path = S.download_starter_sync(cfg.node.data_dir, progress_callback=progress)
        if path is None:
            raise click.ClickException("download failed")
        input_path = str(path)
    if not input_path:
        raise click.UsageError("give a snapshot path or --starter")
    with _store(cfg) as st:
        stats = import_snapshot(st, input_path)
    click.secho(f"✔ Imported {stats.exported} of {stats.total_documents} documents ({stats.skipped} skipped) in {stats.elapsed_ms:.0f} ms", fg="green")


@index_group.command("import-wet")
@click.argument("path_or_url")
def index_import_wet(path_or_url: str) -> None:
    """Import a Common Crawl WET file (local path or URL, .gz supported)."""
    from infomesh_b200.crawler.dedup import DeduplicatorDB
    from infomesh_b200.index.commoncrawl import CommonCrawlImporter

    cfg = load_config()
    with _store(cfg) as st:
        dedup = DeduplicatorDB(str(cfg.node.data_dir / "dedup.db"))
        try:
            s = asyncio.run(CommonCrawlImporter(st, dedup).import_wet_file(path_or_url))
        finally:
            dedup.close()
    click.secho(f"✔ {s.imported}/{s.total_records} records imported ({s.skipped_duplicate} duplicate, {s.skipped_too_short} short, "
                f"{s.skipped_error} errors) in {s.elapsed_ms:.0f} ms", fg="green")


@index_group.command("import-urls")
@click.argument("url_file")
@click.option("--max", "-m", "max_urls", default=10000, help="Maximum URLs to import")
def index_import_urls(url_file: str, max_urls: int) -> None:
    """Register the URLs of a text file for crawling."""
    from infomesh_b200.crawler.dedup import DeduplicatorDB
    from infomesh_b200.index.commoncrawl import CommonCrawlImporter

    cfg = load_config()
    with _store(cfg) as st:
        dedup = DeduplicatorDB(str(cfg.node.data_dir / "dedup.db"))
        try:
            s = asyncio.run(CommonCrawlImporter(st, dedup).import_url_list(url_file, max_urls=max_urls))
        finally:
            dedup.close()
    click.secho(f"✔ {s.imported} new URLs registered ({s.skipped_duplicate} already known)", fg="green")


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
