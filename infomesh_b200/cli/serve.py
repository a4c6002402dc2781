"""Node lifecycle: ``infomesh start [-s SEEDS] [-b|--no-dashboard] [-r ROLE]``, ``stop``, ``update [--check]``,
``status`` and the hidden worker entry ``_serve [--seeds] [--role] [--no-crawl]``
(reference infomesh/cli/serve.py:36-955).

``_serve`` is the long-running process: rotating log file, PID file under the start-up lock, credit sync, the P2P node
on its own thread, then one asyncio loop running the crawl loop, the runtime-status heartbeat, the admin HTTP API and —
when ``[gpu] enabled`` — the HBM-resident index, until SIGTERM/SIGINT."""
from __future__ import annotations

import asyncio
import contextlib
import logging
import os
import signal
import subprocess
import sys
import time
from dataclasses import replace
from logging.handlers import RotatingFileHandler

import click

from infomesh_b200 import __version__
from infomesh_b200 import runtime as RT
from infomesh_b200.config import Config, NodeRole, load_config

ADMIN_API_PORT = 8080


def _serve_cmd(seeds: str | None, role: str | None, no_crawl: bool = False) -> list[str]:
    cmd = [sys.executable, "-m", "infomesh_b200", "_serve"]
    if seeds:
        cmd += ["--seeds", seeds]
    if role:
        cmd += ["--role", role]
    if no_crawl:
        cmd.append("--no-crawl")
    return cmd


def _spawn(cmd: list[str]) -> subprocess.Popen:
    return subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, stdin=subprocess.DEVNULL, start_new_session=True)


def _offer_starter_download(config: Config) -> None:
    from infomesh_b200.index import starter as S

    if not sys.stdin.isatty():
        click.echo("  ℹ Index is empty. Seed it with: infomesh index import --starter")
        return
    info = asyncio.run(S.find_starter_asset(cache_dir=config.node.data_dir))
    if info is None:
        return
    if click.confirm(f"  Index is empty. Download the starter snapshot ({info.size_mb:.0f} MB, {info.release_tag})?", default=True):
        path = S.download_starter_sync(config.node.data_dir)
        if path is None:
            click.secho("  ✖ download failed", fg="yellow")
            return
        from infomesh_b200.index.local_store import LocalStore
        from infomesh_b200.index.snapshot import import_snapshot

        with LocalStore(db_path=config.index.db_path, compression_enabled=config.storage.compression_enabled,
                        compression_level=config.storage.compression_level) as st:
            stats = import_snapshot(st, path)
        click.secho(f"  ✔ imported {stats.exported} documents", fg="green")


@click.command()
@click.option("--seeds", "-s", default=None, help="Seed category (tech-docs, academic, encyclopedia, quickstart, search-strategy)")
@click.option("--background", "-b", is_flag=True, help="Start in the background and return")
@click.option("--no-dashboard", is_flag=True, help="Same as --background")
@click.option("--role", "-r", default=None, type=click.Choice(["full", "crawler", "search"]), help="Override node.role")
def start(seeds: str | None, background: bool, no_dashboard: bool, role: str | None) -> None:
    """Start the node (foreground: initial crawl pass + dashboard; background: detach)."""
    background = background or no_dashboard
    config = load_config()
    if role:
        config = replace(config, node=replace(config.node, role=role))
    lock = RT.StartupLock(config.node.data_dir)
    if not lock.acquire():
        click.secho("Another InfoMesh startup is already in progress.", fg="yellow")
        return
    try:
        running = RT.read_live_pid(config.node.data_dir)
        if running is not None:
            click.secho(f"InfoMesh node is already running (PID {running}).", fg="yellow")
            click.echo("  Stop node:   infomesh stop\n  Check status: infomesh status")
            return
        from infomesh_b200.p2p.keys import ensure_keys

        keys = ensure_keys(config.node.data_dir)
        click.echo(f"InfoMesh v{__version__} starting...\n  Peer ID: {keys.peer_id}\n  Data dir: {config.node.data_dir}")
        from infomesh_b200.version_check import check_pypi_update, format_update_banner

        upd = check_pypi_update(config.node.data_dir)
        if upd is not None:
            click.secho(format_update_banner(upd), fg="yellow", bold=True)
        from infomesh_b200.credits.github_identity import run_first_start_checks

        run_first_start_checks(config, interactive=sys.stdin.isatty(), echo=click.echo)
        from infomesh_b200.resources.preflight import IssueSeverity, run_preflight_checks

        click.echo("  ⏳ Running preflight checks...", nl=False)
        issues = run_preflight_checks(config.node.data_dir, gpu=config.gpu.enabled)
        fatal = [i for i in issues if i.severity == IssueSeverity.ERROR and i.check != "network"]
        click.echo(" ✖" if fatal else " ✔")
        for i in issues:
            err = i.severity == IssueSeverity.ERROR
            click.secho(f"  {'✖' if err else '⚠'} [{i.check}] {i.message}", fg="red" if err else "yellow")
        if fatal:
            click.secho("\nCannot start: fix the errors above first.", fg="red", bold=True)
            raise SystemExit(1)
        from infomesh_b200.index.local_store import LocalStore
        from infomesh_b200.index.starter import needs_starter

        with LocalStore(db_path=config.index.db_path, compression_enabled=config.storage.compression_enabled,
                        compression_level=config.storage.compression_level) as st:
            if needs_starter(int(st.get_stats().get("document_count", 0))):
                _offer_starter_download(config)
        from infomesh_b200.resources.port_check import check_port_and_offer_fix

        port_ok = check_port_and_offer_fix(config.node.listen_port)
        if not port_ok and sys.stdin.isatty() and not click.confirm(
                "  Continue without P2P port access? (local crawl & search will work, but peering won't)", default=False):
            raise SystemExit(1)
        click.echo("  ⏳ Launching node process...", nl=False)
        proc = _spawn(_serve_cmd(seeds, role))
        RT.write_pid_file(config.node.data_dir, proc.pid)
        log_path = config.node.data_dir / "node.log"
        click.echo(f" ✔ (PID {proc.pid})\n  Log: {log_path}")
    finally:
        lock.release()
    if background:
        click.echo(f"\n  Node running in background (no live log).\n  View logs:   tail -f {log_path}\n  Stop node:   infomesh stop")
        return
    click.echo("\n  Launching dashboard...")
    from infomesh_b200.dashboard.app import run_dashboard

    action = run_dashboard(config=config, node_pid=proc.pid)
    if action == "stop_all":
        with contextlib.suppress(ProcessLookupError):
            if RT.request_graceful_stop(proc.pid, timeout_seconds=10.0):
                click.echo(f"\nInfoMesh node stopped (PID {proc.pid}).")
            else:
                click.secho(f"\nInfoMesh node did not exit within timeout (PID {proc.pid}).", fg="yellow")
                return
        RT.clear_pid_file(config.node.data_dir, proc.pid)
    else:
        click.echo(f"\nDashboard closed. Node still running (PID {proc.pid}).\n  Use 'infomesh stop' to stop the node.")


@click.command()
def stop() -> None:
    """Stop the running node."""
    config = load_config()
    pid = RT.read_live_pid(config.node.data_dir)
    if pid is None:
        click.echo("No running InfoMesh node found.")
        return
    try:
        if RT.request_graceful_stop(pid, timeout_seconds=10.0):
            click.echo(f"InfoMesh node stopped (PID {pid}).")
            RT.clear_pid_file(config.node.data_dir, pid)
            RT.mark_runtime_stopped(config.node.data_dir, pid)
        else:
            click.secho(f"InfoMesh node did not exit within timeout (PID {pid}).", fg="yellow")
    except ProcessLookupError:
        click.echo("Node process not found (stale PID file). Cleaning up.")
        RT.clear_pid_file(config.node.data_dir, pid)


@click.command()
@click.option("--check", is_flag=True, help="Only check for updates without installing")
def update(check: bool) -> None:
    """Check for and install updates, restarting a running node."""
    import shutil

    from infomesh_b200.version_check import check_pypi_update, format_update_banner

    config = load_config()
    click.echo(f"Current version: v{__version__}")
    info = check_pypi_update(config.node.data_dir)
    if info is None:
        click.secho("✔ Already up to date.", fg="green")
        return
    click.secho(format_update_banner(info), fg="yellow", bold=True)
    if check:
        return
    uv = shutil.which("uv")
    from infomesh_b200 import DISTRIBUTION

    cmd = [uv, "pip", "install", "--upgrade", DISTRIBUTION] if uv else [sys.executable, "-m", "pip", "install", "--upgrade", DISTRIBUTION]
    click.echo(f"  Running: {' '.join(cmd)}")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        click.secho("  ✖ Upgrade failed:", fg="red")
        click.echo(res.stderr or res.stdout)
        raise SystemExit(1)
    click.secho(f"  ✔ Upgraded to v{info.latest}", fg="green")
    pid = RT.read_live_pid(config.node.data_dir)
    if pid is None:
        click.echo("  Node is not running (start with: infomesh start)")
        return
    try:
        click.echo(f"  ↻ Restarting node (sent SIGTERM to PID {pid})...")
        if not RT.request_graceful_stop(pid, timeout_seconds=10.0):
            click.secho(f"  ✖ Node did not exit within timeout (PID {pid}).", fg="red")
            raise SystemExit(1)
    except ProcessLookupError:
        pass
    RT.clear_pid_file(config.node.data_dir, pid)
    RT.write_pid_file(config.node.data_dir, _spawn(_serve_cmd(None, None)).pid)
    click.secho("  ✔ Node restarted with new version.", fg="green")


def _setup_file_logging(config: Config) -> None:
    handler = RotatingFileHandler(str(config.node.data_dir / "node.log"), maxBytes=10 * 2 ** 20, backupCount=5, encoding="utf-8")
    handler.setLevel(logging.DEBUG)
    logging.basicConfig(handlers=[handler], level=getattr(logging, config.node.log_level.upper(), logging.INFO), format="%(asctime)s %(message)s", force=True)


def _make_credit_sync(config: Config):
    try:
        from infomesh_b200.credits.github_identity import resolve_github_email
        from infomesh_b200.credits.ledger import CreditLedger
        from infomesh_b200.credits.sync import CreditSyncManager, CreditSyncStore
        from infomesh_b200.p2p.keys import ensure_keys

        email = resolve_github_email(config) or ""
        if not email:
            return None
        kp = ensure_keys(config.node.data_dir)
        return CreditSyncManager(CreditLedger(config.node.data_dir / "credits.db", owner_email=email),
                                 CreditSyncStore(config.node.data_dir / "credit_sync.db"), email, key_pair=kp, local_peer_id=kp.peer_id)
    except Exception:  # noqa: BLE001
        logging.getLogger(__name__).warning("credit_sync_init_failed", exc_info=True)
        return None


@click.command(name="_serve", hidden=True)
@click.option("--seeds", "-s", default=None)
@click.option("--role", "-r", default=None, type=click.Choice(["full", "crawler", "search"]))
@click.option("--no-crawl", is_flag=True, help="Serve P2P / API only")
def serve(seeds: str | None, role: str | None, no_crawl: bool) -> None:
    """Internal: the node worker process."""
    config = load_config()
    if role:
        config = replace(config, node=replace(config.node, role=role))
    lock = RT.StartupLock(config.node.data_dir)
    if not lock.acquire():
        click.echo("Another InfoMesh startup is already in progress.")
        raise SystemExit(1)
    running = RT.read_live_pid(config.node.data_dir)
    if running is not None and running != os.getpid():
        lock.release()
        click.echo(f"InfoMesh node already running (PID {running}).")
        raise SystemExit(1)
    _setup_file_logging(config)
    log = logging.getLogger("infomesh.serve")
    RT.write_pid_file(config.node.data_dir, os.getpid())
    lock.release()

    from infomesh_b200.services import AppContext, bootstrap_p2p, create_local_search_fn, republish_local_index

    credit_sync = _make_credit_sync(config)
    ctx = AppContext(config, apply_os_priority=True)

    async def store_replica(*, url: str, title: str, text: str, text_hash: str, language: str) -> bool:
        from infomesh_b200.crawler.parser import ParsedPage
        from infomesh_b200.hashing import content_hash
        from infomesh_b200.services import index_document

        page = ParsedPage(url=url, title=title, text=text, language=language or None, raw_html_hash=content_hash(url), text_hash=text_hash or content_hash(text))
        return index_document(page, ctx.store, ctx.vector_store) is not None

    node, dist_index = bootstrap_p2p(config, credit_sync_manager=credit_sync or ctx.credit_sync_manager,
                                     local_search_fn=create_local_search_fn(config, ctx.store), store_fn=store_replica,
                                     index_submit_receiver=ctx.index_submit_receiver)
    ctx.p2p_node, ctx.distributed_index = node, dist_index

    async def main() -> None:
        from infomesh_b200.api.local_api import serve_admin_api
        from infomesh_b200.crawler.crawl_loop import seed_and_crawl_loop
        from infomesh_b200.mcp.handlers import ToolRuntime
        from infomesh_b200.mcp.server import attach_gpu_index

        started = time.time()
        stop_event = asyncio.Event()
        loop = asyncio.get_running_loop()
        for sig in (signal.SIGTERM, signal.SIGINT):
            with contextlib.suppress(NotImplementedError, RuntimeError):
                loop.add_signal_handler(sig, stop_event.set)
        await asyncio.to_thread(attach_gpu_index, ctx)
        runtime = ToolRuntime(ctx, distributed_index=dist_index, p2p_node=node)

        async def heartbeat() -> None:
            while True:
                gi = getattr(ctx, "gpu_index", None)
                RT.write_runtime_status(config.node.data_dir, RT.build_runtime_status(
                    pid=os.getpid(), role=str(config.node.role), started_at=started, no_crawl=no_crawl,
                    governor_state=ctx.governor.check_and_adjust(), gpu=gi.stats() if gi is not None else None))
                await asyncio.sleep(10)

        tasks = [asyncio.create_task(heartbeat()),
                 asyncio.create_task(serve_admin_api(config, port=ADMIN_API_PORT, runtime=runtime, index_submit_receiver=ctx.index_submit_receiver))]
        if node is not None or dist_index is not None:
            tasks.append(asyncio.create_task(republish_local_index(ctx.store, p2p_node=node, distributed_index=dist_index)))
        if str(config.node.role).lower() != NodeRole.SEARCH and not no_crawl:
            tasks.append(asyncio.create_task(seed_and_crawl_loop(ctx, seed_category=seeds or "tech-docs")))
        else:
            log.info("waiting_mode: P2P active, crawl loop disabled")
        try:
            await stop_event.wait()
            log.info("shutdown requested")
        finally:
            for t in tasks:
                t.cancel()
            for t in tasks:
                with contextlib.suppress(asyncio.CancelledError, Exception):
                    await t
            await ctx.close_async()

    try:
        asyncio.run(main())
    except KeyboardInterrupt:
        pass
    finally:
        if node is not None:
            with contextlib.suppress(Exception):
                node.stop()
        RT.clear_pid_file(config.node.data_dir, os.getpid())
        RT.mark_runtime_stopped(config.node.data_dir, os.getpid())


def _render_p2p_status(config: Config, running: bool) -> None:
    import json

    try:
        st = json.loads((config.node.data_dir / "p2p_status.json").read_text())
    except (OSError, ValueError):
        click.echo("P2P:             stopped")
        return
    fresh = time.time() - float(st.get("timestamp", 0)) < 30
    state = st.get("state", "unknown") if (fresh and running) else "stopped"
    click.echo(f"P2P:             {state}, {st.get('peers', 0) if fresh else 0} peer(s)")
    for a in st.get("listen_addrs", [])[:2]:
        click.echo(f"  listen:        {a}")
    boot = st.get("bootstrap") or {}
    if boot:
        click.echo(f"  bootstrap:     {boot.get('connected', 0)} connected / {boot.get('configured', 0)} configured")
    if fresh and running and not st.get("peers"):
        click.echo("  hint:          add a peer with `infomesh peer add /ip4/HOST/tcp/4001`")


def _render_credit_status(ledger) -> None:
    """Credits / Tier / (state) / GitHub lines, labelled as in reference cli/serve.py:876-901."""
    if ledger is None:
        click.echo("Credits:         N/A (ledger unavailable)")
        return
    ls = ledger.stats()
    click.echo(f"Credits:         {ls.balance:.1f} (earned {ls.total_earned:.1f} / spent {ls.total_spent:.1f})")
    click.echo(f"Tier:            {ls.tier.value} (score {ls.contribution_score:.1f}, search cost {ls.search_cost:.3f})")
    if ls.credit_state.value != "normal":
        extra = (f" ({ls.grace_remaining_hours:.0f} h of grace left)" if ls.credit_state.value == "grace" and ls.grace_remaining_hours is not None
                 else f" (debt {ls.debt_amount:.2f})" if ls.credit_state.value == "debt" else "")
        click.echo(f"Credit state:    {ls.credit_state.value}{extra}")
    if ls.owner_email:
        click.echo(f"GitHub:          {ls.owner_email}\n                 Credits linked across all nodes.")
    else:
        click.echo("GitHub:          " + click.style("not connected", fg="yellow"))
        click.echo("                 Run 'infomesh config github your@email.com' to link.")


@click.command()
def status() -> None:
    """Show node status."""
    from infomesh_b200.services import AppContext

    config = load_config()
    running = RT.read_live_pid(config.node.data_dir) is not None
    with AppContext(replace(config, llm=replace(config.llm, enabled=False))) as ctx:
        click.echo(f"InfoMesh v{__version__}\n{'=' * 30}")
        click.echo(f"Running:         {'yes' if running else 'no'}")
        click.echo(f"Role:            {config.node.role}")
        click.echo(f"Data dir:        {config.node.data_dir}")
        click.echo(f"Index DB:        {config.index.db_path}")
        click.echo(f"Documents:       {ctx.store.get_stats()['document_count']}")
        click.echo(f"Compression:     {'on' if config.storage.compression_enabled else 'off'} (zstd level {config.storage.compression_level})")
        click.echo(f"Vector search:   {'on' if config.index.vector_search else 'off'}")
        if ctx.vector_store is not None:
            vs = ctx.vector_store.get_stats()
            click.echo(f"Embedding model: {vs.get('model', config.index.embedding_model)}\nVector docs:     {vs.get('document_count', 0)}")
        click.echo(f"LLM:             {'on (' + config.llm.runtime + ')' if config.llm.enabled else 'off'}")
        rt = RT.read_runtime_status(config.node.data_dir)
        if rt.get("status") == "running":
            click.echo(f"Load:            {rt.get('degrade_level')} · cpu {rt.get('cpu_percent')}% · mem {rt.get('memory_percent')}% · rss {rt.get('process_memory_mb')} MB")
            if rt.get("gpu"):
                g = rt["gpu"]
                click.echo(f"GPU index:       {g.get('documents', 0)} docs · {g.get('hbm_bytes', 0) / 2 ** 20:.0f} MB HBM · graph {'on' if g.get('cuda_graph') else 'off'}")
        else:
            click.echo(f"GPU plane:       {'enabled' if config.gpu.enabled else 'disabled'} ([gpu] enabled)")
        _render_p2p_status(config, running)
        _render_credit_status(ctx.ledger)
        if ctx.key_pair is not None:
            click.echo(f"Peer ID:         {ctx.key_pair.peer_id}")
    from infomesh_b200.version_check import check_pypi_update, format_update_banner

    upd = check_pypi_update(config.node.data_dir)
    if upd is not None:
        click.secho(format_update_banner(upd), fg="yellow", bold=True)
