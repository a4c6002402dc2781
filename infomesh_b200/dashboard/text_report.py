"""``infomesh dashboard --text``: a static Rich report — node, resources (incl. GPU), index, network, credits
(reference infomesh/dashboard/text_report.py:31-346)."""
from __future__ import annotations

import io
import shutil

from rich.columns import Columns
from rich.console import Console
from rich.panel import Panel
from rich.table import Table
from rich.text import Text

from infomesh_b200 import __version__
from infomesh_b200.config import Config, load_config
from infomesh_b200.dashboard import utils as U
from infomesh_b200.runtime import read_runtime_status


def _make_bar(ratio: float, width: int = 20, color: str = "green") -> Text:
    ratio = max(0.0, min(1.0, ratio))
    fill = int(ratio * width)
    col = "red" if ratio >= 0.9 else "yellow" if ratio >= 0.7 else color
    t = Text()
    t.append("█" * fill, style=col)
    t.append("░" * (width - fill), style="dim")
    t.append(f" {ratio * 100:.0f}%", style=col)
    return t


def _grid(key_width: int = 14) -> Table:
    t = Table.grid(padding=(0, 2))
    t.add_column("key", style="bold", min_width=key_width)
    t.add_column("value")
    return t


def _node_section(config: Config) -> Panel:
    pid = U.get_peer_id(config)
    running, up = U.is_node_running_with_uptime(config)
    t = _grid(12)
    for k, v in (("Peer ID", pid[:16] + "..." if len(pid) > 16 else pid), ("State", "[bold green]🟢 Running[/]" if running else "[bold red]🔴 Stopped[/]"),
                 ("Uptime", U.format_uptime(up)), ("Version", __version__), ("Role", str(config.node.role)), ("Data dir", str(config.node.data_dir)),
                 ("Port", str(config.node.listen_port))):
        t.add_row(k, v)
    return Panel(t, title="[bold]Node[/]", border_style="cyan")


def _resource_section(config: Config) -> Panel:
    t = Table.grid(padding=(0, 1))
    t.add_column("label", min_width=8, style="bold")
    t.add_column("bar", min_width=25)
    try:
        d = config.node.data_dir
        du = shutil.disk_usage(str(d if d.exists() else "/"))
        bar = _make_bar(du.used / du.total if du.total else 0, color="yellow")
        bar.append(f"  {U.format_bytes(du.used)} / {U.format_bytes(du.total)}", style="dim")
        t.add_row("Disk", bar)
    except OSError:
        t.add_row("Disk", Text("N/A", style="dim"))
    try:
        import psutil

        t.add_row("CPU", _make_bar(psutil.cpu_percent(interval=0.1) / 100, color="cyan"))
        mem = psutil.virtual_memory()
        bar = _make_bar(mem.percent / 100)
        bar.append(f"  {U.format_bytes(mem.used)} / {U.format_bytes(mem.total)}", style="dim")
        t.add_row("RAM", bar)
    except ImportError:
        t.add_row("CPU", Text("psutil not installed", style="dim"))
    try:
        import torch

        if torch.cuda.is_available():
            free, total = torch.cuda.mem_get_info()
            bar = _make_bar(1 - free / total, color="magenta")
            bar.append(f"  {U.format_bytes(total - free)} / {U.format_bytes(total)}  {torch.cuda.get_device_name(0)}", style="dim")
            t.add_row("HBM", bar)
    except Exception:  # noqa: BLE001
        pass
    rt = read_runtime_status(config.node.data_dir)
    if rt.get("status") == "running":
        t.add_row("Load", Text(f"{rt.get('degrade_level')}  throttle ×{rt.get('throttle_factor')}", style="dim"))
    return Panel(t, title="[bold]Resources[/]", border_style="cyan")


def _index_section(config: Config) -> Panel:
    t = _grid()
    try:
        from infomesh_b200.index.local_store import LocalStore

        with LocalStore(db_path=config.index.db_path, compression_enabled=config.storage.compression_enabled,
                        compression_level=config.storage.compression_level) as st:
            t.add_row("Documents", f"{st.get_stats()['document_count']:,}")
            if config.index.db_path.exists():
                t.add_row("DB size", U.format_bytes(config.index.db_path.stat().st_size))
            top = st.get_top_domains(limit=5)
            if top:
                t.add_row("Top domains", ", ".join(f"{d} ({c})" for d, c in top))
    except Exception as exc:  # noqa: BLE001
        t.add_row("Error", str(exc))
    gpu = read_runtime_status(config.node.data_dir).get("gpu")
    if isinstance(gpu, dict):
        t.add_row("GPU index", f"{gpu.get('documents', 0):,} docs · {U.format_bytes(gpu.get('hbm_bytes', 0))} · batch {gpu.get('query_batch')}")
    return Panel(t, title="[bold]Index[/]", border_style="cyan")


def _network_section(config: Config) -> Panel:
    t = _grid()
    st = U.read_p2p_status(config)
    if st:
        state = str(st.get("state", "stopped"))
        label = {"running": "[bold green]🟢 Online[/]", "starting": "[bold yellow]🟡 Starting[/]",
                 "error": f"[bold red]🔴 Error {st.get('error', '')}[/]"}.get(state, "[bold red]🔴 Offline[/]")
        t.add_row("P2P State", label)
        t.add_row("Peers", f"{int(st.get('peers', 0) or 0)} connected")
        dht = st.get("dht", {})
        if isinstance(dht, dict) and any(dht.values()):
            t.add_row("DHT keys", f"{dht.get('keys_stored', 0):,} stored, {dht.get('keys_published', 0):,} published")
        bw = st.get("bandwidth", {})
        if isinstance(bw, dict) and any(bw.values()):
            t.add_row("Traffic", f"↑ {U.format_bytes(bw.get('upload_bytes', 0))}  ↓ {U.format_bytes(bw.get('download_bytes', 0))}")
    else:
        t.add_row("P2P State", "[dim]Not started — run infomesh start[/]")
    t.add_row("Port", f"{config.node.listen_port} TCP")
    t.add_row("Bootstrap", f"{len(config.network.bootstrap_nodes)} nodes configured")
    t.add_row("Replication", f"{config.network.replication_factor}x")
    t.add_row("Limits", f"↑ {config.network.upload_limit_mbps:.1f} Mbps  ↓ {config.network.download_limit_mbps:.1f} Mbps")
    return Panel(t, title="[bold]Network[/]", border_style="cyan")


def _credits_section(config: Config) -> Panel:
    t = _grid()
    path = config.node.data_dir / "credits.db"
    if not path.exists():
        t.add_row("Status", "[dim]No credit history yet[/]")
        t.add_row("Hint", "Start crawling to earn credits!")
        return Panel(t, title="[bold]Credits[/]", border_style="cyan")
    try:
        from infomesh_b200.credits.ledger import CreditLedger

        led = CreditLedger(path)
        try:
            s, al = led.stats(), led.search_allowance()
            t.add_row("Balance", f"[bold green]{s.balance:,.2f}[/] credits")
            t.add_row("Tier", U.tier_label(s.tier))
            t.add_row("Earned / spent", f"{s.total_earned:,.2f} / {s.total_spent:,.2f}")
            t.add_row("Search cost", f"{al.search_cost:.3f} ({al.state.value})")
            t.add_row("Score", f"{s.contribution_score:,.2f}")
        finally:
            led.close()
    except Exception as exc:  # noqa: BLE001
        t.add_row("Error", str(exc))
    return Panel(t, title="[bold]Credits[/]", border_style="cyan")


def build_report(config: Config | None = None) -> list[Panel]:
    cfg = config or load_config()
    return [_node_section(cfg), _resource_section(cfg), _index_section(cfg), _network_section(cfg), _credits_section(cfg)]


def render_text_report(config: Config | None = None, *, width: int = 100) -> str:
    buf = io.StringIO()
    con = Console(file=buf, width=width, force_terminal=False, color_system=None)
    panels = build_report(config)
    con.print(Text(f"InfoMesh v{__version__} — dashboard", style="bold"))
    con.print(Columns(panels[:2], equal=True, expand=True))
    for p in panels[2:]:
        con.print(p)
    return buf.getvalue()


def print_text_report(config: Config | None = None) -> None:
    con = Console()
    panels = build_report(config)
    con.print(Columns(panels[:2], equal=True, expand=True))
    for p in panels[2:]:
        con.print(p)


_TAB_SECTIONS = {"overview": (_node_section, _resource_section), "crawl": (_index_section,), "search": (_index_section,),
                 "network": (_network_section,), "credits": (_credits_section,), "settings": (_node_section,)}


def print_dashboard(config: Config | None = None, *, tab: str | None = None) -> None:
    """Rich snapshot on the console; ``tab`` narrows it to the sections of one dashboard tab (reference text_report.py:289)."""
    cfg = config or load_config()
    if tab is None:
        print_text_report(cfg)
        return
    con = Console()
    for section in _TAB_SECTIONS.get(tab.lower(), (_node_section,)):
        con.print(section(cfg))
