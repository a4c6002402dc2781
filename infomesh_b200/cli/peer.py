"""``infomesh peer list | add MADDR | remove MADDR | test`` — manage ``network.bootstrap_nodes``
(reference infomesh/cli/peer.py:12-219)."""
from __future__ import annotations

import asyncio
import json
import time

import click

from infomesh_b200.config import load_config, save_config, set_config_value


def _resolve(nodes: list[str]) -> list[str]:
    from infomesh_b200.p2p.bootstrap import bundled_nodes

    out: list[str] = []
    for n in nodes:
        out += [e["addr"] for e in bundled_nodes() if "addr" in e] if n == "default" else [n]
    return list(dict.fromkeys(out))


@click.group(name="peer")
def peer_group() -> None:
    """Bootstrap peers and connectivity."""


@peer_group.command("list")
def list_peers() -> None:
    """Configured bootstrap nodes and currently connected peers."""
    cfg = load_config()
    click.secho("Configured bootstrap nodes:", bold=True)
    for a in _resolve(cfg.network.bootstrap_nodes) or ["(none)"]:
        click.echo(f"  {a}")
    path = cfg.node.data_dir / "p2p_status.json"
    try:
        st = json.loads(path.read_text())
        fresh = time.time() - float(st.get("timestamp", 0)) < 30
        click.secho(f"\nNode {st.get('state', '?')}{'' if fresh else ' (stale status)'} — {st.get('peers', 0)} peer(s) connected", bold=True)
        for pid in st.get("peer_ids", [])[:50]:
            click.echo(f"  {pid}")
    except (OSError, ValueError):
        click.echo("\nNode not running (no live peer list).")


def _write_nodes(cfg, nodes: list[str]) -> None:
    save_config(set_config_value(cfg, "network.bootstrap_nodes", ",".join(nodes)))


@peer_group.command("add")
@click.argument("multiaddr")
def add(multiaddr: str) -> None:
    """Add a bootstrap node (``/ip4/HOST/tcp/PORT[/p2p/ID]`` or ``HOST:PORT``)."""
    from infomesh_b200.p2p.transport import parse_multiaddr

    try:
        parse_multiaddr(multiaddr)
    except ValueError as exc:
        raise click.ClickException(str(exc)) from None
    cfg = load_config()
    nodes = list(cfg.network.bootstrap_nodes)
    if multiaddr in nodes:
        click.echo("Already configured.")
        return
    _write_nodes(cfg, nodes + [multiaddr])
    click.secho(f"✔ Added {multiaddr}. Restart the node to connect.", fg="green")


@peer_group.command("remove")
@click.argument("multiaddr")
def remove(multiaddr: str) -> None:
    """Remove a bootstrap node."""
    cfg = load_config()
    nodes = list(cfg.network.bootstrap_nodes)
    if multiaddr not in nodes:
        raise click.ClickException("not in the configured bootstrap list")
    nodes.remove(multiaddr)
    _write_nodes(cfg, nodes)
    click.secho(f"✔ Removed {multiaddr}", fg="green")


@peer_group.command("test")
def test() -> None:
    """TCP-probe every configured bootstrap node."""
    from infomesh_b200.p2p.bootstrap import BootstrapNode, check_all_bootstrap_health

    nodes = [BootstrapNode(a, "config") for a in _resolve(load_config().network.bootstrap_nodes)]
    if not nodes:
        click.echo("No bootstrap nodes configured.")
        return
    for h in asyncio.run(check_all_bootstrap_health(nodes)):
        if h.reachable:
            click.secho(f"  ✔ {h.addr}  {h.latency_ms:.0f} ms", fg="green")
        else:
            click.secho(f"  ✖ {h.addr}  unreachable", fg="red")
