"""``infomesh peer list | add MADDR | remove MADDR | test`` — manage ``network.bootstrap_nodes``
(reference infomesh/cli/peer.py:12-219)."""
from __future__ import annotations

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


_ADD_HINT = "Add with: infomesh peer add /ip4/<IP>/tcp/4001/p2p/<PEER_ID>"


def _label(text: str, colour: str) -> str:
    return click.style(text, fg=colour)


def _host_port(multiaddr: str) -> tuple[str, int] | None:
    """``(host, port)`` of a multiaddr (or ``HOST:PORT``), None when it has none."""
    from infomesh_b200.p2p.transport import parse_multiaddr

    try:
        host, port, _ = parse_multiaddr(multiaddr)
    except (ValueError, TypeError):
        return None
    return (host, int(port)) if host and port else None


def _tcp_reachable(host: str, port: int, timeout: float = 5.0) -> bool:
    import socket

    try:
        with socket.create_connection((host, port), timeout=timeout):
            return True
    except OSError:
        return False


@click.group(name="peer")
def peer_group() -> None:
    """Manage P2P peers (add, list, remove bootstrap nodes)."""


@peer_group.command("list")
def list_peers() -> None:
    """Show connected peers and bootstrap nodes."""
    cfg = load_config()
    click.echo("Bootstrap nodes:")
    nodes = _resolve(cfg.network.bootstrap_nodes)
    for line in nodes or ["(none configured)", _ADD_HINT]:
        click.echo(f"  {line}")
    click.echo()
    path = cfg.node.data_dir / "p2p_status.json"
    if not path.exists():
        click.echo("P2P state: not started")
        return
    try:
        st = json.loads(path.read_text())
        if not isinstance(st, dict):
            raise ValueError("status is not an object")
    except (OSError, ValueError):
        click.echo("P2P state: unknown (status file unreadable)")
        return
    stale = time.time() - float(st.get("timestamp", 0) or 0) >= 30 if "timestamp" in st else False
    click.echo(f"P2P state: {st.get('state', 'stopped')}{' (stale status)' if stale else ''}")
    peers = st.get("peer_ids", [])
    peers = peers if isinstance(peers, list) else []
    click.echo(f"Connected peers: {len(peers)}")
    for pid in peers:
        click.echo(f"  {pid}")
    boot = st.get("bootstrap")
    if isinstance(boot, dict) and boot:
        click.echo(f"Bootstrap: {boot.get('connected', 0)} connected, {boot.get('failed', 0)} failed")
        failed = boot.get("failed_addrs", [])
        for addr in failed if isinstance(failed, list) else []:
            click.echo("  " + _label(f"✗ {addr}", "red"))


def _write_nodes(cfg, nodes: list[str]) -> None:
    save_config(set_config_value(cfg, "network.bootstrap_nodes", ",".join(nodes)))


@peer_group.command("add")
@click.argument("multiaddr")
def add(multiaddr: str) -> None:
    """Add a bootstrap node: ``/ip4/1.2.3.4/tcp/4001/p2p/12D3KooW...``."""
    if not multiaddr.startswith(("/ip4/", "/ip6/", "/dns4/", "/dns6/")):
        click.echo(_label("Error: ", "red") + "Invalid multiaddr format.")
        click.echo("Expected: /ip4/<IP>/tcp/<PORT>/p2p/<PEER_ID>")
        return
    if "/p2p/" not in multiaddr:
        click.echo(_label("Warning: ", "yellow") + "No /p2p/<PEER_ID> in address. Connection may fail without peer ID.")
    cfg = load_config()
    nodes = list(cfg.network.bootstrap_nodes)
    if multiaddr in nodes:
        click.echo("Already in bootstrap list.")
        return
    _write_nodes(cfg, [*nodes, multiaddr])
    click.echo(_label("Added: ", "green") + multiaddr)
    click.echo("Restart the node for changes to take effect: infomesh stop && infomesh start")
    target = _host_port(multiaddr)
    if target is None:
        click.echo("(skipped — could not parse IP/port)")
        return
    click.echo(f"Testing TCP {target[0]}:{target[1]}... ", nl=False)
    if _tcp_reachable(*target):
        click.echo(_label("reachable ✓", "green"))
    else:
        click.echo(_label("unreachable ✗", "red"))
        click.echo("  Ensure the bootstrap node is running and port is open (firewall/NSG).")


@peer_group.command("remove")
@click.argument("multiaddr")
def remove(multiaddr: str) -> None:
    """Remove a bootstrap node."""
    cfg = load_config()
    nodes = list(cfg.network.bootstrap_nodes)
    if multiaddr not in nodes:
        click.echo("Not found in bootstrap list.")
        click.echo("Current nodes:")
        for addr in nodes:
            click.echo(f"  {addr}")
        return
    nodes.remove(multiaddr)
    _write_nodes(cfg, nodes)
    click.echo(_label("Removed: ", "green") + multiaddr)


@peer_group.command("test")
def test() -> None:
    """Test connectivity to all bootstrap nodes."""
    nodes = _resolve(load_config().network.bootstrap_nodes)
    if not nodes:
        click.echo("No bootstrap nodes configured.")
        click.echo(_ADD_HINT)
        return
    click.echo(f"Testing {len(nodes)} bootstrap node(s)...\n")
    reachable = 0
    for addr in nodes:
        click.echo(f"  {addr}")
        target = _host_port(addr)
        if target is None:
            click.echo(_label("    SKIP", "yellow") + " — could not parse address")
            continue
        click.echo(f"    TCP {target[0]}:{target[1]} ... ", nl=False)
        if _tcp_reachable(*target):
            reachable += 1
            click.echo(_label("OK ✓", "green"))
        else:
            click.echo(_label("FAIL ✗", "red"))
            click.echo("    → Node not running or port blocked by firewall/NSG")
    click.echo(f"\nResult: {reachable}/{len(nodes)} reachable")
    if reachable == 0:
        click.echo("\nNo bootstrap nodes reachable. Peers cannot be discovered.")
        click.echo("Troubleshooting:")
        for n, tip in enumerate(("Is the bootstrap node running?", "Is TCP port 4001 open in firewall / Azure NSG / AWS SG?", "Is the IP address correct?",
                                 "Try: nc -zv <IP> 4001 (or Test-NetConnection on Windows)"), 1):
            click.echo(f"  {n}. {tip}")
