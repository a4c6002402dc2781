"""``infomesh keys export | rotate`` (reference infomesh/cli/keys.py:10-43)."""
from __future__ import annotations

import click

from infomesh_b200.config import load_config


@click.group("keys")
def keys_group() -> None:
    """Manage the node's Ed25519 identity."""


@keys_group.command("export")
def keys_export() -> None:
    """Print the public key (PEM) and peer id."""
    from infomesh_b200.p2p.keys import ensure_keys, export_public_key

    d = load_config().node.data_dir
    kp = ensure_keys(d)
    click.echo(f"Peer ID: {kp.peer_id}")
    click.echo(export_public_key(d))


@keys_group.command("rotate")
@click.confirmation_option(prompt="This will generate a new key pair. Are you sure?")
def keys_rotate() -> None:
    """Generate a new key pair; the old key signs a revocation record that peers can verify."""
    from infomesh_b200.p2p.keys import rotate_keys

    old, new, rec = rotate_keys(load_config().node.data_dir)
    click.secho("✔ Key rotated", fg="green")
    click.echo(f"  old peer id: {old.peer_id}\n  new peer id: {new.peer_id}\n  revocation record signed by both keys ({rec.reason})")
