"""``infomesh keys export | rotate`` (reference infomesh/cli/keys.py:10-43)."""
from __future__ import annotations

import click

from infomesh_b200.config import load_config


@click.group("keys")
def keys_group() -> None:
    """Manage the node's Ed25519 identity."""


@keys_group.command("export")
def keys_export() -> None:
    """Export the public key (PEM).  Keys are created by ``infomesh start``; without them this says so and changes nothing."""
    from infomesh_b200.p2p.keys import export_public_key

    try:
        click.echo(export_public_key(load_config().node.data_dir))      # the PEM and nothing else: `keys export > node.pem` must work
    except FileNotFoundError as exc:
        click.echo(str(exc))


@keys_group.command("rotate")
@click.confirmation_option(prompt="This will generate a new key pair. Are you sure?")
def keys_rotate() -> None:
    """Rotate the Ed25519 key pair; the old key signs a revocation record that peers can verify."""
    from infomesh_b200.p2p.keys import rotate_keys

    try:
        old, new, record = rotate_keys(load_config().node.data_dir)
    except FileNotFoundError as exc:
        click.echo(str(exc))
        return
    click.echo("Old keys backed up. Revocation record saved.")
    click.echo(f"  Old Peer ID: {old.peer_id}")
    click.echo(f"  New Peer ID: {new.peer_id}")
    click.echo(f"  Reason: {record.reason}")
    click.echo("\nRevocation will be published to DHT on next node start.")
