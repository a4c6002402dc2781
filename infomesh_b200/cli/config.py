"""``infomesh config show | set KEY VALUE | github [EMAIL]`` (reference infomesh/cli/config.py:14-174)."""
from __future__ import annotations

import click

from infomesh_b200.config import DEFAULT_CONFIG_PATH, config_to_dict, load_config, save_config, set_config_value


@click.group("config")
def config_group() -> None:
    """Show or edit the node configuration."""


@config_group.command("show")
def config_show() -> None:
    """Print the effective configuration (secrets redacted)."""
    for section, values in config_to_dict(load_config(), redact=True).items():
        click.secho(f"[{section}]", bold=True)
        for k, v in values.items():
            click.echo(f"  {k} = {v}")
        click.echo()


@config_group.command("set")
@click.argument("key")
@click.argument("value")
def config_set(key: str, value: str) -> None:
    """Set SECTION.KEY to VALUE and persist it (validated and clamped like a config file value)."""
    try:
        new = set_config_value(load_config(), key, value)
    except (KeyError, ValueError, TypeError) as exc:
        raise click.ClickException(str(exc)) from None
    save_config(new)
    section, _, name = key.partition(".")
    click.secho(f"Set {key} = {getattr(getattr(new, section), name)}", fg="green")
    click.echo(f"  saved to {DEFAULT_CONFIG_PATH}")


@config_group.command("github")
@click.argument("email", required=False)
def config_github(email: str | None) -> None:
    """Show or set the GitHub e-mail that links credits across your nodes."""
    from infomesh_b200.credits.github_identity import is_valid_email, resolve_github_email

    cfg = load_config()
    if email is None:
        cur = resolve_github_email(cfg)
        click.echo(f"GitHub identity: {cur}" if cur else "No GitHub identity configured (credits stay node-local).")
        return
    if not is_valid_email(email):
        raise click.ClickException(f"not a valid e-mail address: {email}")
    save_config(set_config_value(cfg, "node.github_email", email))
    click.secho(f"✔ GitHub identity set to {email}", fg="green")
