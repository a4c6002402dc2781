"""Owner identity from the configured / git-global e-mail address (reference infomesh/credits/github_identity.py:33-241):
it only links credits across a user's own nodes; the address itself never leaves the machine (only its SHA-256)."""
from __future__ import annotations

import re
import shutil
import subprocess
from typing import Any

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)
_EMAIL = re.compile(r"^[a-zA-Z0-9._%+-]+@[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}$")


def is_git_installed() -> bool:
    return shutil.which("git") is not None


def is_valid_email(email: str) -> bool:
    return bool(_EMAIL.match(email or ""))


def detect_git_email() -> str | None:
    try:
        res = subprocess.run(["git", "config", "--global", "user.email"], capture_output=True, text=True, timeout=5)  # noqa: S603, S607
    except (FileNotFoundError, subprocess.TimeoutExpired, OSError):
        return None
    email = res.stdout.strip() if res.returncode == 0 else ""
    return email if is_valid_email(email) else None


def resolve_github_email(config: Any) -> str | None:
    """Explicit ``node.github_email`` wins over ``git config --global user.email``."""
    if getattr(config.node, "github_email", ""):
        return config.node.github_email
    return detect_git_email()


def format_startup_message(email: str | None) -> str:
    if email:
        return f"  GitHub:  {email}\n           Credits are linked to this account across all nodes."
    return ("  GitHub:  not connected\n"
            "           Credits are tracked locally on this node only.\n"
            "           Connect your GitHub account to aggregate credits\n"
            "           across all your nodes and search for free forever.\n\n"
            "           To connect:\n"
            '             infomesh config set node.github_email "your@email.com"\n'
            "           Or set git globally:\n"
            '             git config --global user.email "your@email.com"')


def run_first_start_checks(config: Any, interactive: bool = True, *, prompt=input, echo=print) -> str | None:
    """First-start guidance; in interactive mode offers to store an address typed by the user."""
    email = resolve_github_email(config)
    echo(format_startup_message(email))
    if email or not interactive:
        return email
    try:
        typed = prompt("  Enter GitHub email (or leave empty to skip): ").strip()
    except (EOFError, KeyboardInterrupt):
        return None
    if typed and is_valid_email(typed):
        return typed
    if typed:
        echo("  That does not look like an e-mail address; skipping.")
    return None
