"""P2P port reachability: is the port bound, does the host firewall drop it, which cloud are we on (IMDS probes for
Azure / AWS / GCP, WSL detection), can the provider CLI open it, and what to tell the operator otherwise
(reference infomesh/resources/port_check.py:33-1266 — the cloud-specific automation there is far longer; this keeps
the same decisions and command lines in table form)."""
from __future__ import annotations

import json
import shutil
import socket
import subprocess
import sys
from dataclasses import dataclass
from enum import StrEnum
from pathlib import Path

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)
_IMDS_TIMEOUT = 2.0
_RULE_NAME = "infomesh-p2p"


def _validate_port(port: int) -> int:
    if isinstance(port, bool) or not isinstance(port, int) or not 1 <= port <= 65535:
        raise ValueError(f"Invalid TCP port: {port!r}")
    return port


class CloudProvider(StrEnum):
    AWS = "aws"
    AZURE = "azure"
    GCP = "gcp"
    UNKNOWN = "unknown"


@dataclass(frozen=True)
class PortCheckResult:
    port: int
    is_listening: bool
    is_blocked: bool
    provider: CloudProvider
    message: str


@dataclass(frozen=True)
class NsgInfo:
    name: str
    resource_group: str
    source: str

    @staticmethod
    def from_resource_id(resource_id: str, source: str) -> "NsgInfo":
        parts = resource_id.strip().split("/")
        rg = next((parts[i + 1] for i, p in enumerate(parts[:-1]) if p.lower() == "resourcegroups"), "")
        return NsgInfo(parts[-1], rg, source)


def _http_get(url: str, headers: dict[str, str] | None = None, timeout: float = _IMDS_TIMEOUT, method: str = "GET") -> str | None:
    import urllib.request

    try:
        req = urllib.request.Request(url, headers=headers or {}, method=method)
        with urllib.request.urlopen(req, timeout=timeout) as resp:  # noqa: S310 — link-local metadata endpoints only
            return resp.read(1 << 20).decode("utf-8", errors="replace")
    except Exception:  # noqa: BLE001
        return None


def detect_cloud_provider() -> CloudProvider:
    body = _http_get("http://169.254.169.254/metadata/instance?api-version=2021-02-01", {"Metadata": "true"})
    if body and "compute" in body:
        return CloudProvider.AZURE
    token = _http_get("http://169.254.169.254/latest/api/token", {"X-aws-ec2-metadata-token-ttl-seconds": "21600"}, method="PUT")
    hdr = {"X-aws-ec2-metadata-token": token} if token else None
    body = _http_get("http://169.254.169.254/latest/meta-data/instance-id", hdr)
    if body and body.startswith("i-"):
        return CloudProvider.AWS
    body = _http_get("http://metadata.google.internal/computeMetadata/v1/instance/id", {"Metadata-Flavor": "Google"})
    if body and body.strip().isdigit():
        return CloudProvider.GCP
    return CloudProvider.UNKNOWN


def _is_wsl() -> bool:
    try:
        return "microsoft" in Path("/proc/version").read_text().lower()
    except OSError:
        return False


def _run(cmd: list[str], timeout: float = 30.0) -> tuple[int, str]:
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        return r.returncode, (r.stdout or "") + (r.stderr or "")
    except (FileNotFoundError, PermissionError):
        return 127, f"{cmd[0]}: not available"
    except subprocess.TimeoutExpired:
        return 124, f"{cmd[0]}: timed out"


def is_port_listening(port: int) -> bool:
    port = _validate_port(port)
    try:
        with socket.create_connection(("127.0.0.1", port), timeout=1.0):
            return True
    except (OSError, TimeoutError):
        return False


def _check_iptables_allows(port: int) -> bool:
    """False only when an explicit DROP/REJECT rule names the port; undeterminable == allowed."""
    port = _validate_port(port)
    for cmd in (["iptables", "-L", "INPUT", "-n", "--line-numbers"], ["nft", "list", "ruleset"]):
        rc, out = _run(cmd, 5.0)
        if rc != 0:
            continue
        for line in out.splitlines():
            if (f"dpt:{port}" in line or f"dport {port}" in line) and any(w in line for w in ("DROP", "REJECT", "drop", "reject")):
                return False
    return True


def is_port_open_externally(port: int) -> bool:
    port = _validate_port(port)
    rc, out = _run(["ss", "-tln", f"sport = :{port}"], 5.0)
    if rc == 0 and f":{port}" in out:
        return _check_iptables_allows(port)
    return True


def check_port_accessibility(port: int) -> PortCheckResult:
    port = _validate_port(port)
    blocked = not _check_iptables_allows(port)
    return PortCheckResult(port, is_port_listening(port), blocked, detect_cloud_provider(),
                           f"Port {port}/TCP may be blocked by firewall." if blocked else "")


# ------------------------------------------------------------------ provider automation
def _auto_open_azure(port: int) -> tuple[bool, str]:
    if not shutil.which("az"):
        return False, "Azure CLI (az) is not installed."
    meta = _http_get("http://169.254.169.254/metadata/instance?api-version=2021-02-01", {"Metadata": "true"})
    try:
        comp = json.loads(meta or "{}").get("compute", {})
        rg, vm = comp["resourceGroupName"], comp["name"]
    except (ValueError, KeyError):
        return False, "Could not read VM identity from the Azure metadata service."
    rc, out = _run(["az", "vm", "open-port", "--resource-group", rg, "--name", vm, "--port", str(port), "--priority", "1010"], 120.0)
    return (rc == 0, f"Opened {port}/TCP on the NSG of VM {vm}." if rc == 0 else out.strip()[-400:])


def _auto_open_aws(port: int) -> tuple[bool, str]:
    if not shutil.which("aws"):
        return False, "AWS CLI (aws) is not installed."
    token = _http_get("http://169.254.169.254/latest/api/token", {"X-aws-ec2-metadata-token-ttl-seconds": "60"}, method="PUT")
    hdr = {"X-aws-ec2-metadata-token": token} if token else None
    groups = _http_get("http://169.254.169.254/latest/meta-data/security-groups", hdr)
    mac = (_http_get("http://169.254.169.254/latest/meta-data/mac", hdr) or "").strip()
    sg_ids = _http_get(f"http://169.254.169.254/latest/meta-data/network/interfaces/macs/{mac}/security-group-ids", hdr) if mac else None
    if not sg_ids:
        return False, f"Could not determine the instance's security group (groups: {groups or 'unknown'})."
    sg = sg_ids.split()[0]
    rc, out = _run(["aws", "ec2", "authorize-security-group-ingress", "--group-id", sg, "--protocol", "tcp", "--port", str(port),
                    "--cidr", "0.0.0.0/0"], 60.0)
    if rc == 0 or "InvalidPermission.Duplicate" in out:
        return True, f"Port {port}/TCP is open in security group {sg}."
    return False, out.strip()[-400:]


def _auto_open_gcp(port: int) -> tuple[bool, str]:
    if not shutil.which("gcloud"):
        return False, "Google Cloud CLI (gcloud) is not installed."
    rc, out = _run(["gcloud", "compute", "firewall-rules", "create", f"{_RULE_NAME}-{port}", "--allow", f"tcp:{port}",
                    "--direction", "INGRESS", "--source-ranges", "0.0.0.0/0", "--quiet"], 120.0)
    if rc == 0 or "already exists" in out:
        return True, f"Firewall rule {_RULE_NAME}-{port} allows {port}/TCP."
    return False, out.strip()[-400:]


def _auto_open_wsl(port: int) -> tuple[bool, str]:
    rc, ip = _run(["hostname", "-I"], 5.0)
    wsl_ip = ip.split()[0] if rc == 0 and ip.split() else ""
    if not wsl_ip or not shutil.which("powershell.exe"):
        return False, "powershell.exe is not reachable from this WSL session."
    script = (f"netsh interface portproxy add v4tov4 listenport={port} listenaddress=0.0.0.0 connectport={port} "
              f"connectaddress={wsl_ip}; New-NetFirewallRule -DisplayName '{_RULE_NAME}-{port}' -Direction Inbound "
              f"-Action Allow -Protocol TCP -LocalPort {port}")
    rc, out = _run(["powershell.exe", "-Command", f"Start-Process powershell -Verb RunAs -ArgumentList \"{script}\""], 60.0)
    return (rc == 0, f"Forwarded Windows port {port} to WSL ({wsl_ip})." if rc == 0 else out.strip()[-400:])


_MANUAL = {
    CloudProvider.AZURE: ("Azure Portal → VM → Networking → Add inbound port rule: TCP {port}, priority 1010, Allow", 
                          "or: az vm open-port --resource-group <rg> --name <vm> --port {port} --priority 1010"),
    CloudProvider.AWS: ("EC2 Console → Security Groups → Edit inbound rules → Custom TCP {port} from 0.0.0.0/0",
                        "or: aws ec2 authorize-security-group-ingress --group-id <sg-id> --protocol tcp --port {port} --cidr 0.0.0.0/0"),
    CloudProvider.GCP: ("VPC network → Firewall → Create rule: ingress, tcp:{port}, source 0.0.0.0/0",
                        "or: gcloud compute firewall-rules create infomesh-p2p-{port} --allow tcp:{port} --direction INGRESS"),
    CloudProvider.UNKNOWN: ("Open inbound TCP {port} in your router / host firewall",
                            "e.g.: sudo ufw allow {port}/tcp"),
}


def _get_manual_instructions(provider: CloudProvider, port: int) -> str:
    return "\n".join(line.format(port=_validate_port(port)) for line in _MANUAL[provider])


def _get_wsl_manual_instructions(port: int) -> str:
    port = _validate_port(port)
    return "\n".join((
        "In an elevated Windows PowerShell:",
        f"  netsh interface portproxy add v4tov4 listenport={port} listenaddress=0.0.0.0 connectport={port} connectaddress=<WSL IP>",
        f"  New-NetFirewallRule -DisplayName 'infomesh-p2p-{port}' -Direction Inbound -Action Allow -Protocol TCP -LocalPort {port}"))


def check_port_and_offer_fix(port: int) -> bool:
    """Interactive helper used by ``infomesh start``: warn and (with consent) open the port.  Never blocks startup
    in non-interactive sessions."""
    import click

    port = _validate_port(port)
    provider = detect_cloud_provider()
    if provider == CloudProvider.UNKNOWN:
        if _is_wsl():
            click.echo(f"  ℹ P2P port: {port}/TCP (WSL detected — Windows must forward the port)")
            if sys.stdin.isatty() and click.confirm(f"    Configure Windows port forwarding for {port}/TCP?", default=True):
                ok, msg = _auto_open_wsl(port)
                click.secho(f"  {'✓' if ok else '✗'} {msg}", fg="green" if ok else "red")
                if not ok:
                    click.echo(_get_wsl_manual_instructions(port))
                return ok
            click.echo(_get_wsl_manual_instructions(port))
            return True
        click.echo(f"  ℹ P2P port: {port}/TCP")
        click.echo(f"    Ensure port {port}/TCP is open in your firewall for peering.")
        return True
    label = {CloudProvider.AZURE: "Azure NSG", CloudProvider.AWS: "AWS Security Group", CloudProvider.GCP: "GCP Firewall"}[provider]
    click.echo(f"  ℹ P2P port: {port}/TCP (detected: {provider.value.upper()} VM)")
    click.secho(f"  ⚠ Port {port}/TCP may be blocked by {provider.value.upper()} firewall.", fg="yellow")
    if not sys.stdin.isatty():
        click.echo(f"    Run interactively to auto-open port {port}/TCP in {label}.")
        return True
    if click.confirm(f"    Attempt to auto-open port {port}/TCP in {label}?", default=True):
        fn = {CloudProvider.AZURE: _auto_open_azure, CloudProvider.AWS: _auto_open_aws, CloudProvider.GCP: _auto_open_gcp}[provider]
        ok, msg = fn(port)
        if ok:
            click.secho(f"  ✓ {msg}", fg="green")
            return True
        click.secho("  ✗ Auto-open failed.", fg="red")
        click.echo(f"    {msg}")
        click.echo(_get_manual_instructions(provider, port))
        return False
    click.echo(_get_manual_instructions(provider, port))
    return True
