"""P2P port reachability: is the port bound, does the host firewall drop it, which cloud are we on (IMDS probes for
Azure / AWS / GCP, WSL detection), can the provider CLI open it, and what to tell the operator otherwise
(reference infomesh/resources/port_check.py:33-1266).

Automation per environment -- each step shells out to the provider's own CLI, never to an SDK:
* Azure: VM identity from IMDS, then EVERY network security group in the path (NIC-level and subnet-level) gets an inbound
  allow rule; falls back to ``az vm open-port`` when the groups cannot be listed.
* AWS: security-group ids of every interface from IMDSv2 (v1 fallback), ingress authorised in each; a duplicate rule counts
  as success.
* GCP: a tagged ingress firewall rule plus the tag on this instance.
* WSL2: Windows firewall rule + ``netsh portproxy`` forward; a forward left pointing at the VM's previous address (it changes
  at every WSL restart) is re-pointed without asking."""
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


# ------------------------------------------------------------------ instance metadata (link-local IMDS endpoints)
_AZURE_IMDS = "http://169.254.169.254/metadata/instance?api-version=2021-02-01"
_AWS_IMDS = "http://169.254.169.254/latest"
_GCP_IMDS = "http://metadata.google.internal/computeMetadata/v1"


def _get_azure_metadata() -> dict[str, object] | None:
    """The ``compute`` block of the Azure instance document plus the first interface's private IP; ``None`` off Azure."""
    try:
        doc = json.loads(_http_get(_AZURE_IMDS, {"Metadata": "true"}) or "")
        comp = doc["compute"]
    except (ValueError, KeyError, TypeError):
        return None
    nics = doc.get("network", {}).get("interface", [])
    private_ip = ""
    try:
        private_ip = nics[0]["ipv4"]["ipAddress"][0]["privateIpAddress"]
    except (IndexError, KeyError, TypeError):
        pass
    return {"vm_name": comp.get("name", ""), "resource_group": comp.get("resourceGroupName", ""),
            "subscription_id": comp.get("subscriptionId", ""), "location": comp.get("location", ""), "private_ip": private_ip}


def _aws_imds_headers() -> dict[str, str] | None:
    token = _http_get(f"{_AWS_IMDS}/api/token", {"X-aws-ec2-metadata-token-ttl-seconds": "60"}, method="PUT")
    return {"X-aws-ec2-metadata-token": token} if token else None          # IMDSv1 fallback when no token is issued


def _get_aws_metadata() -> dict[str, str]:
    """instance id, region and the security-group ids of EVERY network interface (comma separated)."""
    hdr = _aws_imds_headers()

    def get(path: str) -> str:
        return (_http_get(f"{_AWS_IMDS}/meta-data/{path}", hdr) or "").strip()

    group_ids: list[str] = []
    for mac in get("network/interfaces/macs/").split():
        for sg in get(f"network/interfaces/macs/{mac.rstrip('/')}/security-group-ids").split():
            if sg.startswith("sg-") and sg not in group_ids:
                group_ids.append(sg)
    zone = get("placement/availability-zone")
    return {"instance_id": get("instance-id"), "region": zone[:-1] if zone and zone[-1].isalpha() else zone,
            "security_group_ids": ",".join(group_ids), "security_groups": get("security-groups").replace("\n", ",")}


def _get_gcp_metadata() -> dict[str, str]:
    def get(path: str) -> str:
        return (_http_get(f"{_GCP_IMDS}/{path}", {"Metadata-Flavor": "Google"}) or "").strip()

    zone, network = get("instance/zone"), get("instance/network-interfaces/0/network")
    return {"name": get("instance/name"), "zone": zone.rsplit("/", 1)[-1], "project": get("project/project-id"),
            "network": network.rsplit("/", 1)[-1], "tags": get("instance/tags")}


# ------------------------------------------------------------------ provider automation
def _az_json(args: list[str], timeout: float = 60.0):
    rc, out = _run(["az", *args, "--output", "json"], timeout)
    if rc != 0:
        return None
    try:
        return json.loads(out[out.index("["):] if out.lstrip().startswith("[") else out[out.index("{"):])
    except ValueError:
        return None


def _discover_azure_nsgs(vm_name: str, rg: str, subscription_id: str) -> list[NsgInfo]:
    """Every network security group that filters the VM's traffic: the one attached to each NIC and the one attached to
    each NIC's subnet (Azure evaluates both; opening only one of them leaves the port closed)."""
    sub = ["--subscription", subscription_id] if subscription_id else []
    found: dict[str, NsgInfo] = {}
    for nic_ref in _az_json(["vm", "nic", "list", "--resource-group", rg, "--vm-name", vm_name, *sub]) or []:
        nic = _az_json(["network", "nic", "show", "--ids", str(nic_ref.get("id", "")), *sub]) or {}
        nsg_id = (nic.get("networkSecurityGroup") or {}).get("id")
        if nsg_id:
            found.setdefault(nsg_id.lower(), NsgInfo.from_resource_id(nsg_id, "nic"))
        for ipcfg in nic.get("ipConfigurations", []):
            subnet_id = (ipcfg.get("subnet") or {}).get("id")
            if not subnet_id:
                continue
            subnet = _az_json(["network", "vnet", "subnet", "show", "--ids", subnet_id, *sub]) or {}
            nsg_id = (subnet.get("networkSecurityGroup") or {}).get("id")
            if nsg_id:
                found.setdefault(nsg_id.lower(), NsgInfo.from_resource_id(nsg_id, "subnet"))
    return list(found.values())


def _auto_open_azure(port: int) -> tuple[bool, str]:
    port = _validate_port(port)
    if not shutil.which("az"):
        return False, "Azure CLI (az) is not installed."
    meta = _get_azure_metadata()
    if not meta or not meta["vm_name"] or not meta["resource_group"]:
        return False, "Could not read VM identity from the Azure metadata service."
    vm, rg, sub = str(meta["vm_name"]), str(meta["resource_group"]), str(meta["subscription_id"])
    nsgs = _discover_azure_nsgs(vm, rg, sub)
    if not nsgs:        # no NSG visible (or no permission to list): let the CLI pick the VM's default one
        rc, out = _run(["az", "vm", "open-port", "--resource-group", rg, "--name", vm, "--port", str(port), "--priority", "1010"], 120.0)
        return (rc == 0, f"Opened {port}/TCP on the NSG of VM {vm}." if rc == 0 else out.strip()[-400:])
    opened, errors = [], []
    for nsg in nsgs:
        rc, out = _run(["az", "network", "nsg", "rule", "create", "--resource-group", nsg.resource_group, "--nsg-name", nsg.name,
                        "--name", f"{_RULE_NAME}-{port}", "--priority", "1010", "--direction", "Inbound", "--access", "Allow",
                        "--protocol", "Tcp", "--destination-port-ranges", str(port), *(["--subscription", sub] if sub else [])], 120.0)
        (opened if rc == 0 else errors).append(f"{nsg.name} ({nsg.source})" if rc == 0 else f"{nsg.name}: {out.strip()[-200:]}")
    if errors:
        return False, "; ".join(errors)
    return True, f"Opened {port}/TCP in " + ", ".join(opened) + "."


def _auto_open_aws(port: int) -> tuple[bool, str]:
    port = _validate_port(port)
    if not shutil.which("aws"):
        return False, "AWS CLI (aws) is not installed."
    meta = _get_aws_metadata()
    groups = [g for g in meta["security_group_ids"].split(",") if g]
    if not groups:
        return False, f"Could not determine the instance's security groups (named: {meta['security_groups'] or 'unknown'})."
    region = ["--region", meta["region"]] if meta["region"] else []
    done, errors = [], []
    for sg in groups:           # an instance is reachable only if EVERY attached group... no: ANY group allowing is enough,
        rc, out = _run(["aws", "ec2", "authorize-security-group-ingress", "--group-id", sg, "--protocol", "tcp", "--port", str(port),
                        "--cidr", "0.0.0.0/0", *region], 60.0)   # but opening all of them survives a later detach
        if rc == 0 or "InvalidPermission.Duplicate" in out:
            done.append(sg)
        else:
            errors.append(f"{sg}: {out.strip()[-200:]}")
    if done:
        return True, f"Port {port}/TCP is open in security group(s) {', '.join(done)}."
    return False, "; ".join(errors)


_GCP_TAG = "infomesh-p2p"


def _auto_open_gcp(port: int) -> tuple[bool, str]:
    """A tagged ingress rule plus the tag on this instance (a rule without a matching tag, or a tag without a rule, opens
    nothing)."""
    port = _validate_port(port)
    if not shutil.which("gcloud"):
        return False, "Google Cloud CLI (gcloud) is not installed."
    meta = _get_gcp_metadata()
    project = ["--project", meta["project"]] if meta["project"] else []
    network = ["--network", meta["network"]] if meta["network"] else []
    rc, out = _run(["gcloud", "compute", "firewall-rules", "create", f"{_RULE_NAME}-{port}", "--allow", f"tcp:{port}", "--direction", "INGRESS",
                    "--source-ranges", "0.0.0.0/0", "--target-tags", _GCP_TAG, *network, *project, "--quiet"], 120.0)
    if rc != 0 and "already exists" not in out:
        return False, out.strip()[-400:]
    if meta["name"] and meta["zone"] and _GCP_TAG not in meta["tags"]:
        rc, out = _run(["gcloud", "compute", "instances", "add-tags", meta["name"], "--zone", meta["zone"], "--tags", _GCP_TAG, *project, "--quiet"], 120.0)
        if rc != 0:
            return False, f"Firewall rule created, but tagging the instance failed: {out.strip()[-300:]}"
    return True, f"Firewall rule {_RULE_NAME}-{port} allows {port}/TCP to instances tagged {_GCP_TAG}."


# ------------------------------------------------------------------ WSL2 (Windows forwards the port into the VM)
def _powershell(script: str, timeout: float = 30.0) -> tuple[int, str]:
    return _run(["powershell.exe", "-NoProfile", "-Command", script], timeout)


def _get_wsl_ip() -> str | None:
    """The WSL2 VM's own address (what the Windows side must forward to)."""
    rc, out = _run(["ip", "-4", "-o", "addr", "show", "eth0"], 5.0)
    if rc == 0:
        for tok in out.split():
            if "/" in tok and tok.split("/")[0].count(".") == 3:
                return tok.split("/")[0]
    rc, out = _run(["hostname", "-I"], 5.0)
    return out.split()[0] if rc == 0 and out.split() else None


def _get_wsl_host_ip() -> str | None:
    """The Windows host as seen from inside WSL2: the default gateway (mirrored in /etc/resolv.conf on older builds)."""
    rc, out = _run(["ip", "route", "show", "default"], 5.0)
    parts = out.split()
    if rc == 0 and "via" in parts:
        return parts[parts.index("via") + 1]
    try:
        for line in Path("/etc/resolv.conf").read_text().splitlines():
            if line.startswith("nameserver"):
                return line.split()[1]
    except (OSError, IndexError):
        pass
    return None


def _wsl_firewall_exists(port: int) -> bool:
    port = _validate_port(port)
    rc, out = _powershell(f"Get-NetFirewallRule -DisplayName '{_RULE_NAME}-{port}' -ErrorAction SilentlyContinue | Select-Object -ExpandProperty Enabled")
    return rc == 0 and "true" in out.lower()


def _wsl_portproxy_target(port: int) -> str | None:
    """``connectaddress`` of the v4tov4 portproxy entry that listens on ``port``, or None."""
    port = _validate_port(port)
    rc, out = _run(["netsh.exe", "interface", "portproxy", "show", "v4tov4"], 15.0)
    if rc != 0:
        return None
    for line in out.splitlines():
        cols = line.split()
        if len(cols) >= 4 and cols[1] == str(port) and cols[2].count(".") == 3:       # listen-addr listen-port connect-addr connect-port
            return cols[2]
    return None


def _wsl_update_portproxy(port: int, wsl_ip: str) -> bool:
    """Point an existing forward at the VM's current address (it changes on every WSL restart).  Needs elevation: goes
    through an elevated PowerShell, silently."""
    port = _validate_port(port)
    inner = (f"netsh interface portproxy delete v4tov4 listenport={port} listenaddress=0.0.0.0; "
             f"netsh interface portproxy add v4tov4 listenport={port} listenaddress=0.0.0.0 connectport={port} connectaddress={wsl_ip}")
    rc, _ = _powershell(f"Start-Process powershell -Verb RunAs -WindowStyle Hidden -Wait -ArgumentList '-NoProfile -Command {inner}'", 60.0)
    return rc == 0 and _wsl_portproxy_target(port) == wsl_ip


def _auto_open_wsl(port: int) -> tuple[bool, str]:
    port = _validate_port(port)
    wsl_ip = _get_wsl_ip()
    if not wsl_ip or not shutil.which("powershell.exe"):
        return False, "powershell.exe is not reachable from this WSL session."
    script = (f"netsh interface portproxy add v4tov4 listenport={port} listenaddress=0.0.0.0 connectport={port} "
              f"connectaddress={wsl_ip}; New-NetFirewallRule -DisplayName '{_RULE_NAME}-{port}' -Direction Inbound "
              f"-Action Allow -Protocol TCP -LocalPort {port}")
    rc, out = _powershell(f"Start-Process powershell -Verb RunAs -Wait -ArgumentList \"{script}\"", 60.0)
    return (rc == 0, f"Forwarded Windows port {port} to WSL ({wsl_ip})." if rc == 0 else out.strip()[-400:])


def _check_port_wsl(port: int) -> bool:
    """WSL flow of ``check_port_and_offer_fix``: nothing to do when the Windows firewall rule exists and the forward points
    at the VM's current address; a stale forward is repaired without asking; otherwise offer to set both up."""
    import click

    port = _validate_port(port)
    wsl_ip = _get_wsl_ip()
    click.echo(f"  ℹ P2P port: {port}/TCP (WSL detected — Windows must forward the port)")
    target = _wsl_portproxy_target(port) if shutil.which("netsh.exe") else None
    if target is not None and _wsl_firewall_exists(port):
        if wsl_ip and target != wsl_ip:
            fixed = _wsl_update_portproxy(port, wsl_ip)
            click.secho(f"  {'✓' if fixed else '✗'} Windows forward for {port}/TCP re-pointed from {target} to {wsl_ip}"
                        if fixed else f"  ✗ Windows forwards {port}/TCP to {target}, but this VM is {wsl_ip}", fg="green" if fixed else "red")
            if not fixed:
                click.echo(_get_wsl_manual_instructions(port))
            return fixed
        click.secho(f"  ✓ Windows forwards {port}/TCP to this VM", fg="green")
        return True
    if sys.stdin.isatty() and click.confirm(f"    Configure Windows port forwarding for {port}/TCP?", default=True):
        ok, msg = _auto_open_wsl(port)
        click.secho(f"  {'✓' if ok else '✗'} {msg}", fg="green" if ok else "red")
        if not ok:
            click.echo(_get_wsl_manual_instructions(port))
        return ok
    click.echo(_get_wsl_manual_instructions(port))
    return True


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
            return _check_port_wsl(port)
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
