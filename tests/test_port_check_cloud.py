"""resources/port_check.py cloud + WSL automation with the provider CLIs and metadata services faked (model: reference
tests/test_port_check.py): which commands are issued, in which order, and how their outcomes are reported."""
import json

import pytest

from infomesh_b200.resources import port_check as PC


class Shell:
    """Scripted replacement for port_check._run: records argv, answers by prefix match."""

    def __init__(self, answers):
        self.answers, self.calls = answers, []

    def __call__(self, cmd, timeout=30.0):
        self.calls.append(list(cmd))
        line = " ".join(cmd)
        for prefix, reply in self.answers:
            if prefix in line:
                return reply(cmd) if callable(reply) else reply
        return 0, ""


@pytest.fixture
def cli(monkeypatch):
    monkeypatch.setattr(PC.shutil, "which", lambda name: f"/usr/bin/{name}")

    def install(answers):
        sh = Shell(answers)
        monkeypatch.setattr(PC, "_run", sh)
        return sh

    return install


def imds(monkeypatch, table):
    def fake(url, headers=None, timeout=2.0, method="GET"):
        for key, val in table.items():
            if key in url:
                return val
        return None

    monkeypatch.setattr(PC, "_http_get", fake)


AZ_DOC = json.dumps({"compute": {"name": "vm1", "resourceGroupName": "rg1", "subscriptionId": "sub-9", "location": "koreacentral"},
                     "network": {"interface": [{"ipv4": {"ipAddress": [{"privateIpAddress": "10.0.0.4"}]}}]}})
NIC_NSG = "/subscriptions/sub-9/resourceGroups/rg1/providers/Microsoft.Network/networkSecurityGroups/nic-nsg"
SUBNET_NSG = "/subscriptions/sub-9/resourceGroups/net-rg/providers/Microsoft.Network/networkSecurityGroups/subnet-nsg"
SUBNET = "/subscriptions/sub-9/resourceGroups/net-rg/providers/Microsoft.Network/virtualNetworks/v/subnets/s"


def test_azure_metadata_and_nsg_discovery_cover_nic_and_subnet(monkeypatch, cli):
    imds(monkeypatch, {"metadata/instance": AZ_DOC})
    assert PC._get_azure_metadata() == {"vm_name": "vm1", "resource_group": "rg1", "subscription_id": "sub-9", "location": "koreacentral",
                                        "private_ip": "10.0.0.4"}
    sh = cli([("vm nic list", (0, json.dumps([{"id": "/nic/1"}]))),
              ("network nic show", (0, json.dumps({"networkSecurityGroup": {"id": NIC_NSG}, "ipConfigurations": [{"subnet": {"id": SUBNET}}, {"subnet": {"id": SUBNET}}]}))),
              ("vnet subnet show", (0, json.dumps({"networkSecurityGroup": {"id": SUBNET_NSG}})))])
    nsgs = PC._discover_azure_nsgs("vm1", "rg1", "sub-9")
    assert [(n.name, n.resource_group, n.source) for n in nsgs] == [("nic-nsg", "rg1", "nic"), ("subnet-nsg", "net-rg", "subnet")]
    assert all("--subscription" in c for c in sh.calls)


def test_azure_auto_open_writes_a_rule_into_every_nsg(monkeypatch, cli):
    imds(monkeypatch, {"metadata/instance": AZ_DOC})
    sh = cli([("vm nic list", (0, json.dumps([{"id": "/nic/1"}]))),
              ("network nic show", (0, json.dumps({"networkSecurityGroup": {"id": NIC_NSG}, "ipConfigurations": [{"subnet": {"id": SUBNET}}]}))),
              ("vnet subnet show", (0, json.dumps({"networkSecurityGroup": {"id": SUBNET_NSG}}))), ("nsg rule create", (0, "{}"))])
    ok, msg = PC._auto_open_azure(4001)
    rules = [c for c in sh.calls if "rule" in c and "create" in c]
    assert ok and "nic-nsg (nic)" in msg and "subnet-nsg (subnet)" in msg and len(rules) == 2
    assert all(c[c.index("--destination-port-ranges") + 1] == "4001" and c[c.index("--name") + 1] == "infomesh-p2p-4001" for c in rules)
    assert {c[c.index("--resource-group") + 1] for c in rules} == {"rg1", "net-rg"}


def test_azure_falls_back_to_open_port_and_reports_failures(monkeypatch, cli):
    imds(monkeypatch, {"metadata/instance": AZ_DOC})
    sh = cli([("vm nic list", (1, "AuthorizationFailed")), ("vm open-port", (0, "{}"))])
    assert PC._auto_open_azure(4001) == (True, "Opened 4001/TCP on the NSG of VM vm1.")
    assert any(c[:3] == ["az", "vm", "open-port"] for c in sh.calls)
    cli([("vm nic list", (0, json.dumps([{"id": "/nic/1"}]))), ("network nic show", (0, json.dumps({"networkSecurityGroup": {"id": NIC_NSG}}))),
         ("nsg rule create", (1, "ERROR: priority 1010 already in use"))])
    ok, msg = PC._auto_open_azure(4001)
    assert not ok and "nic-nsg" in msg and "priority" in msg
    imds(monkeypatch, {})
    assert PC._auto_open_azure(4001) == (False, "Could not read VM identity from the Azure metadata service.")
    assert PC._get_azure_metadata() is None


def test_aws_opens_every_security_group_and_tolerates_duplicates(monkeypatch, cli):
    imds(monkeypatch, {"api/token": "tok", "interfaces/macs/0a:01/security-group-ids": "sg-1\nsg-2", "interfaces/macs/0a:02/security-group-ids": "sg-2\nsg-3",
                       "network/interfaces/macs/": "0a:01/\n0a:02/", "placement/availability-zone": "ap-northeast-2c", "meta-data/instance-id": "i-0abc",
                       "meta-data/security-groups": "web\ndb"})
    meta = PC._get_aws_metadata()
    assert meta["security_group_ids"] == "sg-1,sg-2,sg-3" and meta["region"] == "ap-northeast-2" and meta["instance_id"] == "i-0abc"
    sh = cli([("--group-id sg-2", (254, "An error occurred (InvalidPermission.Duplicate)")), ("--group-id sg-3", (254, "UnauthorizedOperation")),
              ("authorize-security-group-ingress", (0, "{}"))])
    ok, msg = PC._auto_open_aws(4001)
    assert ok and "sg-1, sg-2" in msg and "sg-3" not in msg
    assert all(c[c.index("--region") + 1] == "ap-northeast-2" and c[c.index("--port") + 1] == "4001" for c in sh.calls)
    cli([("authorize-security-group-ingress", (254, "UnauthorizedOperation"))])
    ok, msg = PC._auto_open_aws(4001)
    assert not ok and msg.count("UnauthorizedOperation") == 3
    imds(monkeypatch, {})
    assert PC._auto_open_aws(4001)[0] is False


def test_gcp_creates_a_tagged_rule_and_tags_the_instance(monkeypatch, cli):
    table = {"instance/name": "gpu-node", "instance/zone": "projects/1/zones/asia-northeast3-a", "project/project-id": "proj-x",
             "network-interfaces/0/network": "projects/1/networks/default", "instance/tags": '["http-server"]'}
    imds(monkeypatch, table)
    assert PC._get_gcp_metadata() == {"name": "gpu-node", "zone": "asia-northeast3-a", "project": "proj-x", "network": "default", "tags": '["http-server"]'}
    sh = cli([("firewall-rules create", (1, "ERROR: The resource 'infomesh-p2p-4001' already exists")), ("instances add-tags", (0, ""))])
    ok, msg = PC._auto_open_gcp(4001)
    create, tag = sh.calls
    assert ok and "infomesh-p2p" in msg and create[create.index("--target-tags") + 1] == "infomesh-p2p" and create[create.index("--network") + 1] == "default"
    assert tag[:4] == ["gcloud", "compute", "instances", "add-tags"] and tag[4] == "gpu-node" and tag[tag.index("--zone") + 1] == "asia-northeast3-a"
    imds(monkeypatch, dict(table, **{"instance/tags": '["infomesh-p2p"]'}))
    sh = cli([("firewall-rules create", (0, ""))])
    assert PC._auto_open_gcp(4001)[0] and len(sh.calls) == 1                       # already tagged: no second command
    sh = cli([("firewall-rules create", (1, "PERMISSION_DENIED"))])
    assert PC._auto_open_gcp(4001) == (False, "PERMISSION_DENIED")
    imds(monkeypatch, table)
    cli([("firewall-rules create", (0, "")), ("instances add-tags", (1, "denied"))])
    ok, msg = PC._auto_open_gcp(4001)
    assert not ok and "tagging the instance failed" in msg


PORTPROXY = """
Listen on ipv4:             Connect to ipv4:

Address         Port        Address         Port
--------------- ----------  --------------- ----------
0.0.0.0         4001        172.20.1.5      4001
0.0.0.0         8080        172.20.1.5      8080
"""


def test_wsl_helpers_parse_addresses_rules_and_forwards(monkeypatch, cli):
    cli([("ip -4 -o addr show eth0", (0, "2: eth0    inet 172.20.9.9/20 brd 172.20.15.255 scope global eth0")), ("ip route show default", (0, "default via 172.20.0.1 dev eth0")),
         ("portproxy show", (0, PORTPROXY)), ("Get-NetFirewallRule", (0, "True\r\n"))])
    assert PC._get_wsl_ip() == "172.20.9.9" and PC._get_wsl_host_ip() == "172.20.0.1"
    assert PC._wsl_portproxy_target(4001) == "172.20.1.5" and PC._wsl_portproxy_target(4002) is None and PC._wsl_firewall_exists(4001)
    cli([("ip -4", (1, "")), ("hostname -I", (0, "172.20.3.3 fe80::1")), ("Get-NetFirewallRule", (0, "")), ("portproxy show", (1, ""))])
    assert PC._get_wsl_ip() == "172.20.3.3" and not PC._wsl_firewall_exists(4001) and PC._wsl_portproxy_target(4001) is None


def test_wsl_flow_repairs_a_stale_forward_without_asking(monkeypatch, cli, capsys):
    state = {"target": "172.20.1.5"}

    def show(_cmd):
        return 0, PORTPROXY.replace("172.20.1.5      4001", f"{state['target']}      4001")

    def elevate(cmd):
        if "portproxy add" in " ".join(cmd):
            state["target"] = "172.20.9.9"
        return 0, ""

    sh = cli([("ip -4 -o addr show eth0", (0, "inet 172.20.9.9/20")), ("portproxy show", show), ("Get-NetFirewallRule", (0, "True")), ("Start-Process", elevate)])
    assert PC._check_port_wsl(4001) is True
    out = capsys.readouterr().out
    assert "re-pointed from 172.20.1.5 to 172.20.9.9" in out and any("Start-Process" in " ".join(c) for c in sh.calls)
    assert PC._check_port_wsl(4001) is True and "forwards 4001/TCP to this VM" in capsys.readouterr().out      # now consistent: nothing to do


def test_wsl_flow_prints_manual_steps_when_not_interactive(monkeypatch, cli, capsys):
    cli([("ip -4", (0, "inet 172.20.9.9/20")), ("portproxy show", (0, "")), ("Get-NetFirewallRule", (0, ""))])
    monkeypatch.setattr(PC.sys.stdin, "isatty", lambda: False, raising=False)
    assert PC._check_port_wsl(4001) is True
    out = capsys.readouterr().out
    assert "netsh interface portproxy add v4tov4 listenport=4001" in out and "New-NetFirewallRule" in out


@pytest.mark.parametrize("bad", [0, -1, 65536, "80", 80.0, True, None])
def test_every_entry_point_validates_the_port(bad, cli):
    cli([])
    for fn in (PC._auto_open_azure, PC._auto_open_aws, PC._auto_open_gcp, PC._auto_open_wsl, PC._wsl_firewall_exists, PC._wsl_portproxy_target,
               PC._check_iptables_allows, PC.is_port_listening, PC._get_wsl_manual_instructions):
        with pytest.raises(ValueError):
            fn(bad)


@pytest.mark.parametrize("provider,needle", [(PC.CloudProvider.AZURE, "az vm open-port"), (PC.CloudProvider.AWS, "authorize-security-group-ingress"),
                                             (PC.CloudProvider.GCP, "gcloud compute firewall-rules create"), (PC.CloudProvider.UNKNOWN, "ufw allow 4001/tcp")])
def test_manual_instructions_name_the_port_and_the_tool(provider, needle):
    text = PC._get_manual_instructions(provider, 4001)
    assert needle in text and "4001" in text


def test_missing_clis_are_reported_not_raised(monkeypatch):
    monkeypatch.setattr(PC.shutil, "which", lambda name: None)
    assert PC._auto_open_azure(4001) == (False, "Azure CLI (az) is not installed.")
    assert PC._auto_open_aws(4001) == (False, "AWS CLI (aws) is not installed.")
    assert PC._auto_open_gcp(4001) == (False, "Google Cloud CLI (gcloud) is not installed.")
