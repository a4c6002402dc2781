"""Declarative assets stay loadable and consistent with the code they configure (ports, env names, commands)."""
import json
import re
import tomllib
from pathlib import Path

import pytest
import yaml

ROOT = Path(__file__).resolve().parent.parent


def _env_names(text):
    return set(re.findall(r"INFOMESH_[A-Z0-9_]+", text))


def test_every_infomesh_env_var_in_deploy_files_maps_to_a_config_field():
    from dataclasses import fields

    from infomesh_b200.config import Config

    known = {"INFOMESH_API_KEY", "INFOMESH_IMAGE", "INFOMESH_VENV", "INFOMESH_PACKAGE", "INFOMESH_MCP_URL", "INFOMESH_ADMIN_URL", "INFOMESH_B200_PDL"}
    cfg = Config()
    for sec in fields(cfg):
        for f in fields(getattr(cfg, sec.name)):
            known.add(f"INFOMESH_{sec.name.upper()}_{f.name.upper()}")
    files = [ROOT / "Dockerfile", ROOT / "docker-compose.yml", *ROOT.glob("deploy/**/*"), *ROOT.glob("k8s/*.yaml")]
    for f in files:
        if f.is_file():
            unknown = _env_names(f.read_text()) - known
            assert not unknown, f"{f.relative_to(ROOT)} sets unknown settings {sorted(unknown)}"


def test_compose_files_parse_and_expose_the_documented_ports():
    top = yaml.safe_load((ROOT / "docker-compose.yml").read_text())
    assert {"4001:4001", "8080:8080", "8081:8081"} <= set(top["services"]["node1"]["ports"])
    split = yaml.safe_load((ROOT / "deploy/docker-compose.yml").read_text())
    assert split["services"]["search"]["environment"]["INFOMESH_NODE_ROLE"] == "search"
    assert split["services"]["crawler"]["environment"]["INFOMESH_NETWORK_INDEX_SUBMIT_PEERS"].endswith(":8080")
    assert split["networks"]["private"] == {"internal": True}
    assert "dev" in yaml.safe_load((ROOT / "deploy/docker-compose.dev.yml").read_text())["services"]


def test_roles_used_by_deploy_files_are_real_roles():
    from infomesh_b200.config import NodeRole

    roles = {NodeRole.FULL, NodeRole.CRAWLER, NodeRole.SEARCH}
    for f in [ROOT / "deploy/docker-compose.yml", ROOT / "deploy/fly.toml", ROOT / "deploy/helm/infomesh/templates/deployment.yaml", ROOT / "deploy/terraform/main.tf"]:
        for r in re.findall(r'(?:INFOMESH_NODE_ROLE[=:]\s*"?|--role",\s*")([a-z]+)', f.read_text()):
            assert r in roles, (f.name, r)


def test_helm_chart_and_k8s_manifests_are_well_formed():
    chart = yaml.safe_load((ROOT / "deploy/helm/infomesh/Chart.yaml").read_text())
    vals = yaml.safe_load((ROOT / "deploy/helm/infomesh/values.yaml").read_text())
    assert chart["apiVersion"] == "v2" and vals["search"]["gpusPerPod"] == 8 and vals["service"]["adminPort"] == 8080
    tpl = "".join(p.read_text() for p in (ROOT / "deploy/helm/infomesh/templates").glob("*"))
    for ref in re.findall(r"\.Values\.([A-Za-z0-9_.]+)", tpl):
        node = vals
        for part in ref.split("."):
            assert isinstance(node, dict) and part in node, f".Values.{ref} is not defined in values.yaml"
            node = node[part]
    assert "{{" not in re.sub(r"\{\{.*?\}\}", "", tpl, flags=re.S)                   # every action is closed
    for f in (ROOT / "k8s").glob("*.yaml"):
        docs = [d for d in yaml.safe_load_all(f.read_text()) if d]
        assert docs and all("kind" in d and "apiVersion" in d for d in docs), f.name


def test_platform_descriptors_and_workflows_parse():
    fly = tomllib.loads((ROOT / "deploy/fly.toml").read_text())
    assert fly["http_service"]["internal_port"] == 8081 and fly["env"]["INFOMESH_GPU_ENABLED"] == "false"
    rw = json.loads((ROOT / "deploy/railway.json").read_text())
    assert rw["deploy"]["healthcheckPath"] == "/health" and "_serve" in rw["deploy"]["startCommand"]
    assert json.loads((ROOT / "smithery.json").read_text())
    for wf in (ROOT / ".github/workflows").glob("*.yml"):
        doc = yaml.safe_load(wf.read_text())
        assert doc["jobs"] and (True in doc or "on" in doc), wf.name           # PyYAML reads the bare key `on` as True
    ci = (ROOT / ".github/workflows/ci.yml").read_text()
    assert 'pytest tests/ -x -q -m "not gpu"' in ci and "pytest tests/ -x -q -m gpu" in ci and "g.build()" in ci
    tf = (ROOT / "deploy/terraform/main.tf").read_text()
    assert tf.count("{") == tf.count("}") and "INFOMESH_NETWORK_INDEX_SUBMIT_PEERS" in tf


def test_systemd_units_and_scripts_reference_existing_entry_points():
    unit = (ROOT / "deploy/infomesh.service").read_text()
    assert "infomesh_b200" in unit and "[Install]" in unit
    assert "OnCalendar" in (ROOT / "deploy/infomesh-update.timer").read_text()
    sh = (ROOT / "scripts/infomesh-update.sh").read_text()
    assert sh.startswith("#!/usr/bin/env bash") and "set -euo pipefail" in sh and "update --check" in sh
    import ast

    for py in list((ROOT / "scripts").glob("*.py")) + list((ROOT / "examples").glob("*.py")):
        ast.parse(py.read_text(), filename=str(py))
    pkg = json.loads((ROOT / "examples/typescript/package.json").read_text())
    assert "@modelcontextprotocol/sdk" in pkg["dependencies"]
    for ts in ("mcp_client.ts", "http_client.ts"):
        src = (ROOT / "examples/typescript" / ts).read_text()
        assert src.count("(") == src.count(")") and src.count("{") == src.count("}")


def test_example_tool_names_and_routes_exist():
    from infomesh_b200.api.extensions import generate_openapi_spec
    from infomesh_b200.mcp.tools import get_all_tools

    names = {t.name for t in get_all_tools()}
    ts = (ROOT / "examples/typescript/mcp_client.ts").read_text()
    assert set(re.findall(r'call\("([a-z_]+)"', ts)) <= names
    paths = set(generate_openapi_spec()["paths"])
    http = (ROOT / "examples/typescript/http_client.ts").read_text()
    for route in re.findall(r'get(?:<[^(]*>)?\(\s*[`"](/[a-z/]+)', http):
        assert route in paths or route == "/gpu/stats", route


def test_bootstrap_nodes_and_seed_lists_ship_inside_the_package():
    nodes = json.loads((ROOT / "infomesh_b200/bootstrap/nodes.json").read_text())
    assert isinstance(nodes, (list, dict))
    assert len(list((ROOT / "infomesh_b200/seeds").glob("*.txt"))) == 5
    pyproject = tomllib.loads((ROOT / "pyproject.toml").read_text())
    assert pyproject["project"]["scripts"]["infomesh"] == "infomesh_b200.cli:cli"
