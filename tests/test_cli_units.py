"""Command-line paths that need a prepared data directory (model: reference tests/test_cli_*.py, test_version_check.py)."""
import gzip
import json
import os
import time

import pytest
from click.testing import CliRunner


@pytest.fixture()
def node_env(tmp_path, monkeypatch):
    monkeypatch.setenv("INFOMESH_NODE_DATA_DIR", str(tmp_path))
    monkeypatch.setenv("INFOMESH_INDEX_DB_PATH", str(tmp_path / "index.db"))
    monkeypatch.setenv("INFOMESH_INDEX_VECTOR_SEARCH", "false")
    monkeypatch.setenv("INFOMESH_NETWORK_BOOTSTRAP_DNS", "false")
    monkeypatch.setenv("INFOMESH_NETWORK_BOOTSTRAP_GITHUB", "false")
    import infomesh_b200.config as C

    monkeypatch.setattr(C, "DEFAULT_CONFIG_PATH", tmp_path / "config.toml")
    return tmp_path


def _run(*args):
    from infomesh_b200.cli import cli

    return CliRunner().invoke(cli, list(args))


_WET = """WARC/1.0\r
WARC-Type: warcinfo\r
Content-Length: 10\r
\r
software x\r
\r
WARC/1.0\r
WARC-Type: conversion\r
WARC-Target-URI: https://example.org/tmem\r
Content-Length: {n}\r
\r
{body}\r
\r
WARC/1.0\r
WARC-Type: conversion\r
WARC-Target-URI: https://example.org/short\r
Content-Length: 5\r
\r
tiny.\r
\r
"""


def test_import_wet_from_a_local_gzip(node_env):
    body = "Tensor memory holds the accumulators of the fifth generation tensor cores, and is read back with tcgen05 loads. " * 4
    p = node_env / "sample.warc.wet.gz"
    p.write_bytes(gzip.compress(_WET.format(n=len(body.encode()), body=body).encode()))
    out = _run("index", "import-wet", str(p))
    assert out.exit_code == 0, out.output
    assert "Imported 1 documents from WET file" in out.output and "Total records:    1" in out.output and "Skipped (short):  0" in out.output
    again = _run("index", "import-wet", str(p))
    assert "Imported 0 documents from WET file" in again.output and "Skipped (dup):    1" in again.output
    hit = _run("search", "--local", "tensor memory accumulators")
    assert "example.org/tmem" in hit.output


def test_import_needs_a_path_or_starter_and_starter_fails_cleanly_offline(node_env, monkeypatch):
    from infomesh_b200.index import starter as S

    missing = _run("index", "import")
    assert missing.exit_code == 1 and "Missing argument 'INPUT_PATH'" in missing.output and "--starter" in missing.output

    async def none(**_):
        return None

    monkeypatch.setattr(S, "find_starter_asset", none)
    r = _run("index", "import", "--starter")
    assert r.exit_code == 0 and "No starter snapshot found in GitHub Releases." in r.output

    async def found(**_):
        return S.StarterAssetInfo("https://x/s.snapshot", 5 * 2 ** 20, "starter-2026.09", "2026-09-01")

    monkeypatch.setattr(S, "find_starter_asset", found)
    info = _run("index", "import", "--starter", "--info")
    assert info.exit_code == 0 and "Starter Index (remote)" in info.output and "Release:  starter-2026.09" in info.output and "Size:     5.0 MB" in info.output
    monkeypatch.setattr(S, "download_starter_sync", lambda d, progress_callback=None: None)
    failed = _run("index", "import", "--starter")
    assert failed.exit_code == 1 and "Found starter index: 5.0 MB (release starter-2026.09)" in failed.output and "Download failed." in failed.output


def test_feedback_commands_show_recorded_signals(node_env):
    from infomesh_b200.search.feedback import FeedbackStore

    fb = FeedbackStore(str(node_env / "feedback.db"))
    for _ in range(3):
        fb.record_fetch("tmem", "https://example.org/good", 1)
    fb.record_fetch("tmem", "https://example.org/ok", 2)
    fb.close()
    st = _run("feedback", "stats")
    assert "Total signals: 4" in st.output and "Boosted URLs:  2" in st.output and st.output.splitlines()[5].split()[-2:] == ["3", "0"]      # fetch, cite
    top = _run("feedback", "top-urls", "-n", "1")
    assert "example.org/good  (fetch=3 cite=0)" in top.output and "example.org/ok" not in top.output


def test_status_reports_a_live_node_and_its_p2p_state(node_env):
    from infomesh_b200 import runtime as RT

    RT.write_pid_file(node_env, os.getpid())
    RT.write_runtime_status(node_env, {"status": "running", "pid": os.getpid(), "degrade_level": 0, "cpu_percent": 12.0, "memory_percent": 30.0,
                                      "process_memory_mb": 210, "updated_at": time.time(),
                                      "gpu": {"documents": 1000, "hbm_bytes": 64 * 2 ** 20, "cuda_graph": True}})
    (node_env / "p2p_status.json").write_text(json.dumps({"timestamp": time.time(), "state": "running", "peers": 0, "listen_addrs": ["/ip4/0.0.0.0/tcp/4001"],
                                                          "bootstrap": {"connected": 0, "configured": 2}}))
    out = _run("status").output
    assert "Running:         yes" in out and "cpu 12.0%" in out and "GPU index:       1000 docs · 64 MB HBM · graph on" in out
    assert "P2P:             running, 0 peer(s)" in out and "0 connected / 2 configured" in out and "hint:" in out
    RT.clear_pid_file(node_env, os.getpid())
    out = _run("status").output
    assert "Running:         no" in out and "P2P:             stopped" in out
    assert "Credits:         0.0 (earned 0.0 / spent 0.0)" in out and "Tier:            tier_1 (score 0.0, search cost 0.100)" in out
    assert "GitHub:          not connected" in out and "infomesh config github" in out


def test_stop_cleans_up_a_stale_pid_file(node_env):
    from infomesh_b200 import runtime as RT

    (node_env / "infomesh.pid").write_text("999999999")
    r = _run("stop")
    assert r.exit_code == 0 and ("No running InfoMesh node found." in r.output or "stale PID" in r.output)
    assert RT.read_live_pid(node_env) is None


def test_start_refuses_when_a_node_is_already_running(node_env):
    from infomesh_b200 import runtime as RT

    RT.write_pid_file(node_env, os.getpid())
    r = _run("start", "--background")
    assert r.exit_code == 0 and f"already running (PID {os.getpid()})" in r.output
    RT.clear_pid_file(node_env, os.getpid())
    lock = RT.StartupLock(node_env)
    assert lock.acquire()
    try:
        assert "startup is already in progress" in _run("start", "--background").output
    finally:
        lock.release()


def test_serve_command_line_is_assembled_from_the_options():
    import sys

    from infomesh_b200.cli.serve import _serve_cmd

    assert _serve_cmd(None, None) == [sys.executable, "-m", "infomesh_b200", "_serve"]
    assert _serve_cmd("tech-docs", "crawler", no_crawl=True)[4:] == ["--seeds", "tech-docs", "--role", "crawler", "--no-crawl"]


def test_update_check_uses_the_cached_pypi_answer(node_env, monkeypatch):
    from infomesh_b200 import __version__, version_check as V

    monkeypatch.setattr(V, "_fetch_latest_from_pypi", lambda: None)
    assert "Already up to date" in _run("update", "--check").output           # offline: nothing known
    V._write_cache(node_env, "999.0.0")
    r = _run("update", "--check")
    assert f"v{__version__} → v999.0.0" in r.output and "Run: infomesh update" in r.output
    (node_env / V._CACHE_FILE_NAME).write_text(json.dumps({"version": "999.0.0", "ts": 0}))      # expired cache is ignored
    assert V.check_pypi_update(node_env) is None
    monkeypatch.setattr(V, "_fetch_latest_from_pypi", lambda: "0.0.1")
    assert V.check_pypi_update(node_env) is None and V._read_cache(node_env) == "0.0.1"
    tr = V.PeerVersionTracker()
    tr.record("p1", "1000.2.0"), tr.record("p2", "x" * 40), tr.record("p3", "")
    assert tr.peer_versions == {"p1": "1000.2.0"}
    best = V.check_for_update(node_env, tr)
    assert best.source == "peer" and "P2P peer" in V.format_update_banner(best)
    assert V._parse_version("1.2.3rc1") == (1, 2, 3) and V._parse_version("garbage") == (0,)


def test_search_command_formats_scores_and_falls_back_without_gpu(node_env):
    from infomesh_b200.sdk.client import InfoMeshClient

    with InfoMeshClient(str(node_env)) as c:
        for i in range(4):
            c.add_document(f"https://example.org/doc{i}", f"Cluster launch {i}", f"Thread block clusters share distributed shared memory, part {i}. " * 6)
    r = _run("search", "--local", "-n", "2", "distributed shared memory clusters")
    assert r.exit_code == 0 and r.output.count("https://example.org/doc") == 2
    assert "ms" in r.output and "[1] Cluster launch" in r.output and "    Source: https://example.org/doc" in r.output and "BM25=" in r.output


def test_updates_target_this_distribution_not_the_reference_package(node_env, monkeypatch):
    """``infomesh update`` must upgrade ``infomesh-b200``: installing the reference's ``infomesh`` over this package would replace it."""
    import subprocess

    from infomesh_b200 import DISTRIBUTION, version_check as V

    assert DISTRIBUTION == "infomesh-b200" and V._PYPI_URL.endswith("/infomesh-b200/json")
    V._write_cache(node_env, "999.0.0")
    seen = []
    monkeypatch.setattr(subprocess, "run", lambda cmd, **kw: seen.append(cmd) or subprocess.CompletedProcess(cmd, 0, "", ""))
    _run("update")
    assert seen and seen[0][-3:] == ["install", "--upgrade", "infomesh-b200"]
