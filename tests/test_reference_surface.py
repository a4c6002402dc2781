"""When a checkout of the reference is present, every public function / class / method / dataclass field / enum member it
exports must exist under the same path here (scripts/diff_vs_reference.py); skipped elsewhere (e.g. on the GPU box)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not (REF / "infomesh").is_dir(), reason="no reference checkout on this machine")
def test_no_public_name_of_the_reference_is_missing(tmp_path):
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_vs_reference.py"), "surface", str(REF)], capture_output=True, text=True,
                         timeout=600, cwd=tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.startswith("0 public names")


@pytest.mark.skipif(not (REF / "infomesh").is_dir(), reason="no reference checkout on this machine")
def test_pure_functions_agree_with_the_reference_except_for_documented_cases(tmp_path):
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_vs_reference.py"), "behaviour", str(REF)], capture_output=True, text=True,
                         timeout=900, cwd=tmp_path)
    diffs = [ln.split()[1].split("(")[0] for ln in out.stdout.splitlines() if ln.startswith("  DIFF ") or ln.startswith("  UNAVAILABLE ")]
    # deliberate: keyword `length` accepted positionally; our detector's confidence model; <updated> used for Atom dates;
    # completions list the commands this CLI really has
    allowed = {"hashing.short_hash", "crawler.lang_detect.detect_language", "crawler.rss.parse_feed_xml", "api.extensions.get_completion_commands"}
    assert set(diffs) <= allowed, out.stdout[-4000:]
    assert "calls compared" in out.stdout


@pytest.mark.skipif(not (REF / "infomesh").is_dir(), reason="no reference checkout on this machine")
def test_wire_bytes_hashes_ledger_maths_and_search_output_match_the_reference(tmp_path):
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_vs_reference.py"), "interop", str(REF)], capture_output=True, text=True,
                         timeout=900, cwd=tmp_path)
    assert out.returncode == 0 and "0 differences" in out.stdout, out.stdout[-4000:] + out.stderr[-2000:]


@pytest.mark.skipif(not (REF / "infomesh").is_dir(), reason="no reference checkout on this machine")
def test_mcp_tool_outputs_contain_everything_the_reference_returns(tmp_path):
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_vs_reference.py"), "mcp", str(REF)], capture_output=True, text=True,
                         timeout=900, cwd=tmp_path)
    assert out.returncode == 0 and "0 unexpected differences" in out.stdout, out.stdout[-4000:] + out.stderr[-2000:]


@pytest.mark.skipif(not (REF / "infomesh").is_dir(), reason="no reference checkout on this machine")
def test_admin_api_responses_contain_everything_the_reference_returns(tmp_path):
    pytest.importorskip("fastapi")
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_vs_reference.py"), "http", str(REF)], capture_output=True, text=True,
                         timeout=900, cwd=tmp_path)
    assert out.returncode == 0 and ", 0 differences" in out.stdout, out.stdout[-4000:] + out.stderr[-2000:]


def test_signatures_accept_the_reference_calling_conventions():
    """Parameter names, positional order and required-ness of every shared public callable (scripts/diff_signatures.py).  The only
    differences allowed are the transport-facing constructors / handlers that take a ``send`` callable or a decoded payload where
    the reference takes a libp2p host or stream."""
    ref_install = ROOT / "baseline" / "_ref"
    if not (ref_install / "infomesh").exists():
        pytest.skip("reference not installed under baseline/_ref")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "baseline" / "shims"), str(ref_install), str(ROOT)]))
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "diff_signatures.py")], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-1500:]
    heads = sorted(line.split("  ")[0] for line in out.stdout.splitlines() if ":" in line and "missing=" in line)
    assert heads == ["p2p.replication:Replicator.__init__", "p2p.replication:Replicator.handle_replicate_request",
                     "p2p.routing:QueryRouter.__init__", "p2p.routing:QueryRouter.handle_search_request"], out.stdout[-2000:]
