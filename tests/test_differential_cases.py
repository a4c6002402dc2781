"""Differential tests against the installed, unmodified reference (baseline/_ref, see baseline/install_ref.py): every
pure-function call of scripts/scripts_diff_cases.py and every stateful scenario of scripts/diff_scenarios.py is ONE pytest
case -- same inputs through ``infomesh`` and ``infomesh_b200``, normalised outputs must be equal.  Skipped where the
reference is not installed (e.g. on a GPU box snapshot without baseline/_ref)."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
pytestmark = pytest.mark.skipif(not (REF / "infomesh" / "__init__.py").exists(), reason="reference not installed under baseline/_ref")

sys.path.insert(0, str(ROOT / "scripts"))
from diff_scenarios import SCENARIOS  # noqa: E402
from diff_vs_reference import _norm  # noqa: E402
from scripts_diff_cases import CASES  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def reference_importable():
    """Make ``import infomesh`` resolve to baseline/_ref with the stand-in wheels from baseline/shims (structlog, zstandard)."""
    added = [str(ROOT / "baseline" / "shims"), str(REF)]
    sys.path[:0] = added
    yield
    for p in added:
        if p in sys.path:
            sys.path.remove(p)


def _call(pkg: str, mod: str, fn: str, args, kwargs):
    f = getattr(importlib.import_module(f"{pkg}.{mod}"), fn)
    a = [x(pkg) if callable(x) else x for x in args]
    k = {n: (x(pkg) if callable(x) else x) for n, x in kwargs.items()}
    try:
        return "ok", _norm(f(*a, **k))
    except Exception as exc:  # noqa: BLE001 -- raising the same exception type IS matching behaviour
        return "raise", type(exc).__name__


# Deliberate, documented differences (docs/parity.md): `length` accepted positionally; this detector's confidence model;
# <updated> used for Atom dates; completions list the commands this CLI really has.  For these the case only checks that
# this implementation answers without raising.
DELIBERATE = {"hashing.short_hash", "crawler.lang_detect.detect_language", "crawler.rss.parse_feed_xml", "api.extensions.get_completion_commands"}


def _ids():
    for mod, fn, calls in CASES:
        for i, (args, kwargs) in enumerate(calls):
            yield pytest.param(mod, fn, args, kwargs, id=f"{mod}.{fn}[{i}]")


@pytest.mark.parametrize("mod,fn,args,kwargs", list(_ids()))
def test_pure_function_matches_reference(mod, fn, args, kwargs):
    try:
        importlib.import_module(f"infomesh.{mod}")
    except ImportError as exc:
        pytest.skip(f"reference module needs a wheel that is not installed: {exc}")
    ours, theirs = _call("infomesh_b200", mod, fn, args, kwargs), _call("infomesh", mod, fn, args, kwargs)
    if f"{mod}.{fn}" in DELIBERATE:
        assert ours[0] == "ok"
        if ours != theirs and isinstance(ours[1], dict) and isinstance(theirs[1], dict):        # the stable part still has to agree
            stable = {"language", "script", "title", "url", "feed_type"} & set(ours[1]) & set(theirs[1]) - {"script"}
            assert {k: ours[1][k] for k in stable} == {k: theirs[1][k] for k in stable}
        return
    assert ours == theirs


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario_matches_reference(name, tmp_path):
    outs = {}
    for pkg in ("infomesh", "infomesh_b200"):
        d = tmp_path / pkg
        d.mkdir()
        try:
            outs[pkg] = _norm(SCENARIOS[name](pkg, d))
        except ImportError as exc:
            pytest.skip(f"{pkg}: {exc}")
    ours, theirs = outs["infomesh_b200"], outs["infomesh"]
    if isinstance(ours, dict) and isinstance(theirs, dict):
        diff = {k: (theirs.get(k), ours.get(k)) for k in sorted(set(ours) | set(theirs)) if ours.get(k) != theirs.get(k)}
        assert not diff, f"(reference, ours) differ on: {diff}"
    assert ours == theirs


def test_url_validation_is_stricter_than_the_reference_where_it_matters():
    """Two inputs the reference accepts and this repo refuses on purpose (kept out of the sweep above): an IPv4-mapped IPv6
    loopback literal (an SSRF bypass) and a URL far beyond any sane length."""
    from infomesh.security import validate_url as theirs
    from infomesh_b200.security import SSRFError, validate_url as ours

    for url in ("http://[::ffff:127.0.0.1]/", "https://example.org/" + "a" * 3000):
        assert theirs(url)            # accepted upstream
        with pytest.raises(SSRFError):
            ours(url)
