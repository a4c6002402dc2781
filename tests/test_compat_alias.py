"""``infomesh_b200.compat``: code written against the reference's import paths runs on this package."""
import subprocess
import sys
import textwrap

import pytest


@pytest.fixture()
def alias():
    from infomesh_b200 import compat

    assert compat.alias_as_infomesh()
    yield compat
    compat.remove_alias()


def test_alias_resolves_to_the_same_module_objects(alias):
    import infomesh
    import infomesh.config
    from infomesh.index.local_store import LocalStore
    from infomesh.search.query import search_local

    import infomesh_b200
    import infomesh_b200.config
    from infomesh_b200.index import local_store
    from infomesh_b200.search import query

    assert infomesh is infomesh_b200 and infomesh.config is infomesh_b200.config
    assert LocalStore is local_store.LocalStore and search_local is query.search_local
    assert alias.alias_as_infomesh()                      # idempotent
    with pytest.raises(ImportError):
        import infomesh.no_such_module  # noqa: F401


def test_a_reference_style_script_runs_unchanged(tmp_path):
    """The imports and calls of the reference's examples/basic_search.py and credit_status.py, verbatim in style, in a fresh
    interpreter with only INFOMESH_B200_ALIAS set."""
    script = tmp_path / "their_script.py"
    script.write_text(textwrap.dedent('''
        import infomesh_b200                      # noqa: F401 -- the one line a switching user adds (or sets the variable in a .pth)
        from infomesh.config import load_config
        from infomesh.credits.ledger import ActionType, CreditLedger
        from infomesh.index.local_store import LocalStore
        from infomesh.search.formatter import format_fts_results
        from infomesh.search.query import search_local

        config = load_config()
        store = LocalStore(db_path=config.index.db_path, compression_enabled=config.storage.compression_enabled,
                           compression_level=config.storage.compression_level)
        store.add_document(url="https://example.org/asyncio", title="asyncio", text="python asyncio event loop tutorial " * 10,
                           raw_html_hash="r", text_hash="t")
        print(format_fts_results(search_local(store, "python asyncio", limit=5)))
        store.close()
        ledger = CreditLedger(config.node.data_dir / "credits.db")
        ledger.record_action(ActionType.CRAWL, quantity=2)
        print("balance", round(ledger.balance(), 2))
        ledger.close()
    '''))
    env = {"INFOMESH_B200_ALIAS": "1", "INFOMESH_NODE_DATA_DIR": str(tmp_path / "data"), "PYTHONPATH": ":".join(sys.path), "PATH": "/usr/bin:/bin",
           "HOME": str(tmp_path)}
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-800:]
    assert "example.org/asyncio" in out.stdout and "balance 2.0" in out.stdout


def test_alias_is_not_installed_over_a_real_infomesh_package(tmp_path, monkeypatch):
    from infomesh_b200 import compat

    compat.remove_alias()
    fake = tmp_path / "infomesh"
    fake.mkdir()
    (fake / "__init__.py").write_text("MARK = 'the reference'\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    sys.modules.pop("infomesh", None)
    try:
        assert compat.alias_as_infomesh() is False                    # a real package of that name wins
        import infomesh

        assert infomesh.MARK == "the reference"
        assert compat.alias_as_infomesh(force=True) is True           # unless the caller insists
    finally:
        compat.remove_alias()
        sys.modules.pop("infomesh", None)
