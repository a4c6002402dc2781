"""Run code written against the reference package unchanged.

``infomesh_b200`` keeps the reference's module layout and public names (``docs/parity.md``), so the only thing that stops a
script with ``from infomesh.search.query import search_local`` from running on this package is the top-level name.
:func:`alias_as_infomesh` installs an import hook that resolves ``infomesh`` and every ``infomesh.<module>`` to the SAME module
objects as ``infomesh_b200.<module>`` (no copies: ``sys.modules["infomesh.config"] is sys.modules["infomesh_b200.config"]``)::

    import infomesh_b200.compat; infomesh_b200.compat.alias_as_infomesh()
    from infomesh.index.local_store import LocalStore          # -> infomesh_b200.index.local_store

or, without touching the script: ``INFOMESH_B200_ALIAS=1 python -c "import infomesh_b200; ..."`` (the package installs the alias
on import when that variable is set).  The hook refuses to shadow a real ``infomesh`` distribution unless ``force=True``."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys

_ALIAS, _REAL = "infomesh", "infomesh_b200"


def _real_name(alias: str) -> str:
    return _REAL + alias[len(_ALIAS):]


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``infomesh[.x.y]`` -> the module object of ``infomesh_b200[.x.y]``."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != _ALIAS and not fullname.startswith(_ALIAS + "."):
            return None
        try:
            real = importlib.util.find_spec(_real_name(fullname))
        except (ImportError, ValueError):
            return None
        if real is None:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=real.submodule_search_locations is not None)

    def create_module(self, spec):
        return importlib.import_module(_real_name(spec.name))        # the one and only module object, under a second name

    def exec_module(self, module):                                   # already executed under its real name
        return None


def _reference_installed() -> bool:
    for finder in sys.meta_path:
        if isinstance(finder, _AliasFinder):
            continue
        try:
            if finder.find_spec(_ALIAS, None) is not None:
                return True
        except Exception:  # noqa: BLE001 -- a foreign finder that cannot answer is not a reason to fail
            continue
    return False


def alias_as_infomesh(*, force: bool = False) -> bool:
    """Install the hook (idempotent).  Returns False -- and does nothing -- when a real ``infomesh`` package is importable and
    ``force`` is not set: silently replacing an installed reference would be worse than an ImportError."""
    if any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        return True
    if not force and (_ALIAS in sys.modules and not sys.modules[_ALIAS].__name__.startswith(_REAL) or _reference_installed()):
        return False
    sys.meta_path.insert(0, _AliasFinder())
    return True


def remove_alias() -> None:
    """Uninstall the hook and forget the aliased names (the real modules stay loaded)."""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _AliasFinder)]
    for name in [n for n in sys.modules if n == _ALIAS or n.startswith(_ALIAS + ".")]:
        mod = sys.modules[name]
        if getattr(mod, "__name__", "").startswith(_REAL):
            del sys.modules[name]
