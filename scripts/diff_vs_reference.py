#!/usr/bin/env python3
"""Developer tool: differential check against the reference implementation, when a checkout is available.

    python scripts/diff_vs_reference.py surface   [/root/reference]     # public names the reference exports that we lack
    python scripts/diff_vs_reference.py behaviour [/root/reference]     # same inputs through both, outputs compared
    python scripts/diff_vs_reference.py interop   [/root/reference]     # wire bytes, hashes, ledger/trust maths, search output
    python scripts/diff_vs_reference.py mcp       [/root/reference]     # text returned by every MCP tool handler
    python scripts/diff_vs_reference.py http      [/root/reference]     # every route of the local admin API

The reference is imported read-only with tiny stand-ins for the logging / compression wheels that are not installed
here; nothing from it is copied.  Exit code 1 when differences are found."""
from __future__ import annotations

import dataclasses
import enum
import importlib
import inspect
import pkgutil
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _install_stubs() -> None:
    class _Any:
        """Absorbs any attribute access or call (the reference only configures logging with these)."""

        def __getattr__(self, _n):
            return self

        def __call__(self, *a, **k):
            return self

    class _Stub(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith("__"):            # inspect / importlib probe modules for __file__, __path__, ...
                raise AttributeError(n)
            return _Any()

    sys.modules.setdefault("structlog", _Stub("structlog"))
    for optional in ("trafilatura",):                 # imported at module level by the reference's parser
        try:
            importlib.import_module(optional)
        except ImportError:
            sys.modules[optional] = _Stub(optional)
    try:
        import zstandard  # noqa: F401
    except ImportError:
        z = types.ModuleType("zstandard")
        for n in ("ZstdCompressor", "ZstdDecompressor", "ZstdCompressionDict"):
            setattr(z, n, type(n, (), {"__init__": lambda self, *a, **k: None}))
        z.ZstdError = type("ZstdError", (Exception,), {})
        z.train_dictionary = lambda *a, **k: None
        sys.modules["zstandard"] = z


def _ref_modules(ref_root: Path):
    sys.path.insert(0, str(ref_root))
    import infomesh  # type: ignore

    for m in pkgutil.walk_packages(infomesh.__path__, "infomesh.", onerror=lambda _n: None):
        yield m.name


def surface(ref_root: Path) -> int:
    missing, skipped = [], []
    for name in _ref_modules(ref_root):
        try:
            ref = importlib.import_module(name)
        except Exception as exc:  # noqa: BLE001 — optional wheels (trio, libp2p, textual screens, ...)
            skipped.append(f"{name}: {type(exc).__name__}: {str(exc)[:60]}")
            continue
        ours_name = "infomesh_b200" + name[len("infomesh"):]
        try:
            ours = importlib.import_module(ours_name)
        except Exception as exc:  # noqa: BLE001
            missing.append(f"{ours_name}: cannot import ({exc})")
            continue
        for attr, obj in vars(ref).items():
            if attr.startswith("_") or getattr(obj, "__module__", name) != name or not (inspect.isfunction(obj) or inspect.isclass(obj)):
                continue
            if not hasattr(ours, attr):
                missing.append(f"{ours_name}.{attr}")
                continue
            if inspect.isclass(obj):
                mine = getattr(ours, attr)
                for meth, mobj in vars(obj).items():
                    if not meth.startswith("_") and (inspect.isfunction(mobj) or isinstance(mobj, (property, staticmethod, classmethod))) \
                            and not hasattr(mine, meth):
                        missing.append(f"{ours_name}.{attr}.{meth}")
                if dataclasses.is_dataclass(obj) and dataclasses.is_dataclass(mine):
                    want, have = [f.name for f in dataclasses.fields(obj)], [f.name for f in dataclasses.fields(mine)]
                    for f in want:
                        if f not in have:
                            missing.append(f"{ours_name}.{attr} field {f}")
                if isinstance(obj, enum.EnumMeta) and isinstance(mine, enum.EnumMeta):
                    for member in obj.__members__:
                        if member not in mine.__members__:
                            missing.append(f"{ours_name}.{attr}.{member}")
    print(f"{len(missing)} public names of the reference are missing here; {len(skipped)} reference modules could not be imported")
    for m in missing:
        print("  MISSING", m)
    for s in skipped:
        print("  skipped", s)
    return 1 if missing else 0


def _norm(x, depth=0):
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        return {k: _norm(v, depth + 1) for k, v in dataclasses.asdict(x).items()}
    if isinstance(x, enum.Enum):
        return _norm(x.value)
    if isinstance(x, float):
        return round(x, 9)
    if isinstance(x, (list, tuple)):
        return [_norm(v, depth + 1) for v in x]
    if isinstance(x, (set, frozenset)):
        return sorted((_norm(v, depth + 1) for v in x), key=repr)
    if isinstance(x, dict):
        return {str(k): _norm(v, depth + 1) for k, v in sorted(x.items(), key=lambda kv: str(kv[0]))}
    if isinstance(x, (str, int, bool, bytes)) or x is None:
        return x
    if hasattr(x, "__dict__"):
        return {k: _norm(v, depth + 1) for k, v in vars(x).items() if not k.startswith("_")}
    return repr(x)


def behaviour(ref_root: Path) -> int:
    from scripts_diff_cases import CASES  # type: ignore

    list(_ref_modules(ref_root))[:0]
    bad = total = 0
    for mod, fn, calls in CASES:
        try:
            rf = getattr(importlib.import_module("infomesh." + mod), fn)
            of = getattr(importlib.import_module("infomesh_b200." + mod), fn)
        except Exception as exc:  # noqa: BLE001
            print(f"  UNAVAILABLE {mod}.{fn}: {exc}")
            bad += 1
            continue
        for args, kwargs in calls:
            total += 1
            out = []
            for f, pkg in ((rf, "infomesh"), (of, "infomesh_b200")):
                try:        # callables among the arguments build package-specific objects (dataclasses of that side)
                    a = [x(pkg) if callable(x) else x for x in args]
                    k = {n: (x(pkg) if callable(x) else x) for n, x in kwargs.items()}
                    out.append(("ok", _norm(f(*a, **k))))
                except Exception as exc:  # noqa: BLE001
                    out.append(("raise", type(exc).__name__))
            if out[0] != out[1]:
                bad += 1
                print(f"  DIFF {mod}.{fn}{args!r}{kwargs or ''}\n      ref : {str(out[0])[:300]}\n      ours: {str(out[1])[:300]}")
    print(f"{total} calls compared, {bad} differences")
    return 1 if bad else 0


def main() -> int:
    mode = sys.argv[1] if len(sys.argv) > 1 else "surface"
    ref_root = Path(sys.argv[2] if len(sys.argv) > 2 else "/root/reference")
    if not (ref_root / "infomesh").is_dir():
        print(f"no reference checkout at {ref_root}; nothing to compare")
        return 0
    _install_stubs()
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    with tempfile.TemporaryDirectory() as tmp:
        import os

        os.environ.setdefault("INFOMESH_NODE_DATA_DIR", tmp)
        if mode == "http":
            sys.path.insert(0, str(ref_root))
            import diff_http

            return diff_http.run()
        if mode == "mcp":
            sys.path.insert(0, str(ref_root))
            import diff_mcp

            return diff_mcp.run()
        if mode == "interop":
            sys.path.insert(0, str(ref_root))
            import diff_interop

            return diff_interop.run()
        return surface(ref_root) if mode == "surface" else behaviour(ref_root)


if __name__ == "__main__":
    raise SystemExit(main())
