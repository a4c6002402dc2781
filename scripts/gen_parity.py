#!/usr/bin/env python3
"""Regenerate the module table of docs/parity.md: every reference module -> its counterpart here -> the tests that
import it.  `python scripts/gen_parity.py /path/to/reference > /tmp/modules.md`"""
from __future__ import annotations

import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
RENAMES = {"dashboard/screens/__init__.py": None}


def main(ref_root: str) -> int:
    ref = Path(ref_root) / "infomesh"
    tests = {p.name: p.read_text() for p in (ROOT / "tests").glob("test_*.py")}
    rows, missing = [], []
    for f in sorted(ref.rglob("*.py")):
        rel = f.relative_to(ref).as_posix()
        if rel.endswith("__init__.py") and f.stat().st_size < 400:
            continue
        ours = ROOT / "infomesh_b200" / rel
        mod = "infomesh_b200." + rel[:-3].replace("/", ".").removesuffix(".__init__")
        leaf = mod.rsplit(".", 1)
        pats = [re.escape(mod) + r"\b", rf"from {re.escape(leaf[0])} import [^\n]*\b{re.escape(leaf[1])}\b"] if len(leaf) == 2 else [re.escape(mod)]
        used = sorted(n for n, src in tests.items() if any(re.search(p, src) for p in pats))
        if not ours.exists():
            missing.append(rel)
        rows.append(f"| `{rel}` | {'`infomesh_b200/' + rel + '`' if ours.exists() else '**missing**'} | {', '.join(u[5:-3] for u in used) or '—'} |")
    print("| Reference module (`infomesh/…`) | Here | Tests that import it (`tests/test_<name>.py`) |\n|---|---|---|")
    print("\n".join(rows))
    print(f"\n{len(rows)} modules, {len(missing)} without a same-path counterpart: {missing}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
