#!/usr/bin/env python
"""Install the UNMODIFIED reference package into ``baseline/_ref`` (git-ignored; travels to the GPU box with gpurun).

The prescribed command

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

fails in this image: the reference's build backend ``hatchling`` is not in the wheelhouse.  The reference is a pure
Python ``hatchling`` wheel whose only payload is the ``infomesh/`` package (pyproject.toml ``[tool.hatch.build.targets.
wheel] packages = ["infomesh"]``), so an install IS a byte-for-byte copy of that directory.  This script does that copy
and writes ``baseline/_ref/MANIFEST.json`` (sha256 per file) so anyone can check nothing was modified:

    python baseline/install_ref.py            # copy (idempotent)
    python baseline/install_ref.py --verify   # re-hash against /root/reference (when it is mounted)

Third-party wheels the reference imports but this image lacks (structlog, zstandard, chromadb, sentence-transformers)
are NOT installed into _ref; library-only stand-ins live in ``baseline/shims`` (tracked) and go on sys.path after it."""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
DST = HERE / "_ref"


def _hashes(root: Path) -> dict[str, str]:
    out = {}
    for p in sorted(root.rglob("*")):
        if p.is_file() and "__pycache__" not in p.parts:
            out[str(p.relative_to(root))] = hashlib.sha256(p.read_bytes()).hexdigest()
    return out


def install(force: bool = False) -> str:
    pkg = DST / "infomesh"
    if pkg.is_dir() and not force:
        return "present"
    if not (REF_SRC / "infomesh").is_dir():
        return "reference source not mounted"
    if pkg.exists():
        shutil.rmtree(pkg)
    DST.mkdir(parents=True, exist_ok=True)
    shutil.copytree(REF_SRC / "infomesh", pkg, ignore=shutil.ignore_patterns("__pycache__"))
    for extra in ("seeds", "bootstrap"):            # wheel shared-data the package reads at run time
        if (REF_SRC / extra).is_dir() and not (DST / extra).exists():
            shutil.copytree(REF_SRC / extra, DST / extra)
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(REF_SRC), "method": "copytree (pure-Python wheel payload)",
                                                   "files": _hashes(pkg)}, indent=1))
    return "installed"


def verify() -> int:
    want = _hashes(REF_SRC / "infomesh")
    have = _hashes(DST / "infomesh")
    bad = [k for k in want if have.get(k) != want[k]] + [k for k in have if k not in want]
    print(f"{len(want)} files, {len(bad)} differ")
    for b in bad[:20]:
        print("  DIFF", b)
    return 1 if bad else 0


if __name__ == "__main__":
    if "--verify" in sys.argv:
        sys.exit(verify())
    print(install(force="--force" in sys.argv))
