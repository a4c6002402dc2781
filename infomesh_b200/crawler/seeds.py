"""Seed URL lists (``seeds/*.txt``: one URL per line, ``#`` comments) — reference infomesh/crawler/seeds.py:18-76."""
from __future__ import annotations

import sysconfig
from pathlib import Path

CATEGORIES = ("tech-docs", "academic", "encyclopedia", "quickstart", "search-strategy")


def _seed_dirs() -> list[Path]:
    here = Path(__file__).resolve().parent.parent
    return [here / "seeds", here.parent / "seeds", Path(sysconfig.get_path("data") or "") / "share" / "infomesh" / "seeds"]


def _parse_seed_file(path: Path) -> list[str]:
    out = []
    for line in path.read_text("utf-8", errors="replace").splitlines():
        line = line.split("#", 1)[0].strip()
        if line.startswith(("http://", "https://")):
            out.append(line)
    return out


def load_seeds(category: str | None = None, seeds_dir: Path | None = None) -> list[str]:
    """All seeds of one category (file stem) or of every file; duplicates removed, order kept."""
    dirs = [Path(seeds_dir)] if seeds_dir else _seed_dirs()
    for d in dirs:
        if not d.is_dir():
            continue
        files = [d / f"{category}.txt"] if category else sorted(d.glob("*.txt"))
        urls: list[str] = []
        for f in files:
            if f.is_file():
                urls.extend(_parse_seed_file(f))
        if urls:
            return list(dict.fromkeys(urls))
    return []
