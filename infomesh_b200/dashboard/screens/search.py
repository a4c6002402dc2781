"""Search tab: a query box over the local index (reference infomesh/dashboard/screens/search.py:22-216)."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Input, Static


class SearchPane(Vertical):
    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config

    def compose(self) -> ComposeResult:
        yield Input(placeholder="search the local index and press Enter", id="se-input")
        yield Static("[dim]results appear here[/]", id="se-results")

    def run_query(self, query: str) -> str:
        from infomesh_b200.index.local_store import LocalStore
        from infomesh_b200.search.query import search_local

        cfg = self.config
        with LocalStore(db_path=cfg.index.db_path, compression_enabled=cfg.storage.compression_enabled,
                        compression_level=cfg.storage.compression_level) as st:
            res = search_local(st, query, limit=8)
        if not res.results:
            return "No results found."
        rows = [f"[dim]{res.total} results in {res.elapsed_ms:.0f} ms[/]"]
        for i, r in enumerate(res.results, 1):
            snip = r.snippet.replace("<b>", "[bold yellow]").replace("</b>", "[/]").replace("\n", " ")[:180]
            rows.append(f"[bold cyan]{i}. {r.title or r.url}[/]  [dim]{r.combined_score:.3f}[/]\n   {r.url}\n   {snip}")
        return "\n".join(rows)

    def on_input_submitted(self, event: Input.Submitted) -> None:
        q = event.value.strip()
        if q:
            try:
                self.query_one("#se-results", Static).update(self.run_query(q))
            except Exception as exc:  # noqa: BLE001
                self.query_one("#se-results", Static).update(f"[red]search failed: {exc}[/]")
